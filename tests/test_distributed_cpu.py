"""World-size-2 (gloo, CPU) test of the data-parallel host logic of TrainEngine: parameter broadcast,
ONE flat gradient all-reduce per step (loss riding in the same buffer), identical fused AdamW on
every rank, per-rank BatchNorm statistics.  Kernels are emulated by oracle/abi_reference.py."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.helpers import cfg, make_pair


class _Patch:
    def setattr(self, obj, name, val, raising=True):
        setattr(obj, name, val)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, result_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gnnepcsaft_b200 import _lib
    from gnnepcsaft_b200.data.synth import batch_from_packed, make_packed_graphs
    from gnnepcsaft_b200.train import engine as E
    from gnnepcsaft_b200.train import models as M
    from oracle import abi_reference as R
    from oracle import pna_oracle as O
    R.install(_Patch(), _lib, M)
    E._require_cuda = lambda t, what: None

    c = cfg(propagation_depth=1, hidden_dim=16, warmup_steps=4)
    packed = make_packed_graphs(8, seed=9)
    deg = packed.in_degree_histogram()
    # rank-dependent initial weights: the engine must broadcast rank 0's
    o32, _, m = make_pair(c, deg, seed=rank, dtype64=False)
    engine = E.TrainEngine.from_config(m, c)
    assert engine.world == world
    halves = [torch.arange(0, 4), torch.arange(4, 8)]
    losses = []
    for _ in range(2):
        losses.append(float(engine.step(batch_from_packed(packed, halves[rank]))))
    flat = torch.cat([p.data.reshape(-1) for p in m.parameters()])
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    if rank == 0:
        assert torch.equal(gathered[0], gathered[1]), "ranks diverged"
        # reference: two replicas with rank 0's weights, per-replica BN statistics, averaged gradients
        ref0, _, _ = make_pair(c, deg, seed=0, dtype64=False)
        ref1 = O.create_model(c, deg)
        ref1.load_state_dict(ref0.state_dict())
        ref0.train(), ref1.train()
        opt = torch.optim.AdamW(ref0.parameters(), lr=c.learning_rate, weight_decay=c.weight_decay, amsgrad=True,
                                eps=1e-5)
        sched = torch.optim.lr_scheduler.CosineAnnealingWarmRestarts(opt, c.warmup_steps)
        start = torch.cat([p.data.reshape(-1).clone() for p in ref0.parameters()])
        ref_losses = []
        for _ in range(2):
            ls = []
            for r, ref in enumerate((ref0, ref1)):
                ref.zero_grad()
                b = batch_from_packed(packed, halves[r])
                loss = O.mape(ref(b), b.para.view(-1, c.num_para))
                loss.backward()
                ls.append(float(loss))
            for p0, p1 in zip(ref0.parameters(), ref1.parameters()):
                p0.grad = 0.5 * (p0.grad + p1.grad)
            opt.step()
            sched.step()
            with torch.no_grad():
                for p0, p1 in zip(ref0.parameters(), ref1.parameters()):
                    p1.copy_(p0)
            ref_losses.append(sum(ls) / world)
        end = torch.cat([p.data.reshape(-1) for p in ref0.parameters()])
        rel = float((gathered[0] - end).norm() / (end - start).norm())
        ok = rel < 5e-2 and all(abs(a - b) <= 1e-4 * abs(b) + 1e-6 for a, b in zip(losses, ref_losses))
        with open(result_path, "w") as f:
            f.write(f"{int(ok)} rel={rel} losses={losses} ref={ref_losses}")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_data_parallel_step(tmp_path):
    result = tmp_path / "result.txt"
    mp.spawn(_worker, args=(2, _free_port(), str(result)), nprocs=2, join=True)
    text = result.read_text()
    assert text.startswith("1"), text


def _store_worker(rank, world, port, result_path):
    """Store-resident epoch, sharded DistributedSampler-style over two ranks (padded compact batches)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gnnepcsaft_b200 import _lib
    from gnnepcsaft_b200.data.compact import CompactGraphStore
    from gnnepcsaft_b200.data.synth import make_packed_graphs
    from gnnepcsaft_b200.train import engine as E
    from gnnepcsaft_b200.train import models as M
    from oracle import abi_reference as R
    R.install(_Patch(), _lib, M)
    E._require_cuda = lambda t, what: None

    c = cfg(propagation_depth=1, hidden_dim=16, warmup_steps=4)
    packed = make_packed_graphs(24, seed=3)
    _, _, m = make_pair(c, packed.in_degree_histogram(), seed=rank, dtype64=False)
    engine = E.TrainEngine.from_config(m, c)
    store = CompactGraphStore(packed, device="cpu")
    seen = [ids for ids in store.epoch(4, generator=torch.Generator().manual_seed(5), rank=rank, world=world)]
    mean_loss = engine.fit_epoch(store, 4, generator=torch.Generator().manual_seed(5), rank=rank)
    flat = torch.cat([p.data.reshape(-1) for p in m.parameters()])
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    mine = torch.cat(seen)
    allids = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allids, mine)
    if rank == 0:
        ok = (torch.equal(gathered[0], gathered[1]) and sorted(torch.cat(allids).tolist()) == list(range(24))
              and int(engine.step_count) == 3 and mean_loss == mean_loss)
        with open(result_path, "w") as f:
            f.write(f"{int(ok)} steps={int(engine.step_count)} loss={mean_loss}")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_store_epoch(tmp_path):
    result = tmp_path / "result.txt"
    mp.spawn(_store_worker, args=(2, _free_port(), str(result)), nprocs=2, join=True)
    text = result.read_text()
    assert text.startswith("1"), text
