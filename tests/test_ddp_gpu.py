"""Data-parallel training step on >= 2 real GPUs (NCCL): identical parameters on every rank after the flat
gradient all-reduce + fused AdamW, equal to a two-replica reference computed with the oracle on the CPU."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.helpers import cfg, make_pair

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, result_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from gnnepcsaft_b200.data.synth import batch_from_packed, make_packed_graphs
    from gnnepcsaft_b200.train.engine import TrainEngine
    from oracle import pna_oracle as O

    c = cfg(propagation_depth=3, warmup_steps=4)
    packed = make_packed_graphs(64, seed=9)
    deg = packed.in_degree_histogram()
    o32, _, m = make_pair(c, deg, seed=rank, dtype64=False)       # rank-dependent init: engine must broadcast rank 0's
    m.to(dev)
    engine = TrainEngine.from_config(m, c)
    halves = [torch.arange(0, 32), torch.arange(32, 64)]
    losses = [float(engine.step(batch_from_packed(packed, halves[rank]).to(dev))) for _ in range(2)]
    flat = torch.cat([p.data.reshape(-1) for p in m.parameters()])
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    if rank == 0:
        same = bool(torch.equal(gathered[0], gathered[1]))
        ref0, _, _ = make_pair(c, deg, seed=0, dtype64=False)
        ref1 = O.create_model(c, deg)
        ref1.load_state_dict(ref0.state_dict())
        ref0.train(), ref1.train()
        opt = torch.optim.AdamW(ref0.parameters(), lr=c.learning_rate, weight_decay=c.weight_decay, amsgrad=True, eps=1e-5)
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
        rel = float((gathered[0].cpu() - end).norm() / (end - start).norm())
        ok = same and rel < 5e-2 and all(abs(a - b) <= 1e-4 * abs(b) + 1e-6 for a, b in zip(losses, ref_losses))
        with open(result_path, "w") as f:
            f.write(f"{int(ok)} same={same} rel={rel} losses={losses} ref={ref_losses}")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_data_parallel_step(tmp_path):
    result = tmp_path / "result.txt"
    mp.spawn(_worker, args=(2, _free_port(), str(result)), nprocs=2, join=True)
    text = result.read_text()
    assert text.startswith("1"), text


def _store_worker(rank, world, port, result_path):
    """Store-resident training with one captured CUDA graph per shape bucket on every rank: the ranks draw different
    batches (different buckets, captures at different steps), the NCCL all-reduce is part of the captured graph."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from gnnepcsaft_b200.data.compact import CompactGraphStore
    from gnnepcsaft_b200.data.synth import make_packed_graphs
    from gnnepcsaft_b200.train.engine import TrainEngine

    c = cfg(propagation_depth=2, warmup_steps=4)
    packed = make_packed_graphs(3000, seed=11)
    _, _, m = make_pair(c, packed.in_degree_histogram(), seed=rank, dtype64=False)
    m.to(dev)
    engine = TrainEngine.from_config(m, c)
    store = CompactGraphStore(packed, device=dev)
    gen = torch.Generator().manual_seed(100 + rank)
    losses = []
    for it in range(10):
        nb = 256 if (it + rank) % 3 else 512            # different bucket sequences on the two ranks
        ids = torch.randperm(3000, generator=gen)[:nb]
        losses.append(engine.step_from_store(store, ids).clone())
    torch.cuda.synchronize()
    flat = torch.cat([p.data.reshape(-1) for p in m.parameters()])
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    ls = torch.stack(losses).reshape(-1)
    gl = [torch.zeros_like(ls) for _ in range(world)]
    dist.all_gather(gl, ls)
    if rank == 0:
        same = bool(torch.equal(gathered[0], gathered[1]))
        same_loss = bool(torch.equal(gl[0], gl[1]))              # the loss is the rank mean on every rank
        finite = bool(torch.isfinite(ls).all())
        ok = same and same_loss and finite and engine.graph_replays >= 5
        with open(result_path, "w") as f:
            f.write(f"{int(ok)} same_params={same} same_loss={same_loss} finite={finite} replays={engine.graph_replays} "
                    f"buckets={len(engine._buckets)} losses={[round(float(v), 5) for v in ls]}")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_graphed_store_training(tmp_path):
    result = tmp_path / "result.txt"
    mp.spawn(_store_worker, args=(2, _free_port(), str(result)), nprocs=2, join=True)
    text = result.read_text()
    print(text)
    assert text.startswith("1"), text
