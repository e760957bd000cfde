"""Every CUDA entry point of the C ABI against its CPU restatement (oracle/abi_reference.py) on
identical inputs.  Integer / index outputs must be bit-exact; fp32 outputs within 1e-5 (relative
to the tensor's scale) unless stated."""
import math

import pytest
import torch

from gnnepcsaft_b200 import _lib
from gnnepcsaft_b200.data.synth import batch_from_packed, make_packed_graphs
from oracle import abi_reference as R

pytestmark = pytest.mark.gpu
DEV = "cuda"


def run_both(name, *args):
    """Run reference on CPU clones and the CUDA entry point on device clones; returns both arg lists."""
    cpu = [a.clone() if isinstance(a, torch.Tensor) else a for a in args]
    gpu = [a.to(DEV) if isinstance(a, torch.Tensor) else a for a in args]
    getattr(R, name)(*cpu)
    _lib.call(name, *gpu, _lib.stream_ptr())
    torch.cuda.synchronize()
    return cpu, gpu


def close(a_gpu, a_cpu, tol=1e-5, what=""):
    a_gpu = a_gpu.cpu().double()
    a_cpu = a_cpu.double()
    scale = max(float(a_cpu.abs().max()), 1e-30)
    err = float((a_gpu - a_cpu).abs().max()) / scale
    assert err <= tol, f"{what}: max-normalised error {err:.3e} > {tol}"


def random_graph(n, e, seed):
    g = torch.Generator().manual_seed(seed)
    ei = torch.randint(0, n, (2, e), generator=g)
    ea = torch.stack([torch.randint(0, d, (e,), generator=g) for d in (5, 6, 2)], 1)
    return ei, ea


def build_csr_cpu(ei, ea, n, self_loops):
    e = ei.shape[1]
    ep = e + (n if self_loops else 0)
    i32 = torch.int32
    out = dict(rowptr=torch.zeros(n + 1, dtype=i32), col_src=torch.zeros(max(ep, 1), dtype=i32),
               col_code=torch.zeros(max(ep, 1), dtype=torch.uint8), col_eid=torch.zeros(max(ep, 1), dtype=i32),
               srcptr=torch.zeros(n + 1, dtype=i32), src_slot=torch.zeros(max(ep, 1), dtype=i32))
    R.gnn_build_csr(ei, ea, n, e, self_loops, out["rowptr"], out["col_src"], out["col_code"], out["col_eid"],
                    out["srcptr"], out["src_slot"], None, None, 0)
    return out, ep


@pytest.mark.parametrize("n,e,loops", [(1, 0, True), (1, 0, False), (7, 0, True), (50, 160, True), (50, 160, False),
                                       (5000, 14000, True), (9000, 3, False)])
def test_build_csr_bit_exact(n, e, loops):
    ei, ea = random_graph(n, e, seed=n + e)
    ref, ep = build_csr_cpu(ei, ea, n, 1 if loops else 0)
    i32 = dict(dtype=torch.int32, device=DEV)
    rowptr, srcptr = torch.empty(n + 1, **i32), torch.empty(n + 1, **i32)
    col_src, col_eid, src_slot = (torch.empty(max(ep, 1), **i32) for _ in range(3))
    col_code = torch.empty(max(ep, 1), dtype=torch.uint8, device=DEV)
    err = torch.zeros(1, **i32)
    nbytes = _lib.query("gnn_build_csr_workspace_bytes", n, e)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=DEV)
    _lib.call("gnn_build_csr", ei.to(DEV), ea.to(DEV), n, e, 1 if loops else 0, rowptr, col_src, col_code, col_eid,
              srcptr, src_slot, err, ws, nbytes, _lib.stream_ptr())
    torch.cuda.synchronize()
    assert int(err.item()) == 0
    assert torch.equal(rowptr.cpu(), ref["rowptr"]) and torch.equal(srcptr.cpu(), ref["srcptr"])
    for k, t in (("col_src", col_src), ("col_code", col_code), ("col_eid", col_eid), ("src_slot", src_slot)):
        assert torch.equal(t[:ep].cpu(), ref[k][:ep]), k
    # properties: slots of a node are in PyG edge order and the self-loop is last
    if loops and ep:
        rp = rowptr.cpu()
        last = col_eid.cpu()[(rp[1:] - 1).long()]
        assert torch.equal(last, torch.arange(n, dtype=torch.int32) + e)


def test_build_csr_flags_out_of_range():
    ei = torch.tensor([[0, 1, 9], [1, 0, 1]])
    ea = torch.zeros(3, 3, dtype=torch.int64)
    n, e = 3, 3
    i32 = dict(dtype=torch.int32, device=DEV)
    bufs = [torch.zeros(8, **i32) for _ in range(5)]
    code = torch.zeros(8, dtype=torch.uint8, device=DEV)
    err = torch.zeros(1, **i32)
    nbytes = _lib.query("gnn_build_csr_workspace_bytes", n, e)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=DEV)
    _lib.call("gnn_build_csr", ei.to(DEV), ea.to(DEV), n, e, 1, bufs[0], bufs[1], code, bufs[2], bufs[3], bufs[4], err,
              ws, nbytes, _lib.stream_ptr())
    assert int(err.item()) != 0


def test_graph_ptr_with_empty_graphs():
    batch = torch.tensor([0, 0, 2, 2, 2, 5])
    n, b = 6, 8
    ng, gp = torch.zeros(n, dtype=torch.int32), torch.zeros(b + 1, dtype=torch.int32)
    cpu, gpu = run_both("gnn_graph_ptr_from_batch", batch, n, b, ng, gp, None)
    assert torch.equal(gpu[3].cpu(), cpu[3]) and torch.equal(gpu[4].cpu(), cpu[4])
    assert cpu[4].tolist() == [0, 2, 2, 5, 5, 5, 6, 6, 6]
    # malformed batch vectors are reported (GNN_INPUT_ERR_BATCH_RANGE = 4, _UNSORTED = 8), never silently accepted
    for bad, code in ((torch.tensor([0, 0, 2, 1, 2, 5]), 8), (torch.tensor([0, 0, 2, 2, 9, 9]), 4)):
        err = torch.zeros(1, dtype=torch.int32)
        cpu, gpu = run_both("gnn_graph_ptr_from_batch", bad, n, b, ng, gp, err)
        assert int(gpu[-1]) == code and int(cpu[-1]) == code
    sp = torch.zeros(2, dtype=torch.int32)
    cpu, gpu = run_both("gnn_graph_ptr_single", 17, sp)
    assert gpu[-1].tolist() == [0, 17] == cpu[-1].tolist()


@pytest.mark.parametrize("skewed", [False, True])
def test_device_collate_bit_exact(skewed):
    packed = make_packed_graphs(300, seed=11, skewed=skewed)
    g = torch.Generator().manual_seed(0)
    ids = torch.randperm(300, generator=g)[:129]
    ref = batch_from_packed(packed, ids)
    from gnnepcsaft_b200.data.store import GraphStore
    store = GraphStore(packed, device=DEV)
    out = store.batch(ids.to(DEV))
    for k in ("x", "edge_index", "edge_attr", "batch", "ptr"):
        assert torch.equal(getattr(out, k).cpu(), getattr(ref, k)), k
    assert torch.equal(out.para.cpu(), ref.para)


def test_atom_embed_fwd_bwd():
    packed = make_packed_graphs(40, seed=2)
    x = packed.x
    n, H = x.shape[0], 64
    rows = 174
    tables = torch.randn(rows, H)
    off = torch.tensor([0, 119, 124, 136, 148, 158, 164, 170, 172], dtype=torch.int32)
    out = torch.zeros(n, H)
    err = torch.zeros(1, dtype=torch.int32)
    cpu, gpu = run_both("gnn_atom_embed_fwd", x, n, tables, off, rows, H, out, err)
    close(gpu[-2], cpu[-2], 1e-6, "atom embedding")
    assert int(gpu[-1]) == 0 and int(cpu[-1]) == 0
    if n > 0:      # an out-of-range feature is clamped AND reported (the reference's embedding lookup raises)
        xb = x.clone()
        xb[n // 2, 0] = 119
        cpu, gpu = run_both("gnn_atom_embed_fwd", xb, n, tables, off, rows, H, out, err)
        assert int(gpu[-1]) == 16 and int(cpu[-1]) == 16
        close(gpu[-2], cpu[-2], 1e-6, "atom embedding (clamped)")
    dout = torch.randn(n, H)
    dt = torch.zeros(rows, H)
    cpu, gpu = run_both("gnn_atom_embed_bwd", x, n, dout, off, rows, H, dt)
    close(gpu[-1], cpu[-1], 1e-5, "atom table gradient")


@pytest.mark.parametrize("H", [32, 64, 128, 256, 48])
@pytest.mark.parametrize("graphs", [3, 700])
def test_atom_embed_bwd_private_copies(H, graphs):
    """H = 64 accumulates the table gradient in private shared-memory copies fed by a cp.async ring, the other widths
    take the atomic kernel: int64 and uint8 features, few and many nodes (every CTA, every ring slot and partial
    trips in use), accumulation onto a non-zero gradient."""
    packed = make_packed_graphs(graphs, seed=5)
    x = packed.x
    n, rows = x.shape[0], 174
    off = torch.tensor([0, 119, 124, 136, 148, 158, 164, 170, 172], dtype=torch.int32)
    dout = torch.randn(n, H)
    dt = torch.randn(rows, H)
    cpu, gpu = run_both("gnn_atom_embed_bwd", x, n, dout, off, rows, H, dt)
    close(gpu[-1], cpu[-1], 2e-5, "atom table gradient (int64 features)")
    cpu8, gpu8 = run_both("gnn_atom_embed_u8_bwd", x.to(torch.uint8), n, dout, off, rows, H, dt)
    close(gpu8[-1], cpu[-1], 2e-5, "atom table gradient (uint8 features)")


@pytest.mark.parametrize("H,L", [(64, 3), (16, 1)])
def test_bond_table_fwd_bwd(H, L):
    F, T = H, 2
    # fake flat arena: per layer [wee F*H | bee F | wpre_t0 F*3F | bpre F | wpre_t1 F*3F | bpre F]
    per_tower = F * 3 * F + F
    layer = F * H + F + T * per_tower
    arena = torch.randn(L * layer) * 0.2
    bond_tables = torch.randn(13, H)
    enc, tab = torch.zeros(L, 60, F), torch.zeros(L, 60, T * F)

    def run(fn, *args):
        # the kernels address layers through pointer strides into one arena: pass views of it
        cpu_arena, gpu_arena = arena.clone(), arena.to(DEV)
        res = []
        for base, call in ((cpu_arena, lambda *a: getattr(R, fn)(*a)),
                           (gpu_arena, lambda *a: _lib.call(fn, *a, _lib.stream_ptr()))):
            res.append(args_for(base, *args))
            call(*res[-1])
        torch.cuda.synchronize()
        return res

    def args_for(base, *args):
        dev = base.device
        return [a(base) if callable(a) else (a.to(dev).clone() if isinstance(a, torch.Tensor) else a) for a in args]

    wee = lambda b: b[0:F * H]
    bee = lambda b: b[F * H:F * H + F]
    wpre = lambda b: b[F * H + F:]
    cpu, gpu = run("gnn_bond_table_fwd", bond_tables, wee, bee, wpre, per_tower, layer, L, H, T, enc, tab)
    close(gpu[-2], cpu[-2], 1e-5, "enc")
    close(gpu[-1], cpu[-1], 1e-5, "tab")
    enc_ref = cpu[-2]
    dtab = torch.randn(L, 60, T * F)
    denc = torch.zeros(L, 60, F)
    dbond = torch.zeros(13, H)
    garena_cpu, garena_gpu = torch.zeros_like(arena), torch.zeros_like(arena).to(DEV)
    for base, garena, call in ((arena.clone(), garena_cpu, lambda *a: R.gnn_bond_table_bwd(*a)),
                               (arena.to(DEV), garena_gpu,
                                lambda *a: _lib.call("gnn_bond_table_bwd", *a, _lib.stream_ptr()))):
        dev = base.device
        db = dbond.to(dev).clone()
        call(dtab.to(dev), bond_tables.to(dev), wee(base), wpre(base), per_tower, layer, L, enc_ref.to(dev), H, T,
             denc.to(dev).clone(), db, wee(garena), bee(garena), wpre(garena))
        if dev.type == "cpu":
            db_cpu = db
        else:
            db_gpu = db
    torch.cuda.synchronize()
    close(garena_gpu, garena_cpu, 1e-5, "weight grads")
    close(db_gpu, db_cpu, 1e-5, "bond embedding grads")


def _agg_inputs(n, e, T, F, seed, loops=True, quantise=False):
    ei, ea = random_graph(n, e, seed)
    ref, ep = build_csr_cpu(ei, ea, n, 1 if loops else 0)
    C = T * F
    g = torch.Generator().manual_seed(seed)
    P = torch.randn(n, 2 * C, generator=g)
    tab = torch.randn(60, C, generator=g) * 0.3
    if quantise:   # exact ties and zero variance
        P = (P * 2).round() / 2
        tab = (tab * 2).round() / 2
    return ref, ep, P, tab


@pytest.mark.parametrize("n,e,T,F,loops,quant", [(64, 150, 2, 64, True, False), (64, 150, 2, 64, False, True),
                                                 (33, 40, 2, 16, True, True), (20, 300, 2, 256, True, False),
                                                 (200, 500, 2, 40, False, False)])
def test_aggregate_fwd_bwd(n, e, T, F, loops, quant):
    ref, ep, P, tab = _agg_inputs(n, e, T, F, seed=n + F, loops=loops, quantise=quant)
    C = T * F
    A = torch.zeros(n, T * 4 * F)
    cpu, gpu = run_both("gnn_pna_aggregate_fwd", P, tab, ref["rowptr"], ref["col_src"], ref["col_code"], n, T, F, A)
    A_cpu, A_gpu = cpu[-1], gpu[-1].cpu()
    a4c, a4g = A_cpu.view(n, T, 4, F), A_gpu.view(n, T, 4, F)
    close(a4g[:, :, :3], a4c[:, :, :3], 1e-5, "mean/min/max")
    # std: elementwise, except threshold-adjacent elements (var within 1e-6 of the 1e-5 floor) may flip to 0
    sd_c, sd_g = a4c[:, :, 3], a4g[:, :, 3]
    near = (sd_c.double() ** 2 - 1e-5).abs() < 2e-6
    bad = ((sd_g - sd_c).abs() > 1e-5 * max(float(sd_c.abs().max()), 1.0)) & ~near & ~((sd_c == 0) & (sd_g ** 2 < 1.2e-5))
    assert not bad.any(), f"{int(bad.sum())} std elements differ away from the threshold"
    # empty segments are exactly zero
    deg = ref["rowptr"][1:] - ref["rowptr"][:-1]
    assert torch.all(A_gpu[deg == 0] == 0)

    gA = torch.randn(n, T * 4 * F)
    dP, dq, dtab = torch.zeros(n, 2 * C), torch.zeros(max(ep, 1), C), torch.zeros(60, C)
    # feed the CUDA forward's A to both sides so that only the backward arithmetic is compared
    nb = _lib.query("gnn_pna_aggregate_bwd_workspace_bytes", T, F)
    zws = torch.zeros(nb, dtype=torch.uint8)
    cpu, gpu = run_both("gnn_pna_aggregate_bwd", gA, P, tab, A_gpu, ref["rowptr"], ref["col_src"], ref["col_code"], n,
                        T, F, dP, dq, dtab, zws, nb)
    close(gpu[-5][:, :C], cpu[-5][:, :C], 1e-5, "dPdst")
    close(gpu[-4][:ep], cpu[-4][:ep], 2e-5, "dq")
    close(gpu[-3], cpu[-3], 1e-4, "dtab")
    assert int(gpu[-2].count_nonzero()) == 0, "workspace must be left all-zero"
    dq_gpu = gpu[-4].cpu()
    dP2 = torch.zeros(n, 2 * C)
    cpu, gpu = run_both("gnn_scatter_src_bwd", dq_gpu, ref["srcptr"], ref["src_slot"], n, C, dP2)
    close(gpu[-1][:, C:], cpu[-1][:, C:], 1e-5, "dPsrc")
    dPsrc_two_pass = gpu[-1][:, C:].cpu()
    # dq = NULL: source gradients accumulate atomically inside the kernel (no dq buffer, no second pass)
    dP3, dtab3 = torch.full((n, 2 * C), 7.0), torch.zeros(60, C)
    cpu, gpu = run_both("gnn_pna_aggregate_bwd", gA, P, tab, A_gpu, ref["rowptr"], ref["col_src"], ref["col_code"], n,
                        T, F, dP3, None, dtab3, zws, nb)
    close(gpu[-5][:, :C], cpu[-5][:, :C], 1e-5, "dPdst (atomic variant)")
    close(gpu[-5][:, C:], cpu[-5][:, C:], 2e-5, "dPsrc (atomic variant)")
    close(gpu[-5][:, C:], dPsrc_two_pass, 2e-5, "dPsrc atomic vs two-pass")
    close(gpu[-3], cpu[-3], 1e-4, "dtab (atomic variant)")
    assert int(gpu[-2].count_nonzero()) == 0


def test_general_message_path():
    n, e, T, F = 50, 140, 2, 16
    ref, ep, P, tab = _agg_inputs(n, e, T, F, seed=9, loops=False, quantise=True)
    C = T * F
    h = torch.zeros(ep, C)
    cpu, gpu = run_both("gnn_edge_hidden_fwd", P, tab, ref["rowptr"], ref["col_src"], ref["col_code"], n, C, 1, h)
    close(gpu[-1], cpu[-1], 1e-6, "edge hidden")
    m = cpu[-1]
    A = torch.zeros(n, T * 4 * F)
    cpu, gpu = run_both("gnn_pna_aggregate_msg_fwd", m, ref["rowptr"], n, T, F, A)
    close(gpu[-1], cpu[-1], 2e-5, "aggregate(msg)")
    A_gpu = gpu[-1].cpu()
    gA, dm = torch.randn(n, T * 4 * F), torch.zeros(ep, C)
    cpu, gpu = run_both("gnn_pna_aggregate_msg_bwd", gA, m, A_gpu, ref["rowptr"], n, T, F, dm)
    close(gpu[-1], cpu[-1], 2e-5, "dm")
    dh = torch.randn(ep, C)
    dP, dtab = torch.zeros(n, 2 * C), torch.zeros(60, C)
    cpu, gpu = run_both("gnn_edge_hidden_bwd", dh, ref["rowptr"], ref["col_code"], n, C, dP, dtab)
    close(gpu[-2][:, :C], cpu[-2][:, :C], 1e-5, "dPdst(general)")
    close(gpu[-1], cpu[-1], 1e-5, "dtab(general)")


@pytest.mark.parametrize("M,N,K,ldx,ldw,ldy", [(1000, 64, 64, 64, 192, 256), (1, 5, 16, 16, 16, 5),
                                               (333, 32, 832, 832, 832, 64), (129, 256, 64, 70, 64, 256),
                                               (4096, 16, 32, 32, 32, 16), (77, 10, 20, 20, 20, 10)])
def test_linear_fwd_dgrad_wgrad(M, N, K, ldx, ldw, ldy):
    g = torch.Generator().manual_seed(M + N + K)
    X, W, b = torch.randn(M, ldx, generator=g), torch.randn(N, ldw, generator=g) * 0.1, torch.randn(N, generator=g)
    for act in (0, 1):
        Y = torch.full((M, ldy), 7.0)
        cpu, gpu = run_both("gnn_linear_fwd", X, ldx, W, ldw, b, M, N, K, act, Y, ldy)
        close(gpu[-2], cpu[-2], 1e-5, "linear fwd")          # untouched columns stay 7.0 on both sides
    dY = torch.randn(M, ldy, generator=g)
    for acc in (0, 1):
        dX = torch.randn(M, ldx, generator=g)
        cpu, gpu = run_both("gnn_linear_dgrad", dY, ldy, W, ldw, M, N, K, acc, dX, ldx)
        close(gpu[-2], cpu[-2], 1e-5, "linear dgrad")
    dW, db = torch.randn(N, ldw, generator=g), torch.randn(N, generator=g)
    cpu, gpu = run_both("gnn_linear_wgrad", dY, ldy, X, ldx, M, N, K, dW, ldw, db)
    close(gpu[-3], cpu[-3], 2e-5, "linear wgrad")
    close(gpu[-1], cpu[-1], 2e-5, "bias grad")
    cpu, gpu = run_both("gnn_linear_wgrad", dY, ldy, X, ldx, M, N, K, dW, ldw, None)
    close(gpu[-3], cpu[-3], 2e-5, "linear wgrad (no bias)")


@pytest.mark.parametrize("n,F", [(300, 64), (70, 16), (1, 8), (130, 128)])
def test_pna_post_fwd_dgrad_wgrad(n, F):
    T, Fo = 2, F // 2
    g = torch.Generator().manual_seed(n + F)
    x, A = torch.randn(n, F, generator=g), torch.randn(n, T * 4 * F, generator=g)
    rowptr = torch.zeros(n + 1, dtype=torch.int32)
    rowptr[1:] = torch.randint(0, 6, (n,), generator=g).cumsum(0).to(torch.int32)
    delta = torch.tensor([0.93])
    scal = torch.zeros(n, 2)
    cpu, gpu = run_both("gnn_pna_scalers", rowptr, delta, n, scal)
    close(gpu[-1], cpu[-1], 1e-6, "scalers")
    scal = cpu[-1]
    # two towers stored with a gap (as in the parameter arena: weight, bias, weight, bias)
    wts = Fo * 13 * F + Fo + 4
    wbuf = torch.randn(T * wts, generator=g) * 0.1
    bts = wts
    boff = Fo * 13 * F

    def views(base):
        return base[0:], base[boff:]

    for act in (0, 1):
        O = torch.zeros(n, T * Fo)
        res = []
        for dev in ("cpu", DEV):
            base = wbuf.to(dev).clone()
            w, b = views(base)
            args = [x.to(dev), A.to(dev), scal.to(dev), w, wts, b, bts, n, F, Fo, T, act, O.to(dev).clone()]
            if dev == "cpu":
                R.gnn_pna_post_fwd(*args)
            else:
                _lib.call("gnn_pna_post_fwd", *args, _lib.stream_ptr())
            res.append(args[-1])
        close(res[1], res[0], 1e-5, "post fwd")
    dO = torch.randn(n, T * Fo, generator=g)
    res = []
    for dev in ("cpu", DEV):
        base = wbuf.to(dev).clone()
        w, _ = views(base)
        dx, gA = torch.ones(n, F).to(dev), torch.zeros(n, T * 4 * F).to(dev)
        args = [dO.to(dev), scal.to(dev), w, wts, n, F, Fo, T, dx, gA]
        R.gnn_pna_post_dgrad(*args) if dev == "cpu" else _lib.call("gnn_pna_post_dgrad", *args, _lib.stream_ptr())
        res.append((dx, gA))
    close(res[1][0], res[0][0], 1e-5, "post dgrad dx")
    close(res[1][1], res[0][1], 1e-5, "post dgrad gA")
    res = []
    for dev in ("cpu", DEV):
        gbase = torch.zeros(T * wts).to(dev)
        gw, gb = views(gbase)
        args = [dO.to(dev), x.to(dev), A.to(dev), scal.to(dev), n, F, Fo, T, gw, wts, gb, bts]
        R.gnn_pna_post_wgrad(*args) if dev == "cpu" else _lib.call("gnn_pna_post_wgrad", *args, _lib.stream_ptr())
        res.append(gbase)
    close(res[1], res[0], 2e-5, "post wgrad")


@pytest.mark.parametrize("rows,C", [(5000, 64), (37, 16), (2, 5), (4096, 32)])
def test_batchnorm_fwd_bwd(rows, C):
    g = torch.Generator().manual_seed(rows + C)
    Y = torch.randn(rows, C, generator=g) * 2 + 0.5
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.2
    skip = torch.randn(rows, C, generator=g)
    nbytes = _lib.query("gnn_bn_workspace_bytes", rows, C)
    for training in (1, 0):
        rm, rv = torch.randn(C, generator=g) * 0.3, torch.rand(C, generator=g) + 0.5
        nbt = torch.tensor([3], dtype=torch.int64)
        stats = torch.zeros(2 * C)
        ws = torch.zeros(nbytes, dtype=torch.uint8)
        cpu, gpu = run_both("gnn_bn_stats", Y, rows, C, training, 1e-5, 0.1, rm, rv, nbt, stats, None, ws, nbytes)
        for i, what in ((6, "running_mean"), (7, "running_var"), (9, "stats")):
            close(gpu[i], cpu[i], 1e-5, what)
        assert int(gpu[8].item()) == int(cpu[8].item()) == (4 if training else 3)
        stats = cpu[9]
        kernel_running_mean, kernel_running_var = gpu[6].cpu(), gpu[7].cpu()
        # cross-check against torch BatchNorm1d itself
        bn = torch.nn.BatchNorm1d(C)
        bn.weight.data, bn.bias.data = gamma.clone(), beta.clone()
        bn.running_mean.copy_(rm), bn.running_var.copy_(rv)
        bn.train(bool(training))
        Yr = Y.clone().requires_grad_(True)
        ref_out = torch.relu(bn(Yr)) + skip
        out = torch.zeros(rows, C)
        cpu, gpu = run_both("gnn_bn_act_fwd", Y, stats, gamma, beta, skip, rows, C, 1, out)
        close(gpu[-1], ref_out.detach(), 1e-5, "bn+relu+skip vs torch")
        if training:   # running statistics follow torch (momentum 0.1, unbiased variance)
            close(kernel_running_mean, bn.running_mean, 1e-5, "running_mean vs torch")
            close(kernel_running_var, bn.running_var, 1e-5, "running_var vs torch")
        dOut = torch.randn(rows, C, generator=g)
        ref_out.backward(dOut)
        dY, dga, dbe = torch.zeros(rows, C), torch.zeros(C), torch.zeros(C)
        cpu, gpu = run_both("gnn_bn_act_bwd", dOut, Y, stats, gamma, beta, rows, C, 1, training, dY, dga, dbe, None, ws,
                            nbytes)
        # normalise by the scale of the incoming gradient: with very few rows the BN backward is
        # dominated by cancellation and the result is orders of magnitude below its inputs
        in_scale = float((dOut.abs().max() * (gamma * stats[C:]).abs().max()))
        err = float((gpu[9].cpu() - Yr.grad).abs().max()) / in_scale
        assert err <= 2e-5, f"bn backward vs torch autograd: {err:.2e}"
        close(gpu[10], bn.weight.grad, 2e-5, "dgamma")
        close(gpu[11], bn.bias.grad, 2e-5, "dbeta")


@pytest.mark.parametrize("rows,valid,C", [(5000, 4321, 64), (40, 37, 16), (300, 300, 32), (9, 2, 8)])
def test_batchnorm_padded_rows(rows, valid, C):
    """valid_rows: a padded batch must give exactly the statistics / gradients of its first `valid` rows (whatever the
    padding rows contain) and zero gradient rows for the padding."""
    g = torch.Generator().manual_seed(rows + C)
    Y = torch.randn(rows, C, generator=g) * 2 + 0.5
    Y[valid:] = 1e6                                  # garbage in the padding must not leak into anything
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.2
    nbytes = _lib.query("gnn_bn_workspace_bytes", rows, C)
    ws = torch.zeros(nbytes, dtype=torch.uint8, device=DEV)
    vr = torch.tensor([valid], dtype=torch.int32, device=DEV)
    dOut = torch.randn(rows, C, generator=g)
    dOut[valid:] = -1e6
    res = {}
    for name, r, v in (("padded", rows, vr), ("plain", valid, None)):
        rm, rv = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)
        nbt = torch.zeros(1, dtype=torch.int64, device=DEV)
        stats = torch.zeros(2 * C, device=DEV)
        Yd, dOd = Y[:r].contiguous().to(DEV), dOut[:r].contiguous().to(DEV)
        _lib.call("gnn_bn_stats", Yd, r, C, 1, 1e-5, 0.1, rm, rv, nbt, stats, v, ws, nbytes, _lib.stream_ptr())
        dY = torch.full((r, C), 3.0, device=DEV)
        dga, dbe = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
        _lib.call("gnn_bn_act_bwd", dOd, Yd, stats, gamma.to(DEV), beta.to(DEV), r, C, 1, 1, dY, dga, dbe, v, ws, nbytes,
                  _lib.stream_ptr())
        torch.cuda.synchronize()
        res[name] = [t.cpu() for t in (stats, rm, rv, dY, dga, dbe)]
    for a, b, what in zip(res["padded"], res["plain"], ("stats", "running_mean", "running_var", "dY", "dgamma", "dbeta")):
        if what == "dY":
            assert torch.all(a[valid:] == 0), "gradient of the padding rows must be zero"
            a = a[:valid]
        close(a, b, 1e-6, what)
    cpu, gpu = run_both("gnn_bn_stats", Y, rows, C, 1, 1e-5, 0.1, torch.zeros(C), torch.ones(C),
                        torch.zeros(1, dtype=torch.int64), torch.zeros(2 * C), vr.cpu(), ws.cpu(), nbytes)
    close(gpu[9], cpu[9], 1e-5, "padded stats vs reference")


def test_pool_mape():
    g = torch.Generator().manual_seed(0)
    counts = torch.tensor([3, 0, 5, 1, 40])
    n, b, C = int(counts.sum()), 5, 64
    gp = torch.zeros(b + 1, dtype=torch.int32)
    gp[1:] = counts.cumsum(0).to(torch.int32)
    ng = torch.repeat_interleave(torch.arange(b), counts).to(torch.int32)
    x = torch.randn(n, C, generator=g)
    out = torch.zeros(b, C)
    cpu, gpu = run_both("gnn_segment_pool_fwd", x, gp, b, C, out)
    close(gpu[-1], cpu[-1], 1e-6, "pool")
    dout = torch.randn(b, C, generator=g)
    for acc in (0, 1):
        dx = torch.randn(n, C, generator=g)
        cpu, gpu = run_both("gnn_segment_pool_bwd", dout, ng, n, C, acc, dx)
        close(gpu[-1], cpu[-1], 1e-6, "pool bwd")
    pred, tgt = torch.randn(700, 5, generator=g), torch.rand(700, 5, generator=g) * 4 + 0.5
    tgt[0, 0] = 0.0                                       # exercises the 1.17e-6 clamp
    loss, dpred = torch.zeros(1), torch.zeros(700, 5)
    cpu, gpu = run_both("gnn_mape_fwd_bwd", pred, tgt, 3500, 0.5, loss, dpred, None, 5)
    close(gpu[4], cpu[4], 1e-6, "mape")
    close(gpu[5], cpu[5], 1e-6, "dmape")
    # padded batch: only the first `valid` rows carry a loss, the rest gets a zero gradient
    vr = torch.tensor([613], dtype=torch.int32)
    loss, dpred = torch.zeros(1), torch.full((700, 5), 7.0)
    cpu, gpu = run_both("gnn_mape_fwd_bwd", pred, tgt, 3500, 1.0, loss, dpred, vr, 5)
    close(gpu[4], cpu[4], 1e-6, "mape (valid rows)")
    close(gpu[5], cpu[5], 1e-6, "dmape (valid rows)")
    assert torch.all(gpu[5][613:] == 0)
    ref = ((pred[:613] - tgt[:613]).abs() / tgt[:613].abs().clamp(min=1.17e-6)).mean()
    assert abs(float(gpu[4]) - float(ref)) <= 1e-5 * float(ref)


def test_adamw_amsgrad_cosine_restarts_vs_torch():
    """Fused optimiser against torch.optim.AdamW(amsgrad) + CosineAnnealingWarmRestarts over 25 steps."""
    g = torch.Generator().manual_seed(0)
    n, T0 = 10007, 10
    p0 = torch.randn(n, generator=g)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([ref], lr=1e-3, weight_decay=1e-2, amsgrad=True, eps=1e-5)
    sched = torch.optim.lr_scheduler.CosineAnnealingWarmRestarts(opt, T0)
    p = p0.to(DEV)
    m, v, vmax = (torch.zeros(n, device=DEV) for _ in range(3))
    step = torch.zeros(1, dtype=torch.int64, device=DEV)
    for it in range(25):
        grad = torch.randn(n, generator=g) * (1 + it % 3)
        ref.grad = grad.clone()
        opt.step()
        sched.step()
        _lib.call("gnn_adamw_amsgrad_step", p, (grad * 4).to(DEV), m, v, vmax, n, step, 1e-3, 0.9, 0.999, 1e-5, 1e-2,
                  T0, 0.25, _lib.stream_ptr())
        _lib.call("gnn_step_increment", step, _lib.stream_ptr())
    torch.cuda.synchronize()
    assert int(step.item()) == 25
    close(p, ref.detach(), 2e-6, "parameters after 25 fused AdamW steps")


# ------------------------------------------------------------------ tcgen05 / TMEM dense stages
# precision 0 = TF32x3 (fp32-level: 1e-5), 2 = bf16 operands (stated
# tolerance 2e-2 of the tensor scale)
TC_TOL = {0: 1e-5, 2: 2e-2}


@pytest.mark.parametrize("precision", [0, 2])
@pytest.mark.parametrize("M,N,K,ldx,ldw,ldy", [(1000, 64, 64, 64, 192, 256), (1, 8, 16, 16, 16, 8),
                                               (333, 32, 832, 832, 832, 64), (129, 256, 64, 72, 64, 256),
                                               (4096, 16, 32, 32, 32, 16), (77, 12, 20, 20, 20, 12),
                                               (5000, 64, 100, 100, 100, 64), (20000, 64, 256, 256, 256, 64)])
def test_tc_linear_fwd_dgrad_wgrad(precision, M, N, K, ldx, ldw, ldy):
    tol = TC_TOL[precision]
    g = torch.Generator().manual_seed(M + N + K)
    X, W, b = torch.randn(M, ldx, generator=g), torch.randn(N, ldw, generator=g) * 0.1, torch.randn(N, generator=g)
    for act in (0, 1):
        Y = torch.full((M, ldy), 7.0)
        for ws in (None, torch.zeros(4 << 20, dtype=torch.uint8)):         # without / with the packed-weight path
            cpu, gpu = run_both("gnn_tc_linear_fwd", X, ldx, W, ldw, b, M, N, K, act, Y, ldy, precision, ws, 4 << 20)
            close(gpu[-5], cpu[-5], tol, "tc linear fwd")
    dY = torch.randn(M, ldy, generator=g)
    for acc in (0, 1):
        dX = torch.randn(M, ldx, generator=g)
        for ws in (None, torch.zeros(4 << 20, dtype=torch.uint8)):
            cpu, gpu = run_both("gnn_tc_linear_dgrad", dY, ldy, W, ldw, M, N, K, acc, dX, ldx, precision, ws, 4 << 20)
            close(gpu[-5], cpu[-5], tol, "tc linear dgrad")
    dW, db = torch.randn(N, ldw, generator=g), torch.randn(N, generator=g)
    cpu, gpu = run_both("gnn_tc_linear_wgrad", dY, ldy, X, ldx, M, N, K, dW, ldw, db, precision)
    close(gpu[-4], cpu[-4], 2 * tol, "tc linear wgrad")
    close(gpu[-2], cpu[-2], 2e-5 if precision < 2 else tol, "tc linear wgrad bias")
    cpu, gpu = run_both("gnn_tc_linear_wgrad", dY, ldy, X, ldx, M, N, K, dW, ldw, None, precision)
    close(gpu[-4], cpu[-4], 2 * tol, "tc linear wgrad (no bias)")


@pytest.mark.parametrize("precision", [0, 2])
@pytest.mark.parametrize("n,F", [(300, 64), (70, 16), (1, 8), (130, 128), (257, 256), (90, 40)])
def test_tc_pna_post_fwd_dgrad_wgrad(precision, n, F):
    tol = TC_TOL[precision]
    T, Fo = 2, F // 2
    g = torch.Generator().manual_seed(n + F)
    x, A = torch.randn(n, F, generator=g), torch.randn(n, T * 4 * F, generator=g)
    scal = torch.rand(n, 2, generator=g) + 0.5
    wts = Fo * 13 * F + Fo + 4
    wbuf = torch.randn(T * wts, generator=g) * 0.1
    bts, boff = wts, Fo * 13 * F
    for act in (0, 1):
        res = []
        for dev in ("cpu", DEV):
            base = wbuf.to(dev).clone()
            args = [x.to(dev), A.to(dev), scal.to(dev), base, wts, base[boff:], bts, n, F, Fo, T, act,
                    torch.zeros(n, T * Fo).to(dev), precision, torch.zeros(4 << 20, dtype=torch.uint8).to(dev), 4 << 20]
            R.gnn_tc_pna_post_fwd(*args) if dev == "cpu" else _lib.call("gnn_tc_pna_post_fwd", *args, _lib.stream_ptr())
            res.append(args[-4])
        close(res[1], res[0], tol, "tc post fwd")
    dO = torch.randn(n, T * Fo, generator=g)
    res = []
    for dev in ("cpu", DEV):
        base = wbuf.to(dev).clone()
        dx, gA = torch.ones(n, F).to(dev), torch.zeros(n, T * 4 * F).to(dev)
        args = [dO.to(dev), scal.to(dev), base, wts, n, F, Fo, T, dx, gA, precision,
                torch.zeros(4 << 20, dtype=torch.uint8).to(dev), 4 << 20]
        R.gnn_tc_pna_post_dgrad(*args) if dev == "cpu" else _lib.call("gnn_tc_pna_post_dgrad", *args, _lib.stream_ptr())
        res.append((dx, gA))
    close(res[1][0], res[0][0], tol, "tc post dgrad dx")
    close(res[1][1], res[0][1], tol, "tc post dgrad gA")
    gA2 = torch.zeros(n, T * 4 * F, device=DEV)                   # dx = NULL: only gA
    _lib.call("gnn_tc_pna_post_dgrad", dO.to(DEV), scal.to(DEV), wbuf.to(DEV), wts, n, F, Fo, T, None, gA2, precision,
              torch.zeros(4 << 20, dtype=torch.uint8, device=DEV), 4 << 20, _lib.stream_ptr())
    close(gA2, res[0][1], tol, "tc post dgrad gA (dx=NULL)")
    res = []
    for dev in ("cpu", DEV):
        gbase = torch.zeros(T * wts).to(dev)
        args = [dO.to(dev), x.to(dev), A.to(dev), scal.to(dev), n, F, Fo, T, gbase, wts, gbase[boff:], bts, precision]
        R.gnn_tc_pna_post_wgrad(*args) if dev == "cpu" else _lib.call("gnn_tc_pna_post_wgrad", *args, _lib.stream_ptr())
        res.append(gbase)
    close(res[1], res[0], 2 * tol, "tc post wgrad")


def test_tc_rejects_unaligned_operands():
    """The tensor-core producers use 128-bit accesses: shapes that are not 4-column aligned are refused
    (the host falls back to the SIMT kernels for them), never silently mis-read."""
    lib = _lib.load()
    X = torch.zeros(64, 70, device=DEV)
    W = torch.zeros(8, 70, device=DEV)
    Y = torch.zeros(64, 8, device=DEV)
    rc = lib.gnn_tc_linear_fwd(X.data_ptr(), 70, W.data_ptr(), 70, None, 64, 8, 70, 0, Y.data_ptr(), 8, 0, None, 0,
                               _lib.stream_ptr())
    assert rc == -3
    rc = lib.gnn_tc_linear_dgrad(Y.data_ptr(), 5, W.data_ptr(), 70, 64, 5, 70, 0, X.data_ptr(), 70, 0, None, 0,
                                 _lib.stream_ptr())
    assert rc == -3


@pytest.mark.parametrize("precision", [0, 2])
@pytest.mark.parametrize("n,F", [(1000, 64), (70, 16), (3, 8), (300, 128)])
def test_tc_fused_projection_stage(precision, n, F):
    """proj_fwd / dx_dgrad / proj_wgrad: the four projection Linears (and the x part of the post stage) in one GEMM each."""
    tol = TC_TOL[precision]
    T, Fo, C = 2, F // 2, 2 * F
    g = torch.Generator().manual_seed(n + F)
    x = torch.randn(n, F, generator=g)
    pre_ts = F * 3 * F + F + 4            # [weight F x 3F | bias F | pad] per tower
    pre = torch.randn(T * pre_ts, generator=g) * 0.1
    post_ts = Fo * 13 * F + Fo + 4
    post = torch.randn(T * post_ts, generator=g) * 0.1
    ws = lambda dev: torch.zeros(4 << 20, dtype=torch.uint8, device=dev)
    res = []
    for dev in ("cpu", DEV):
        b = pre.to(dev)
        args = [x.to(dev), b, pre_ts, b[F * 3 * F:], pre_ts, n, F, T, torch.zeros(n, 2 * C, device=dev), precision, ws(dev), 4 << 20]
        R.gnn_tc_pna_proj_fwd(*args) if dev == "cpu" else _lib.call("gnn_tc_pna_proj_fwd", *args, _lib.stream_ptr())
        res.append(args[8])
    close(res[1], res[0], tol, "proj fwd")
    dP, dO = torch.randn(n, 2 * C, generator=g), torch.randn(n, T * Fo, generator=g)
    for acc in (0, 1):
        res = []
        for dev in ("cpu", DEV):
            dx = torch.ones(n, F, device=dev) * 0.5
            args = [dP.to(dev), dO.to(dev), pre.to(dev), pre_ts, post.to(dev), post_ts, n, F, Fo, T, acc, dx, precision,
                    ws(dev), 4 << 20]
            R.gnn_tc_pna_dx_dgrad(*args) if dev == "cpu" else _lib.call("gnn_tc_pna_dx_dgrad", *args, _lib.stream_ptr())
            res.append(dx)
        close(res[1], res[0], tol, "fused dx dgrad")
    res = []
    for dev in ("cpu", DEV):
        gb = torch.zeros(T * pre_ts, device=dev)
        args = [dP.to(dev), x.to(dev), n, F, T, gb, pre_ts, gb[F * 3 * F:], pre_ts, precision]
        R.gnn_tc_pna_proj_wgrad(*args) if dev == "cpu" else _lib.call("gnn_tc_pna_proj_wgrad", *args, _lib.stream_ptr())
        res.append(gb)
    close(res[1], res[0], 2 * tol, "proj wgrad")


# ------------------------------------------------------------------ folded destination projection
@pytest.mark.parametrize("F,L", [(64, 3), (16, 1), (8, 2), (128, 2), (40, 1)])
def test_fold_fwd_bwd(F, L):
    T, Fo = 2, F // 2
    g = torch.Generator().manual_seed(F)
    pre_ts = F * 3 * F + F + 4
    post_ts = Fo * 13 * F + Fo + 4
    ls = T * (pre_ts + post_ts) + 8                 # layer l: [pre towers | post towers | pad]
    buf = torch.randn(L * ls, generator=g) * 0.1
    post = buf[T * pre_ts:]
    cpu, gpu = run_both("gnn_pna_fold_fwd", post, post_ts, buf, pre_ts, ls, L, F, Fo, T, torch.zeros(L, T, 3, Fo, F))
    close(gpu[-1], cpu[-1], 1e-5, "Vx")
    dVx = torch.randn(L, T, 3, Fo, F, generator=g)
    res = []
    for dev in ("cpu", DEV):
        gb, wb = torch.randn(L * ls, generator=torch.Generator().manual_seed(1)).to(dev), buf.to(dev)   # += semantics
        args = [dVx.to(dev), wb[T * pre_ts:], post_ts, wb, pre_ts, ls, L, F, Fo, T, gb[T * pre_ts:], gb]
        R.gnn_pna_fold_bwd(*args) if dev == "cpu" else _lib.call("gnn_pna_fold_bwd", *args, _lib.stream_ptr())
        res.append(gb)
    close(res[1], res[0], 1e-5, "dwpost / dwpre")


@pytest.mark.parametrize("n,e,T,F,quant", [(64, 150, 2, 64, False), (33, 40, 2, 16, True), (20, 300, 2, 256, False),
                                           (200, 500, 2, 40, False), (64, 150, 2, 64, True)])
def test_aggregate_src_fwd_bwd(n, e, T, F, quant):
    ref, ep, P, tab = _agg_inputs(n, e, T, F, seed=n + F, loops=True, quantise=quant)
    C = T * F
    Ps = P[:, C:].contiguous()
    A = torch.zeros(n, T * 4 * F)
    cpu, gpu = run_both("gnn_pna_aggregate_src_fwd", Ps, tab, ref["rowptr"], ref["col_src"], ref["col_code"], n, T, F, A)
    A_cpu, A_gpu = cpu[-1], gpu[-1].cpu()
    a4c, a4g = A_cpu.view(n, T, 4, F), A_gpu.view(n, T, 4, F)
    close(a4g[:, :, :3], a4c[:, :, :3], 1e-5, "mean/min/max of q")
    sd_c, sd_g = a4c[:, :, 3], a4g[:, :, 3]
    near = (sd_c.double() ** 2 - 1e-5).abs() < 2e-6
    bad = ((sd_g - sd_c).abs() > 1e-5 * max(float(sd_c.abs().max()), 1.0)) & ~near & ~((sd_c == 0) & (sd_g ** 2 < 1.2e-5))
    assert not bad.any()
    # the unfolded kernel on [Pdst | Psrc] must give the same aggregates shifted by Pdst (std unchanged)
    A_full = torch.zeros(n, T * 4 * F, device=DEV)
    _lib.call("gnn_pna_aggregate_fwd", P.to(DEV), tab.to(DEV), ref["rowptr"].to(DEV), ref["col_src"].to(DEV),
              ref["col_code"].to(DEV), n, T, F, A_full, _lib.stream_ptr())
    af = A_full.cpu().view(n, T, 4, F)
    pd = P[:, :C].view(n, T, 1, F)
    close(af[:, :, :3], a4g[:, :, :3] + pd, 1e-5, "unfolded == folded + Pdst")
    assert torch.equal(af[:, :, 3], a4g[:, :, 3])

    gA = torch.randn(n, T * 4 * F)
    nb = _lib.query("gnn_pna_aggregate_bwd_workspace_bytes", T, F)
    zws = torch.zeros(nb, dtype=torch.uint8)
    dPs, dq, dtab = torch.zeros(n, C), torch.zeros(max(ep, 1), C), torch.zeros(60, C)
    cpu, gpu = run_both("gnn_pna_aggregate_src_bwd", gA, Ps, tab, A_gpu, ref["rowptr"], ref["col_src"],
                        ref["col_code"], n, T, F, dPs, dq, dtab, zws, nb)
    close(gpu[-4][:ep], cpu[-4][:ep], 2e-5, "dq")
    close(gpu[-3], cpu[-3], 1e-4, "dtab")
    assert int(gpu[-2].count_nonzero()) == 0
    dq_gpu = gpu[-4].cpu()
    cpu, gpu = run_both("gnn_slot_sum_by_source", dq_gpu, ref["srcptr"], ref["src_slot"], n, C, torch.zeros(n, C), C)
    close(gpu[-2], cpu[-2], 1e-5, "dPsrc")
    two_pass = gpu[-2].cpu()
    dPs3, dtab3 = torch.full((n, C), 7.0), torch.zeros(60, C)
    cpu, gpu = run_both("gnn_pna_aggregate_src_bwd", gA, Ps, tab, A_gpu, ref["rowptr"], ref["col_src"],
                        ref["col_code"], n, T, F, dPs3, None, dtab3, zws, nb)
    close(gpu[-5], cpu[-5], 2e-5, "dPsrc (atomic variant)")
    close(gpu[-5], two_pass, 2e-5, "dPsrc atomic vs two-pass")
    close(gpu[-3], cpu[-3], 1e-4, "dtab (atomic variant)")


@pytest.mark.parametrize("precision", [0, 2])
@pytest.mark.parametrize("n,F", [(1000, 64), (70, 16), (3, 8), (300, 128), (257, 256), (90, 40)])
def test_tc_folded_stage(precision, n, F):
    """projsrc_fwd / postf_fwd / postf_wgrad / dxf_dgrad / projsrc_wgrad against their CPU restatements."""
    tol = TC_TOL[precision]
    T, Fo, C = 2, F // 2, 2 * F
    g = torch.Generator().manual_seed(n + F)
    x, A = torch.randn(n, F, generator=g), torch.randn(n, T * 4 * F, generator=g)
    scal = torch.rand(n, 2, generator=g) + 0.5
    pre_ts = F * 3 * F + F + 4
    pre = torch.randn(T * pre_ts, generator=g) * 0.1
    post_ts = Fo * 13 * F + Fo + 4
    post = torch.randn(T * post_ts, generator=g) * 0.1
    boff = Fo * 13 * F
    Vx = torch.randn(T, 3, Fo, F, generator=g) * 0.1
    ws = lambda dev: torch.zeros(4 << 20, dtype=torch.uint8, device=dev)

    def both(name, make_args, out_index):
        res = []
        for dev in ("cpu", DEV):
            args = make_args(dev)
            getattr(R, name)(*args) if dev == "cpu" else _lib.call(name, *args, _lib.stream_ptr())
            res.append(args[out_index])
        torch.cuda.synchronize()
        return res

    res = both("gnn_tc_pna_projsrc_fwd", lambda dev: [x.to(dev), pre.to(dev), pre_ts, pre.to(dev)[F * 3 * F:], pre_ts, n, F, T,
                                                      torch.zeros(n, C, device=dev), precision, ws(dev), 4 << 20], 8)
    close(res[1], res[0], tol, "projsrc fwd")
    for act in (0, 1):
        def mk(dev):
            b = post.to(dev)
            return [x.to(dev), A.to(dev), scal.to(dev), Vx.to(dev), b, post_ts, b[boff:], post_ts, n, F, Fo, T, act,
                    torch.zeros(n, T * Fo, device=dev), precision, ws(dev), 4 << 20]
        res = both("gnn_tc_pna_postf_fwd", mk, 13)
        close(res[1], res[0], tol, "postf fwd")
    dO, dPs = torch.randn(n, T * Fo, generator=g), torch.randn(n, C, generator=g)
    out = []
    for dev in ("cpu", DEV):
        gb, dv = torch.zeros(T * post_ts, device=dev), torch.zeros(T, 3, Fo, F, device=dev)
        args = [dO.to(dev), x.to(dev), A.to(dev), scal.to(dev), n, F, Fo, T, gb, post_ts, gb[boff:], post_ts, dv, precision]
        R.gnn_tc_pna_postf_wgrad(*args) if dev == "cpu" else _lib.call("gnn_tc_pna_postf_wgrad", *args, _lib.stream_ptr())
        out.append((gb, dv))
    close(out[1][0], out[0][0], 2 * tol, "postf wgrad (aggregate columns, bias)")
    close(out[1][1], out[0][1], 2 * tol, "postf wgrad dVx")
    for acc in (0, 1):
        res = both("gnn_tc_pna_dxf_dgrad", lambda dev: [dPs.to(dev), dO.to(dev), scal.to(dev), pre.to(dev), pre_ts, Vx.to(dev),
                                                        n, F, Fo, T, acc, torch.ones(n, F, device=dev) * 0.5, precision,
                                                        ws(dev), 4 << 20], 11)
        close(res[1], res[0], tol, "folded dx dgrad")

    def mkw(dev):
        gb = torch.zeros(T * pre_ts, device=dev)
        return [dPs.to(dev), x.to(dev), n, F, T, gb, pre_ts, gb[F * 3 * F:], pre_ts, precision]
    res = both("gnn_tc_pna_projsrc_wgrad", mkw, 5)
    close(res[1], res[0], 2 * tol, "projsrc wgrad")


@pytest.mark.parametrize("precision", [0, 2])
@pytest.mark.parametrize("M,N,K", [(1000, 64, 64), (4096, 32, 64), (130, 16, 32), (1, 8, 16)])
def test_bn_fold_linear_and_skip_gemm(precision, M, N, K):
    """Inference fast path: BatchNorm(eval) folded into the Linear, ReLU and skip add in the GEMM epilogue."""
    tol = TC_TOL[precision]
    g = torch.Generator().manual_seed(M + N)
    X, W, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) * 0.2, torch.randn(N, generator=g)
    gamma, beta = torch.rand(N, generator=g) + 0.5, torch.randn(N, generator=g)
    mean, var = torch.randn(N, generator=g), torch.rand(N, generator=g) + 0.1
    skip = torch.randn(M, N, generator=g)
    cpu, gpu = run_both("gnn_bn_fold_linear", W, K, b, gamma, beta, mean, var, 1e-5, N, K, torch.zeros(N, K), torch.zeros(N))
    close(gpu[-2], cpu[-2], 1e-6, "folded weight")
    close(gpu[-1], cpu[-1], 1e-6, "folded bias")
    Wf, bf = cpu[-2], cpu[-1]
    # the folded GEMM reproduces Linear -> BatchNorm1d(eval) -> ReLU -> + skip
    bn = torch.nn.BatchNorm1d(N).eval()
    with torch.no_grad():
        bn.weight.copy_(gamma), bn.bias.copy_(beta), bn.running_mean.copy_(mean), bn.running_var.copy_(var)
        want = torch.relu(bn(X @ W.t() + b)) + skip
    for sk in (skip, None):
        res = []
        for dev in ("cpu", DEV):
            args = [X.to(dev), K, Wf.to(dev), K, bf.to(dev), M, N, K, 1, None if sk is None else sk.to(dev), N,
                    torch.zeros(M, N, device=dev), N, precision, torch.zeros(4 << 20, dtype=torch.uint8, device=dev), 4 << 20]
            R.gnn_tc_linear_skip_fwd(*args) if dev == "cpu" else _lib.call("gnn_tc_linear_skip_fwd", *args, _lib.stream_ptr())
            res.append(args[11])
        close(res[1], res[0], tol, "skip gemm")
        if sk is not None:
            close(res[1], want, max(tol, 2e-5), "vs torch BatchNorm1d(eval)")


@pytest.mark.parametrize("n,e,F,act,store", [(700, 1500, 64, 0, True), (700, 1500, 64, 1, False), (130, 300, 16, 0, True),
                                             (1, 0, 8, 0, True), (260, 900, 128, 0, True), (90, 200, 40, 1, True),
                                             (300, 9000, 64, 0, True)])
def test_tc_postf_agg_fwd_fused(n, e, F, act, store):
    """gnn_tc_pna_postf_agg_fwd (aggregates computed by the GEMM's producers) against the two-kernel path on the GPU
    (aggregates BIT-identical to gnn_pna_aggregate_src_fwd, outputs identical up to the K order of the fp32
    accumulation) and against the CPU restatement.  (300, 9000): in-degrees above the 3-edge prefetch window."""
    T, Fo, C = 2, F // 2, 2 * F
    ref, ep, P, tab = _agg_inputs(n, e, T, F, seed=n + F, loops=True, quantise=(F == 16))
    Ps = P[:, C:].contiguous()
    g = torch.Generator().manual_seed(n + e + F)
    x = torch.randn(n, F, generator=g)
    scal = torch.rand(n, 2, generator=g) + 0.5
    post_ts = Fo * 13 * F + Fo + 4
    post = torch.randn(T * post_ts, generator=g) * 0.1
    Vx = torch.randn(T, 3, Fo, F, generator=g) * 0.1
    boff = Fo * 13 * F
    ws = torch.zeros(8 << 20, dtype=torch.uint8, device=DEV)
    dev_args = [t.to(DEV) for t in (x, Ps, tab, ref["rowptr"], ref["col_src"], ref["col_code"], scal, Vx, post)]
    xd, Pd, tabd, rp, cs, cc, scd, Vxd, pd = dev_args
    O = torch.full((n, T * Fo), 7.0, device=DEV)
    A = torch.full((n, T * 4 * F), -1.0, device=DEV) if store else None
    _lib.call("gnn_tc_pna_postf_agg_fwd", xd, Pd, tabd, rp, cs, cc, scd, Vxd, pd, post_ts, pd[boff:], post_ts, n, F, Fo, T,
              act, O, A, 0, ws, ws.numel(), _lib.stream_ptr())
    # two-kernel path
    A2 = torch.zeros(n, T * 4 * F, device=DEV)
    O2 = torch.zeros(n, T * Fo, device=DEV)
    _lib.call("gnn_pna_aggregate_src_fwd", Pd, tabd, rp, cs, cc, n, T, F, A2, _lib.stream_ptr())
    _lib.call("gnn_tc_pna_postf_fwd", xd, A2, scd, Vxd, pd, post_ts, pd[boff:], post_ts, n, F, Fo, T, act, O2, 0, ws,
              ws.numel(), _lib.stream_ptr())
    torch.cuda.synchronize()
    if store:
        assert torch.equal(A, A2), "fused aggregates must be bit-identical to the stand-alone kernel"
    close(O, O2.cpu(), 5e-6, "fused vs two-kernel post forward")
    # CPU restatement
    Oc = torch.zeros(n, T * Fo)
    Ac = torch.zeros(n, T * 4 * F)
    R.gnn_tc_pna_postf_agg_fwd(x, Ps, tab, ref["rowptr"], ref["col_src"], ref["col_code"], scal, Vx, post, post_ts,
                               post[boff:], post_ts, n, F, Fo, T, act, Oc, Ac, 0)
    close(O, Oc, 1e-5 if F != 16 else 1e-5, "fused post forward vs reference")
