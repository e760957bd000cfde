"""Drop-in `PNAPCSAFT` for the reference's gnnepcsaft/train/models.py:28-135, on sm_100a kernels.

Same constructor (`hidden_dim, pna_params, mlp_params`), same `forward(data)` contract
(`data.x [N,9] i64`, `data.edge_index [2,E] i64`, `data.edge_attr [E,3] i64`, `data.batch [N] i64`
or None -> `float32 [B, num_para]`) and the same `state_dict()` keys, so
reference checkpoints load with `strict=True`.  The submodules below only *hold* parameters
under the reference's names; the arithmetic runs in the CUDA library behind the C ABI
(include/gnnepcsaft_b200.h).  There is no CPU / PyG / torch_scatter fallback: CPU tensors raise.
"""

from __future__ import annotations

import collections
import contextlib
import dataclasses
import math
import os
from typing import List, Optional

import torch
from torch import nn

from .. import _lib
from ..data.graph import ATOM_FEATURE_DIMS, BOND_FEATURE_DIMS, NUM_BOND_CODES

TOWERS = 2  # reference models.py:76
# Source-node gradient scatter: by default 128-bit atomic adds inside the aggregate backward kernel (no per-slot dq
# buffer, no second pass: 139 us per layer on B200); GNN_B200_ATOMIC_SCATTER=0 selects the two-pass variant with a
# fixed summation order (per-slot gradients dq + gnn_slot_sum_by_source: 134 + 49 us per layer).  The weight-gradient
# split-K and bond-table sums use floating-point atomics in both variants.
DETERMINISTIC_SCATTER = os.environ.get("GNN_B200_ATOMIC_SCATTER", "1") == "0"
DEFAULT_MATMUL = "tf32x3"
# Fold the destination projection of PNAConv into post_nns' weights (csrc/fold.cu): the aggregate kernels gather
# source projections only.  Needs pre_layers == 1, self-loops (every node has an in-edge) and a tensor-core matmul mode;
# GNN_B200_FOLD=0 keeps the [Pdst | Psrc] formulation everywhere.
FOLD_DST = os.environ.get("GNN_B200_FOLD", "1") != "0"
# Compute the aggregates inside the post GEMM's operand producers (gnn_tc_pna_postf_agg_fwd) instead of in a separate
# kernel that writes them to HBM (folded formulation + TF32x3 only).  "infer" (default): forward passes that keep
# nothing for a backward pass - the aggregates then never reach HBM (154 vs 162 us per layer on 4096 graphs, 0.1 GB
# instead of 0.5 GB of DRAM traffic); "all": training forwards too (the kernel then also stores the aggregates:
# 183 us, slower than the two-kernel path); "off": never.
FUSE_AGG = os.environ.get("GNN_B200_FUSE_AGG", "infer")
# Store-resident inference of the default architecture: the whole PNAConv stack of a tile of <= 128 nodes (whole
# graphs) in one persistent CTA (gnn_tc_pna_stack_infer: node features and source projections stay in shared memory,
# one launch for all layers).  Bit-identical to the per-layer kernels, but measured slower on B200 (1.59 vs 1.46 ms per
# 4096 graphs: its two-stage operand ring leaves the tensor pipe waiting), so it is opt-in:
# GNN_B200_TILE_INFER=1.
TILE_INFER = os.environ.get("GNN_B200_TILE_INFER", "0") == "1"


@dataclasses.dataclass
class PnaconvsParams:
    "Parameters for pna convolutions (reference models.py:28-37)."
    propagation_depth: int
    pre_layers: int
    post_layers: int
    deg: torch.Tensor
    dropout: float = 0.0
    skip_connections: bool = False
    self_loops: bool = False


@dataclasses.dataclass
class ReadoutMLPParams:
    "Parameters for the MLP layers (reference models.py:40-45)."
    num_mlp_layers: int
    num_para: int
    dropout: float = 0.0


# ----------------------------------------------------------------------------- parameter holders
class _AtomEncoder(nn.Module):           # ogb AtomEncoder: atom_embedding_list.{k}.weight
    def __init__(self, emb_dim):
        super().__init__()
        self.atom_embedding_list = nn.ModuleList()
        for dim in ATOM_FEATURE_DIMS:
            emb = nn.Embedding(dim, emb_dim)
            nn.init.xavier_uniform_(emb.weight.data)
            self.atom_embedding_list.append(emb)


class _BondEncoder(nn.Module):           # ogb BondEncoder: bond_embedding_list.{k}.weight
    def __init__(self, emb_dim):
        super().__init__()
        self.bond_embedding_list = nn.ModuleList()
        for dim in BOND_FEATURE_DIMS:
            emb = nn.Embedding(dim, emb_dim)
            nn.init.xavier_uniform_(emb.weight.data)
            self.bond_embedding_list.append(emb)


class _DegreeScalerAggregation(nn.Module):   # PyG DegreeScalerAggregation buffers
    def __init__(self, deg: torch.Tensor):
        super().__init__()
        deg = deg.to(torch.float)
        num_nodes = int(deg.sum())
        bins = torch.arange(deg.numel())
        self.register_buffer("avg_deg_lin", torch.tensor([float((bins * deg).sum()) / num_nodes], dtype=torch.float))
        self.register_buffer("avg_deg_log",
                             torch.tensor([float(((bins + 1).log() * deg).sum()) / num_nodes], dtype=torch.float))


class _PNAConv(nn.Module):               # PyG PNAConv parameter layout (towers=2, divide_input=False)
    def __init__(self, hidden, deg, pre_layers, post_layers):
        super().__init__()
        f_in, f_out = hidden, hidden // TOWERS
        self.aggr_module = _DegreeScalerAggregation(deg)
        self.edge_encoder = nn.Linear(hidden, f_in)
        self.pre_nns = nn.ModuleList()
        self.post_nns = nn.ModuleList()
        for _ in range(TOWERS):
            mods = [nn.Linear(3 * f_in, f_in)]
            for _ in range(pre_layers - 1):
                mods += [nn.ReLU(), nn.Linear(f_in, f_in)]
            self.pre_nns.append(nn.Sequential(*mods))
            mods = [nn.Linear(13 * f_in, f_out)]
            for _ in range(post_layers - 1):
                mods += [nn.ReLU(), nn.Linear(f_out, f_out)]
            self.post_nns.append(nn.Sequential(*mods))
        self.lin = nn.Linear(hidden, hidden)


class _BatchNorm(nn.Module):             # PyG BatchNorm: `.module` = BatchNorm1d
    def __init__(self, channels):
        super().__init__()
        self.module = nn.BatchNorm1d(channels)


# ----------------------------------------------------------------------------- helpers
class KernelProfiler:
    """Optional CUDA-event timing of ABI calls on the launching stream (used by bench.py)."""

    def __init__(self, watch=None):
        self.watch = set(watch) if watch is not None else None
        self.records = []          # (name, start_event, end_event)

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for name, s, e in self.records:
            tot, cnt = out.get(name, (0.0, 0))
            out[name] = (tot + s.elapsed_time(e), cnt + 1)
        return out          # name -> (total ms, calls)


_PROFILER: Optional[KernelProfiler] = None
ABI_CALLS = 0               # number of C-ABI calls issued by this process (each launches >= 1 kernel)


def set_profiler(prof: Optional[KernelProfiler]):
    global _PROFILER
    _PROFILER = prof


def _c(name, *args):
    global ABI_CALLS
    ABI_CALLS += 1
    prof = _PROFILER
    if prof is not None and (prof.watch is None or name in prof.watch):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        _lib.call(name, *args, _lib.stream_ptr())
        e.record()
        prof.records.append((name, s, e))
    else:
        _lib.call(name, *args, _lib.stream_ptr())


_SIDE_STREAMS = {}


def _side_stream(dev: torch.device, i: int = 0) -> "torch.cuda.Stream":
    """Extra streams per device for small jobs that only have to finish before the end of the backward pass."""
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), i)
    st = _SIDE_STREAMS.get(key)
    if st is None:
        st = _SIDE_STREAMS[key] = torch.cuda.Stream(device=dev)
    return st


# dense-stage dispatch: which kernel family runs the Linear layers
MATMUL_MODES = {"simt": -1, "tf32x3": 0, "bf16": 2}


def _al(t, *dims):
    """128-bit accessible: 16-byte aligned base and all given dimensions multiples of 4."""
    return (t.data_ptr() & 15) == 0 and all((d & 3) == 0 for d in dims)


def _lin_fwd(mm, X, ldx, W, ldw, bias, M, N, K, act, Y, ldy, plan=None, key=None, pa=None):
    if mm < 0 or not _al(X, ldx, K):
        _c("gnn_linear_fwd", X, ldx, W, ldw, bias, M, N, K, act, Y, ldy)
    else:
        _gemm_site(plan, pa if key is not None else None, key, N, K, Y.device, "gnn_tc_linear_fwd", 5,
                   X, ldx, W, ldw, bias, M, N, K, act, Y, ldy, mm)


def _lin_dgrad(mm, dY, lddy, W, ldw, M, N, K, accumulate, dX, lddx, pa=None, key=None):
    if mm < 0 or not _al(dY, lddy, N):
        _c("gnn_linear_dgrad", dY, lddy, W, ldw, M, N, K, accumulate, dX, lddx)
    else:
        _gemm_site(None, pa if key is not None else None, key, K, N, dX.device, "gnn_tc_linear_dgrad", 4,
                   dY, lddy, W, ldw, M, N, K, accumulate, dX, lddx, mm)


def _lin_wgrad(mm, dY, lddy, X, ldx, M, N, K, dW, lddw, db):
    if mm < 0 or not (_al(dY, lddy, N) and _al(X, ldx, K)):
        _c("gnn_linear_wgrad", dY, lddy, X, ldx, M, N, K, dW, lddw, db)
    else:
        _c("gnn_tc_linear_wgrad", dY, lddy, X, ldx, M, N, K, dW, lddw, db, mm)


def _post_fwd(mm, *args, ws=None):
    if mm < 0:
        _c("gnn_pna_post_fwd", *args)
    else:
        buf, nb = ws if ws is not None else (_pack_ws(args[-1].device), _pack_ws(args[-1].device).numel())
        _c("gnn_tc_pna_post_fwd", *args, mm, buf, nb)


def _post_dgrad(mm, *args, pa=None, key=None):
    if mm < 0:
        _c("gnn_pna_post_dgrad", *args)
    else:
        # (dO, scal, wpost, tower stride, n, F, F_out, towers, dx, gA): image = towers x [4F rows, K = 3 F_out]; with
        # dx given the entry point runs further GEMMs through the same workspace -> shared scratch
        n_rows, k = args[7] * 4 * args[5], 3 * args[6]
        _gemm_site(None, pa if (key is not None and args[8] is None) else None, key, n_rows, k, args[-1].device,
                   "gnn_tc_pna_post_dgrad", 4, *args, mm)


def _post_wgrad(mm, *args):
    if mm < 0:
        _c("gnn_pna_post_wgrad", *args)
    else:
        _c("gnn_tc_pna_post_wgrad", *args, mm)


class _Workspace:
    """Grow-only scratch buffers per device (BN partials, CSR scratch).  A buffer that is outgrown is RETIRED, not
    freed: captured CUDA graphs keep launching kernels with its address, so its memory must never be handed to
    another tensor (a replay would scribble BatchNorm partial sums over it).  Growth doubles, so a handful of small
    buffers is retired at most."""

    def __init__(self):
        self.buf = {}
        self.retired = []

    def _grow(self, key, nbytes, device, zero):
        old = self.buf.get((key, device))
        size = max(int(nbytes), 256, 2 * old.numel() if old is not None else 0)
        t = (torch.zeros if zero else torch.empty)(size, dtype=torch.uint8, device=device)
        if old is not None:
            self.retired.append(old)
        self.buf[(key, device)] = t
        return t

    def get(self, key, nbytes, device):
        t = self.buf.get((key, device))
        if t is None or t.numel() < nbytes:
            t = self._grow(key, nbytes, device, False)
        return t

    def get_zeroed(self, key, nbytes, device):
        """Scratch whose users leave it all-zero (so it is cleared only when (re)allocated)."""
        t = self.buf.get((key, device))
        if t is None or t.numel() < nbytes:
            t = self._grow(key, nbytes, device, True)
        return t


_WS = _Workspace()
PACK_WS_BYTES = 32 << 20      # covers the wide model (H = 256: 7.9 MB for the post GEMM image)


def _pack_ws(device):
    """Scratch for the packed weight operand of the tensor-core GEMMs (stream-ordered reuse)."""
    return _WS.get("pack", PACK_WS_BYTES, device)


class _InferPlan:
    """Weight-derived tensors of the inference fast path - bond tables, folded post / lin weights, the packed
    tensor-core image of every GEMM call site - kept across forwards for as long as the weights do not change, so that
    an eval forward (and a captured eval graph) launches only the kernels that depend on the batch.  Buffers are
    allocated once and refilled IN PLACE when the weights change: captured graphs keep their addresses."""

    def __init__(self, device):
        self.device = device
        self.token = None           # weights version the contents were computed for
        self.pending = None
        self.filling = True         # this forward (re)computes the contents
        self.tensors = {}
        self.sites = {}

    def tensor(self, key, shape, dtype=torch.float32):
        t = self.tensors.get(key)
        if t is None:
            if not self.filling:
                raise RuntimeError(f"inference plan: {key} was not produced by the pass that filled the plan")
            t = torch.empty(shape, dtype=dtype, device=self.device)
            self.tensors[key] = t
        return t

    def site(self, key, nbytes):
        """(pack workspace, signed size) of a GEMM call site: negative size = image already packed (header)."""
        buf = self.sites.get(key)
        if buf is None or buf.numel() < nbytes:
            if not self.filling:
                raise RuntimeError(f"inference plan: the packed image of {key} was not produced by the pass that "
                                   "filled the plan")
            buf = torch.empty((int(nbytes) + 4095) // 4096 * 4096, dtype=torch.uint8, device=self.device)
            self.sites[key] = buf
        return buf, (buf.numel() if self.filling else -buf.numel())


class _PackAhead:
    """Training steps under CUDA-graph capture: the pack kernel of every tensor-core GEMM call site (weights -> the
    swizzled hi/lo tile image) runs on a second stream that depends on the weights only, into a buffer of its own, and
    the GEMM is launched with the `already packed` convention of the ABI (row count 0 = pack only, negative workspace
    size = packed).  In the captured graph the ~45 pack kernels of a step form a branch beside the data path instead of
    sitting between consecutive GEMMs.
    The per-site buffers are allocated by EAGER steps (the engine runs one before it captures) and live as long as the
    model: a block handed out by the caching allocator inside a capture may have belonged to a tensor whose last
    main-stream user the pack branch is not ordered after.  Eager steps keep the in-line pack (two more host calls per
    GEMM would cost more than the few microseconds of GPU time); a site without a buffer packs in line too."""

    PACK_AHEAD = os.environ.get("GNN_B200_PACK_AHEAD", "1") != "0"
    _streams = {}

    def __init__(self, model, dev, capturing):
        self.dev = dev
        self.capturing = capturing
        key = dev.index if dev.index is not None else torch.cuda.current_device()
        if not hasattr(model, "_train_sites"):
            model._train_sites = {}
        self.sites = model._train_sites.setdefault(key, {})      # per device: the model may be moved
        if capturing:
            if key not in _PackAhead._streams:
                _PackAhead._streams[key] = torch.cuda.Stream(device=dev)
            self.side = _PackAhead._streams[key]
            fork = torch.cuda.Event()
            fork.record()                       # the weights (and the folded Vx) are final from here on
            self.side.wait_event(fork)

    @staticmethod
    def begin(model, dev, training, save, mm):
        if _PackAhead.PACK_AHEAD and training and save and mm >= 0 and dev.type == "cuda":
            return _PackAhead(model, dev, torch.cuda.is_current_stream_capturing())
        return None

    def call(self, key, nbytes, name, n_idx, args):
        buf = self.sites.get(key)
        if not self.capturing:
            if buf is None or buf.numel() < nbytes:
                self.sites[key] = torch.empty((int(nbytes) + 4095) // 4096 * 4096, dtype=torch.uint8, device=self.dev)
            buf = None
        if buf is None or buf.numel() < nbytes:
            ws = _pack_ws(self.dev)
            _c(name, *args, ws, ws.numel())
            return
        a0 = list(args)
        a0[n_idx] = 0
        with torch.cuda.stream(self.side):
            _c(name, *a0, buf, buf.numel())
            ev = torch.cuda.Event()
            ev.record()
        torch.cuda.current_stream().wait_event(ev)
        _c(name, *args, buf, -buf.numel())


def _gemm_site(plan, pa, key, n_rows, k, dev, name, n_idx, *args):
    """Tensor-core GEMM with a packed weight operand; `args` = the ABI arguments before (pack_ws, pack_ws_bytes) and
    args[n_idx] the row count.  plan: inference (image persists while the weights do); pa: training step."""
    if plan is not None:
        ws, nb = plan.site(key, _pack_bound(n_rows, k))
        _c(name, *args, ws, nb)
    elif pa is not None:
        pa.call(key, _pack_bound(n_rows, k), name, n_idx, args)
    else:
        ws = _pack_ws(dev)
        _c(name, *args, ws, ws.numel())


def _pack_bound(n_rows: int, k: int) -> int:
    """Upper bound of the packed image (tf32 hi + lo planes) of a weight operand with n_rows rows and K = k."""
    return 8 * (n_rows + 64) * ((k + 31) // 32 * 32 + 32)


def _site_ws(plan, key, n_rows, k, device):
    if plan is None:
        ws = _pack_ws(device)
        return ws, ws.numel()
    return plan.site(key, _pack_bound(n_rows, k))


_INPUT_ERRORS = ((1, IndexError, "edge_index out of range [0, num_nodes)"),
                 (2, IndexError, "edge_attr (bond feature) out of range"),
                 (4, IndexError, "batch id out of range [0, num_graphs)"),
                 (8, ValueError, "data.batch is not sorted (PyG Batch vectors are; sort the nodes by graph)"),
                 (16, IndexError, "x (atom feature) out of range"))
# GNN_B200_CHECK=1: read the device-side input error word right after every forward (synchronises).  Default: the
# word of a forward is read, without blocking, by a later forward / step once its kernels have finished, and by
# every call that synchronises anyway (`check_input_errors()`, TrainEngine.loss_value, smoke tests).
CHECK_INPUTS_EAGERLY = os.environ.get("GNN_B200_CHECK", "0") == "1"
_PENDING_ERR = collections.deque()


def _raise_input_error(code: int):
    for bit, exc, msg in _INPUT_ERRORS:
        if code & bit:
            raise exc(f"gnnepcsaft_b200: {msg} (the reference's torch indexing raises on the same input)")


def check_input_errors(block: bool = True):
    """Raise IndexError / ValueError for malformed inputs of earlier forwards (the kernels clamp and flag instead of
    faulting).  block=False only looks at forwards whose kernels have already completed."""
    if _PENDING_ERR and torch.cuda.is_available() and torch.cuda.is_current_stream_capturing():
        return                      # event queries are not allowed while a CUDA graph is being captured
    while _PENDING_ERR:
        err, ev = _PENDING_ERR[0]
        if not block and not ev.query():
            return
        _PENDING_ERR.popleft()
        code = int(err.item())
        if code:
            _PENDING_ERR.clear()
            _raise_input_error(code)


class GraphStructure:
    """Device-side CSR / slot lists of one batch (gnn_build_csr, gnn_graph_ptr_from_batch)."""

    def __init__(self, edge_index, edge_attr, batch, num_nodes, num_graphs, self_loops, ptr=None):
        dev = edge_index.device
        n, e = int(num_nodes), int(edge_index.shape[1])
        ep = e + (n if self_loops else 0)
        self.n, self.e, self.ep, self.b = n, e, ep, int(num_graphs)
        i32 = dict(dtype=torch.int32, device=dev)
        self.rowptr = torch.empty(n + 1, **i32)
        self.srcptr = torch.empty(n + 1, **i32)
        self.col_src = torch.empty(max(ep, 1), **i32)
        self.col_eid = torch.empty(max(ep, 1), **i32)
        self.src_slot = torch.empty(max(ep, 1), **i32)
        self.col_code = torch.empty(max(ep, 1), dtype=torch.uint8, device=dev)
        self.err = torch.zeros(1, **i32)
        nbytes = _lib.query("gnn_build_csr_workspace_bytes", n, e)
        ws = _WS.get("csr", nbytes, dev)
        _c("gnn_build_csr", edge_index, edge_attr, n, e, 1 if self_loops else 0, self.rowptr, self.col_src,
           self.col_code, self.col_eid, self.srcptr, self.src_slot, self.err, ws, ws.numel())
        self.graph_ptr = torch.empty(self.b + 1, **i32)
        if batch is None:
            self.node_graph = None
            _c("gnn_graph_ptr_single", n, self.graph_ptr)
        else:
            self.node_graph = torch.empty(max(n, 1), **i32)
            _c("gnn_graph_ptr_from_batch", batch, n, self.b, self.node_graph, self.graph_ptr, self.err)

    def defer_check(self):
        """Queue this batch's error word (call after the last kernel that writes it was enqueued)."""
        if self.err.is_cuda and torch.cuda.is_current_stream_capturing():
            return
        if CHECK_INPUTS_EAGERLY or not self.err.is_cuda:
            _raise_input_error(int(self.err.item()))
            return
        ev = torch.cuda.Event()
        ev.record()
        _PENDING_ERR.append((self.err, ev))

    def check(self):
        """Synchronising validity check (the reference would raise IndexError inside torch)."""
        _raise_input_error(int(self.err.item()))


def _device_guard(dev: torch.device):
    """Make `dev` the current CUDA device for the enclosed launches (streams, scratch buffers and the SM count all
    follow the current device); a no-op for the host tensors of the CPU-side host-logic tests."""
    return torch.cuda.device(dev) if dev.type == "cuda" else contextlib.nullcontext()


def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(
            f"gnnepcsaft_b200: {what} is on {t.device}; this implementation has no CPU path "
            "(move the model and the batch to a CUDA device)")


# ----------------------------------------------------------------------------- the model
class PNAPCSAFT(nn.Module):
    """Graph neural network to predict ePC-SAFT parameters (reference models.py:48-135)."""

    def __init__(self, hidden_dim: int, pna_params: PnaconvsParams, mlp_params: ReadoutMLPParams):
        super().__init__()
        if hidden_dim % 4 != 0:
            raise ValueError("hidden_dim must be divisible by 4 (towers=2, head uses H//2 and H//4)")
        if hidden_dim % 8 != 0:
            raise ValueError("gnnepcsaft_b200 kernels use 128-bit accesses on per-tower rows: hidden_dim must be a "
                             "multiple of 8 (the reference accepts multiples of 4; its configs use 64 / 128 / 256)")
        self.hidden_dim = hidden_dim
        self.pna_params = pna_params
        self.mlp_params = mlp_params
        self.convs = nn.ModuleList()
        self.batch_norms = nn.ModuleList()
        self.node_embed = _AtomEncoder(hidden_dim)
        self.edge_embed = _BondEncoder(hidden_dim)
        for _ in range(pna_params.propagation_depth):
            self.convs.append(_PNAConv(hidden_dim, pna_params.deg, pna_params.pre_layers, pna_params.post_layers))
            self.batch_norms.append(_BatchNorm(hidden_dim))
        self.mlp = nn.Sequential()
        for _ in range(mlp_params.num_mlp_layers):
            self.mlp.append(nn.Linear(hidden_dim, hidden_dim))
            self.mlp.append(nn.BatchNorm1d(hidden_dim))
            self.mlp.append(nn.ReLU())
            self.mlp.append(nn.Dropout(p=mlp_params.dropout))
        self.mlp.append(nn.Sequential(
            nn.Linear(hidden_dim, hidden_dim // 2), nn.BatchNorm1d(hidden_dim // 2), nn.ReLU(),
            nn.Dropout(p=mlp_params.dropout),
            nn.Linear(hidden_dim // 2, hidden_dim // 4), nn.BatchNorm1d(hidden_dim // 4), nn.ReLU(),
            nn.Dropout(p=mlp_params.dropout),
            nn.Linear(hidden_dim // 4, mlp_params.num_para)))
        self._arena: Optional[torch.Tensor] = None
        self._arena_offsets: List[int] = []
        self._arena_ends = ()
        self._row_off = None
        self._plan = None           # inference plans by kernel path
        self._plan_tensors = None
        self._weights_epoch = 0     # bumped by everything that changes weights / BatchNorm buffers through raw pointers
        # dense stages: "tf32x3" = tcgen05 with fp32 operands split into tf32 hi+lo (fp32-level accuracy),
        # "bf16" = tcgen05 bf16 operands / fp32 accumulate (torch "medium" matmul precision, which the
        # reference itself sets at train.py:319), "simt" = plain fp32 FMA kernels
        self.matmul_precision = os.environ.get("GNN_B200_MATMUL", DEFAULT_MATMUL)

    def _fold_dst(self, mm: int) -> bool:
        pp = self.pna_params
        return (FOLD_DST and mm >= 0 and pp.pre_layers == 1 and bool(pp.self_loops)
                and (self.hidden_dim // TOWERS) % 4 == 0)

    def _matmul_mode(self) -> int:
        try:
            mm = MATMUL_MODES[self.matmul_precision]
        except KeyError:
            raise ValueError(f"matmul_precision must be one of {sorted(MATMUL_MODES)}") from None
        return mm

    # ---- flat parameter arena -------------------------------------------------------------
    def _param_list(self):
        return list(self.parameters())

    def _arena_valid(self) -> bool:
        if self._arena is None:
            return False
        # the first and the last parameter still view the arena (nn.Module keeps the Parameter objects; walking the
        # module tree for them on every call cost 0.1 ms of host time per forward)
        base = self._arena.data_ptr()
        for p, off in self._arena_ends:
            if p.data_ptr() != base + 4 * off or p.device != self._arena.device:
                return False
        return True

    def flatten_parameters(self) -> torch.Tensor:
        """(Re)build the flat fp32 arena; every parameter becomes a view at a 16-byte aligned offset.

        Layer l's conv parameters sit at a constant stride, the 9 atom tables / 3 bond tables are
        contiguous; kernels rely on both.  Called lazily by forward after `.to()` / `.cuda()`.
        """
        ps = self._param_list()
        for p in ps:
            if p.dtype != torch.float32:
                raise RuntimeError("gnnepcsaft_b200 kernels are fp32; float64 / half models are not supported "
                                   "(the reference's `.to(torch.float64)` evaluation must stay on fp32 here)")
        offsets, total = [], 0
        for p in ps:
            offsets.append(total)
            total += (p.numel() + 3) // 4 * 4
        arena = torch.zeros(total, dtype=torch.float32, device=ps[0].device)
        with torch.no_grad():
            for p, off in zip(ps, offsets):
                view = arena[off:off + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
        self._arena, self._arena_offsets = arena, offsets
        self._arena_ends = ((ps[0], offsets[0]), (ps[-1], offsets[-1]))
        self._row_off = None
        return arena

    def arena(self) -> torch.Tensor:
        if not self._arena_valid():
            self.flatten_parameters()
        return self._arena

    def _offset_of(self, p: torch.Tensor) -> int:
        return (p.data_ptr() - self._arena.data_ptr()) // 4

    def _infer_plan(self, device, kind: str = "layers", training: Optional[bool] = None) -> _InferPlan:
        """The inference plan for the current weights (see _InferPlan); `filling` tells the forward whether its
        weight-derived tensors have to be (re)computed."""
        arena = self.arena()
        if self._plan_tensors is None or self._plan_tensors[0] is not arena:
            # (walking the module tree costs ~0.2 ms per call: the tensors whose versions make up the token are listed
            # once per arena; nn.Module keeps Parameter / buffer objects across in-place updates and load_state_dict)
            self._plan_tensors = (arena, self._param_list() + list(self.buffers()))
        token = (arena.data_ptr(), arena._version, self._weights_epoch, sum(t._version for t in self._plan_tensors[1]),
                 self.matmul_precision, TILE_INFER, FUSE_AGG, FOLD_DST)
        if self._plan is None:
            self._plan = {}
        plan = self._plan.get(kind)             # one plan per kernel path ("layers" / "tiles"): different contents
        if plan is None or plan.device != device:
            plan = self._plan[kind] = _InferPlan(device)
        plan.filling = plan.token != token
        plan.pending = token
        return plan

    def _consts(self, device):
        if self._row_off is None or self._row_off.device != device:
            off = [0]
            for d in ATOM_FEATURE_DIMS[:-1]:
                off.append(off[-1] + d)
            self._row_off = torch.tensor(off, dtype=torch.int32, device=device)
        return self._row_off

    # ---- forward ----------------------------------------------------------------------------
    def forward(self, data):
        """Forward pass of the model (reference models.py:105-135)."""
        x, edge_index, edge_attr, batch = data.x, data.edge_index, data.edge_attr, getattr(data, "batch", None)
        _require_cuda(x, "data.x")
        # every launch of this call goes to the current stream OF THE DATA'S DEVICE, whatever device is current
        with _device_guard(x.device):
            arena = self.arena()
            _require_cuda(arena, "the model")
            _same_device(x.device, model=arena, edge_index=edge_index, edge_attr=edge_attr, batch=batch)
            check_input_errors(block=False)
            if batch is None:
                num_graphs = 1
            else:
                num_graphs = getattr(data, "num_graphs", None)
                if num_graphs is None:
                    ptr = getattr(data, "ptr", None)
                    num_graphs = int(ptr.numel()) - 1 if ptr is not None else int(batch.max()) + 1
            params = self._param_list()
            if torch.is_grad_enabled() and any(p.requires_grad for p in params):
                return _PNAFunction.apply(self, x, edge_index, edge_attr, batch, int(num_graphs), *params)
            out, _ = _forward_impl(self, x, edge_index, edge_attr, batch, int(num_graphs), save=False)
            return out


def _same_device(dev, **tensors):
    for name, t in tensors.items():
        if t is not None and t.device != dev:
            raise RuntimeError(f"gnnepcsaft_b200: {name} is on {t.device} but data.x is on {dev}; the model and every "
                               "tensor of the batch must live on the same CUDA device")


def _check_inputs(x, edge_index, edge_attr, batch):
    for name, t in (("x", x), ("edge_index", edge_index), ("edge_attr", edge_attr)):
        _require_cuda(t, f"data.{name}")
        if t.dtype != torch.int64:
            raise TypeError(f"data.{name} must be int64 (OGB integer features), got {t.dtype}")
    if x.dim() != 2 or x.shape[1] != 9:
        raise ValueError(f"data.x must be [N, 9], got {tuple(x.shape)}")
    if edge_index.dim() != 2 or edge_index.shape[0] != 2:
        raise ValueError(f"data.edge_index must be [2, E], got {tuple(edge_index.shape)}")
    if edge_attr.shape[0] != edge_index.shape[1] or (edge_attr.numel() and edge_attr.shape[1] != 3):
        raise ValueError(f"data.edge_attr must be [E, 3], got {tuple(edge_attr.shape)}")
    if batch is not None and (batch.dtype != torch.int64 or batch.shape[0] != x.shape[0]):
        raise ValueError("data.batch must be int64 [N]")


class _Saved:
    pass


def _bn_forward(y, bn: nn.BatchNorm1d, training: bool, skip, relu: bool, valid=None):
    """`valid` (device int32[1]): padded batch, only the first *valid rows enter the statistics."""
    rows, ch = y.shape
    dev = y.device
    stats = torch.empty(2 * ch, dtype=torch.float32, device=dev)
    use_batch = training or bn.running_mean is None
    nbytes = _lib.query("gnn_bn_workspace_bytes", rows, ch)
    ws = _WS.get("bn", nbytes, dev)
    if use_batch and rows == 1 and valid is None:
        raise ValueError(f"Expected more than 1 value per channel when training, got input size {tuple(y.shape)}")
    if bn.momentum is None:     # torch: cumulative moving average, factor 1 / num_batches_tracked (after the increment)
        if y.is_cuda and torch.cuda.is_current_stream_capturing():
            raise RuntimeError("gnnepcsaft_b200: BatchNorm momentum=None reads num_batches_tracked on the host and "
                               "cannot be captured in a CUDA graph (use the eager step)")
        momentum = 1.0 / (int(bn.num_batches_tracked) + 1) if (use_batch and bn.num_batches_tracked is not None) else 0.0
    else:
        momentum = bn.momentum
    _c("gnn_bn_stats", y, rows, ch, 1 if use_batch else 0, float(bn.eps), float(momentum), bn.running_mean,
       bn.running_var, bn.num_batches_tracked if use_batch else None, stats, valid, ws, ws.numel())
    out = torch.empty_like(y)
    _c("gnn_bn_act_fwd", y, stats, bn.weight, bn.bias, skip, rows, ch, 1 if relu else 0, out)
    return out, stats, use_batch


def _lin_bn_eval(mm, X, lin: nn.Linear, bn: nn.BatchNorm1d, skip, relu: bool, plan=None, key=None):
    """Inference fast path: Linear -> BatchNorm(eval) -> ReLU (-> + skip) as ONE tensor-core GEMM with the running
    statistics folded into the weights (gnn_bn_fold_linear + gnn_tc_linear_skip_fwd).  Returns None when the shape
    needs the general path."""
    rows, k, nout = X.shape[0], lin.in_features, lin.out_features
    if mm < 0 or bn.running_mean is None or (k & 3) or (nout & 3) or not _al(X, k) or rows == 0:
        return None
    dev = X.device
    if plan is not None:
        wf, bf = plan.tensor((key, "wf"), (nout, k)), plan.tensor((key, "bf"), (nout,))
    else:
        wf = torch.empty(nout, k, dtype=torch.float32, device=dev)
        bf = torch.empty(nout, dtype=torch.float32, device=dev)
    if plan is None or plan.filling:
        _c("gnn_bn_fold_linear", lin.weight, k, lin.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var,
           float(bn.eps), nout, k, wf, bf)
    out = torch.empty(rows, nout, dtype=torch.float32, device=dev)
    ws, nb = _site_ws(plan, key, nout, k, dev)
    _c("gnn_tc_linear_skip_fwd", X, k, wf, k, bf, rows, nout, k, 1 if relu else 0, skip, nout, out, nout, mm, ws, nb)
    return out


def _forward_impl(model: PNAPCSAFT, x_idx, edge_index, edge_attr, batch, num_graphs, save: bool, graph=None,
                  training: Optional[bool] = None):
    """The forward kernel chain.  `graph`: a prebuilt `GraphStructure`, or a compact store batch
    (`data.compact.CompactBatch`: uint8 atom features + int32 CSR straight from `gnn_compact_collate`, possibly padded
    to a shape bucket - then x_idx / edge_index / edge_attr / batch are ignored and the output has one extra row, the
    padding graph)."""
    compact = graph is not None and getattr(graph, "is_compact", False)
    pp, mp = model.pna_params, model.mlp_params
    valid_nodes = valid_graphs = None
    if compact:
        if bool(graph.self_loops) != bool(pp.self_loops):
            raise ValueError("the compact batch was built with self_loops=%s but the model has self_loops=%s"
                             % (graph.self_loops, pp.self_loops))
        if pp.pre_layers != 1 or DETERMINISTIC_SCATTER:
            raise NotImplementedError(
                "gnnepcsaft_b200: compact store batches carry no per-source slot lists; models with pre_layers > 1 and "
                "the fixed-order gradient scatter (GNN_B200_ATOMIC_SCATTER=0) take PyG-layout batches (GraphStore / "
                "forward(data))")
        x_idx, num_graphs = graph.x, graph.b
        _require_cuda(x_idx, "the compact batch")
        valid_nodes, valid_graphs = graph.valid[0:1], graph.valid[1:2]
    else:
        _check_inputs(x_idx, edge_index, edge_attr, batch)
    if (model.training if training is None else training) and (pp.dropout > 0.0 or mp.dropout > 0.0):
        raise NotImplementedError("gnnepcsaft_b200: dropout > 0 in training mode is not implemented "
                                  "(the reference configs use dropout_rate = 0.0)")
    dev = x_idx.device
    mm = model._matmul_mode()
    H = model.hidden_dim
    F, Fo, T = H, H // TOWERS, TOWERS
    C = T * F
    L = pp.propagation_depth
    x_idx = x_idx.contiguous()
    if not compact:
        edge_index = edge_index.contiguous()
        edge_attr = edge_attr.contiguous()
        if batch is not None:
            batch = batch.contiguous()
    n = int(graph.n) if compact else int(x_idx.shape[0])
    training = model.training if training is None else bool(training)      # engines pass the mode: no module-tree flips
    f32 = dict(dtype=torch.float32, device=dev)

    g = graph if graph is not None else GraphStructure(edge_index, edge_attr, batch, n, num_graphs, pp.self_loops)
    fold = model._fold_dst(mm)
    row_off = model._consts(dev)
    sv = _Saved() if save else None
    # inference with a tensor-core matmul mode: weight-derived tensors and packed weight images persist in the model's
    # inference plan and are only recomputed when the weights changed (fill)
    use_tiles = (TILE_INFER and compact and not save and not training and fold and mm == 0 and L > 0 and H == 64
                 and pp.post_layers == 1 and bool(getattr(graph, "tileable", False))
                 and all(b.module.running_mean is not None for b in model.batch_norms))
    plan = model._infer_plan(dev, "tiles" if use_tiles else "layers") if (not save and not training and mm >= 0) else None
    fill = plan is None or plan.filling
    if training:
        model._weights_epoch += 1       # the BatchNorm running statistics are about to change (raw-pointer kernels)

    # AtomEncoder: the 9 tables are contiguous in the arena
    atom_tab = model.node_embed.atom_embedding_list[0].weight
    total_rows = sum(ATOM_FEATURE_DIMS)
    x = torch.empty(n, H, **f32)
    _c("gnn_atom_embed_u8_fwd" if compact else "gnn_atom_embed_fwd", x_idx, n, atom_tab, row_off, total_rows, H, x,
       g.err)
    if graph is None:
        g.defer_check()

    # bond-code tables of all layers in one launch
    conv0 = model.convs[0] if L > 0 else None
    enc = tab = None
    layer_stride = tower_stride_pre = 0
    if L > 0:
        bond_tab = model.edge_embed.bond_embedding_list[0].weight
        layer_stride = (model._offset_of(model.convs[1].edge_encoder.weight) -
                        model._offset_of(conv0.edge_encoder.weight)) if L > 1 else 0
        tower_stride_pre = model._offset_of(conv0.pre_nns[1][0].weight) - model._offset_of(conv0.pre_nns[0][0].weight)
        if plan is not None:
            enc, tab = plan.tensor("enc", (L, NUM_BOND_CODES, F)), plan.tensor("tab", (L, NUM_BOND_CODES, C))
        else:
            enc = torch.empty(L, NUM_BOND_CODES, F, **f32)
            tab = torch.empty(L, NUM_BOND_CODES, C, **f32)
        if fill:
            _c("gnn_bond_table_fwd", bond_tab, conv0.edge_encoder.weight, conv0.edge_encoder.bias,
               conv0.pre_nns[0][0].weight, tower_stride_pre, layer_stride, L, H, T, enc, tab)

    Vx_all = None
    if fold and L > 0:      # folded x columns of post_nns[t][0] for all layers (one launch)
        tsp = model._offset_of(conv0.post_nns[1][0].weight) - model._offset_of(conv0.post_nns[0][0].weight)
        Vx_all = plan.tensor("Vx", (L, T, 3, Fo, F)) if plan is not None else torch.empty(L, T, 3, Fo, F, **f32)
        if fill:
            _c("gnn_pna_fold_fwd", conv0.post_nns[0][0].weight, tsp, conv0.pre_nns[0][0].weight, tower_stride_pre,
               layer_stride, L, F, Fo, T, Vx_all)

    pa = _PackAhead.begin(model, dev, training, save, mm)     # after the fold: the dgrad images read Vx

    if use_tiles:
        x = _conv_stack_tiles(model, x, graph, tab, Vx_all, layer_stride, tower_stride_pre, plan)

    layers = []
    for l in range(L if use_tiles else 0, L):
        conv, bn = model.convs[l], model.batch_norms[l].module
        rec = _Saved()
        rec.x_in = x
        # node projections  P = [Pdst_t0 | Pdst_t1 | Psrc_t0 | Psrc_t1]   (folded: [Psrc_t0 | Psrc_t1] only)
        P = torch.empty(n, C if fold else 2 * C, **f32)
        mm_proj = 0 if mm == 2 else mm      # bf16 mode: messages stay fp32-accurate (min/max/std are selection-sensitive)
        pre0 = conv.pre_nns[0][0]
        ts_pre = model._offset_of(conv.pre_nns[1][0].weight) - model._offset_of(pre0.weight)
        bs_pre = model._offset_of(conv.pre_nns[1][0].bias) - model._offset_of(pre0.bias)
        rec.pre_strides = (ts_pre, bs_pre)
        tower_stride_post = model._offset_of(conv.post_nns[1][0].weight) - model._offset_of(conv.post_nns[0][0].weight)
        bias_stride_post = model._offset_of(conv.post_nns[1][0].bias) - model._offset_of(conv.post_nns[0][0].bias)
        rec.post_strides = (tower_stride_post, bias_stride_post)
        rec.Vx = None
        if fold:
            rec.Vx = Vx_all[l]
            _gemm_site(plan, pa, ("projsrc", l), C, F, dev, "gnn_tc_pna_projsrc_fwd", 5,
                       x, pre0.weight, ts_pre, pre0.bias, bs_pre, n, F, T, P, mm_proj)
        elif mm_proj >= 0:
            _gemm_site(plan, pa, ("proj", l), 2 * C, F, dev, "gnn_tc_pna_proj_fwd", 5,
                       x, pre0.weight, ts_pre, pre0.bias, bs_pre, n, F, T, P, mm_proj)
        else:
            for t in range(T):
                w = conv.pre_nns[t][0].weight
                b = conv.pre_nns[t][0].bias
                _lin_fwd(mm_proj, x, H, w, 3 * F, b, n, F, F, 0, P[:, t * F:], 2 * C)
                _lin_fwd(mm_proj, x, H, w[:, F:], 3 * F, None, n, F, F, 0, P[:, C + t * F:], 2 * C)
        fuse_agg = (fold and mm == 0 and n > 0 and Fo <= 64 and
                    (FUSE_AGG == "all" or (FUSE_AGG == "infer" and not save)))
        # the aggregates only reach HBM when the backward pass needs them
        A = torch.empty(n, T * 4 * F, **f32) if (save or not fuse_agg) else None
        rec.P = P
        if fuse_agg:
            pass                                    # aggregated by the producers of the post GEMM below
        elif fold:
            _c("gnn_pna_aggregate_src_fwd", P, tab[l], g.rowptr, g.col_src, g.col_code, n, T, F, A)
        elif pp.pre_layers == 1:
            _c("gnn_pna_aggregate_fwd", P, tab[l], g.rowptr, g.col_src, g.col_code, n, T, F, A)
        else:
            hs = []
            h = torch.empty(max(g.ep, 1), C, **f32)
            _c("gnn_edge_hidden_fwd", P, tab[l], g.rowptr, g.col_src, g.col_code, n, C, 1, h)
            hs.append(h)
            for k in range(1, pp.pre_layers):
                h2 = torch.empty(max(g.ep, 1), C, **f32)
                last = k == pp.pre_layers - 1
                for t in range(T):
                    lin = conv.pre_nns[t][2 * k]
                    _lin_fwd(mm, h[:, t * F:], C, lin.weight, F, lin.bias, g.ep, F, F, 0 if last else 1,
                       h2[:, t * F:], C)
                hs.append(h2)
                h = h2
            rec.hs = hs
            _c("gnn_pna_aggregate_msg_fwd", h, g.rowptr, n, T, F, A)
        rec.A = A
        scal = torch.empty(n, 2, **f32)
        _c("gnn_pna_scalers", g.rowptr, conv.aggr_module.avg_deg_log, n, scal)
        rec.scal = scal
        # post_nns (virtual 13F concat) -> lin
        O = torch.empty(n, H, **f32)
        post_rows = T * 3 * ((Fo + 63) // 64) * 64         # rows of the packed post image (towers x scaler groups x tiles)
        if fuse_agg:
            _gemm_site(plan, pa, ("postf_agg", l), post_rows, 5 * F + 32, dev, "gnn_tc_pna_postf_agg_fwd", 12,
                       x, P, tab[l], g.rowptr, g.col_src, g.col_code, scal, rec.Vx,
                       conv.post_nns[0][0].weight, tower_stride_post, conv.post_nns[0][0].bias, bias_stride_post, n, F,
                       Fo, T, 1 if pp.post_layers > 1 else 0, O, A, mm)
        elif fold:
            _gemm_site(plan, pa, ("postf", l), post_rows, 5 * F, dev, "gnn_tc_pna_postf_fwd", 8,
                       x, A, scal, rec.Vx, conv.post_nns[0][0].weight, tower_stride_post,
                       conv.post_nns[0][0].bias, bias_stride_post, n, F, Fo, T, 1 if pp.post_layers > 1 else 0, O, mm)
        else:
            _post_fwd(mm, x, A, scal, conv.post_nns[0][0].weight, tower_stride_post, conv.post_nns[0][0].bias,
               bias_stride_post, n, F, Fo, T, 1 if pp.post_layers > 1 else 0, O,
               ws=_site_ws(plan, ("post", l), post_rows, 5 * F, dev) if mm >= 0 else None)
        os_ = [O]
        for k in range(1, pp.post_layers):
            O2 = torch.empty(n, H, **f32)
            last = k == pp.post_layers - 1
            for t in range(T):
                lin = conv.post_nns[t][2 * k]
                _lin_fwd(mm, O[:, t * Fo:], H, lin.weight, Fo, lin.bias, n, Fo, Fo, 0 if last else 1,
                   O2[:, t * Fo:], H, plan, ("postk", l, k, t), pa)
            os_.append(O2)
            O = O2
        rec.os = os_
        fused = None
        if not training and not save:       # inference: lin + BatchNorm(running stats) + ReLU + skip in one GEMM
            fused = _lin_bn_eval(mm, O, conv.lin, bn, rec.x_in if pp.skip_connections else None, True, plan, ("lin", l))
        if fused is not None:
            x = fused
        else:
            y = torch.empty(n, H, **f32)
            _lin_fwd(mm, O, H, conv.lin.weight, H, conv.lin.bias, n, H, H, 0, y, H, plan, ("lin_plain", l), pa)
            rec.y = y
            x, rec.stats, rec.bn_batch = _bn_forward(y, bn, training, rec.x_in if pp.skip_connections else None, True,
                                                     valid_nodes)
        if save:
            layers.append(rec)

    # global_add_pool
    pooled = torch.empty(num_graphs, H, **f32)
    _c("gnn_segment_pool_fwd", x, g.graph_ptr, num_graphs, H, pooled)

    # MLP head
    head = []
    head_keys = []
    hcur = pooled

    def head_block(lin: nn.Linear, bn: Optional[nn.BatchNorm1d]):
        nonlocal hcur
        hkey = ("head", len(head_keys))
        head_keys.append(hkey)
        if bn is not None and not training and not save:
            fused = _lin_bn_eval(mm, hcur, lin, bn, None, True, plan, hkey)
            if fused is not None:
                hcur = fused
                return
        rec = _Saved()
        rec.inp, rec.lin, rec.bn = hcur, lin, bn
        z = torch.empty(hcur.shape[0], lin.out_features, **f32)
        _lin_fwd(mm, hcur, lin.in_features, lin.weight, lin.in_features, lin.bias, hcur.shape[0],
           lin.out_features, lin.in_features, 0, z, lin.out_features, plan, hkey, pa)
        rec.z = z
        if bn is not None:
            hcur, rec.stats, rec.bn_batch = _bn_forward(z, bn, training, None, True, valid_graphs)
        else:
            hcur = z
        head.append(rec)

    for i in range(mp.num_mlp_layers):
        head_block(model.mlp[4 * i], model.mlp[4 * i + 1])
    tail = model.mlp[4 * mp.num_mlp_layers]
    head_block(tail[0], tail[1])
    head_block(tail[4], tail[5])
    head_block(tail[8], None)
    out = hcur
    if plan is not None and plan.filling and not (dev.type == "cuda" and torch.cuda.is_current_stream_capturing()):
        plan.token = plan.pending       # contents now match the weights (a capture only records the fill kernels)

    if save:
        sv.g, sv.layers, sv.head, sv.enc, sv.tab = g, layers, head, enc, tab
        sv.x_idx, sv.x_final, sv.n, sv.num_graphs = x_idx, x, n, num_graphs
        sv.layer_stride, sv.tower_stride_pre = layer_stride, tower_stride_pre
        sv.training, sv.mm, sv.fold = training, mm, fold
        sv.compact, sv.valid_nodes, sv.valid_graphs = compact, valid_nodes, valid_graphs
        sv.pa = pa
    return out, sv


def _conv_stack_tiles(model: PNAPCSAFT, x, g, tab, Vx_all, layer_stride, tower_stride_pre, plan=None):
    """All PNAConv + BatchNorm(eval) + ReLU + skip layers of a compact batch in ONE launch (gnn_tc_pna_stack_infer)."""
    pp = model.pna_params
    dev = x.device
    H = model.hidden_dim
    F, Fo, T, L = H, H // TOWERS, TOWERS, pp.propagation_depth
    n = x.shape[0]
    f32 = dict(dtype=torch.float32, device=dev)
    conv0 = model.convs[0]
    nbytes = _lib.query("gnn_tc_infer_image_bytes", L)
    if plan is not None:
        wf_all, bf_all = plan.tensor("tile_wf", (L, H, H)), plan.tensor("tile_bf", (L, H))
        images = plan.tensor("tile_images", (nbytes,), torch.uint8)
        avg = plan.tensor("tile_avg", (L,))
    else:
        wf_all, bf_all = torch.empty(L, H, H, **f32), torch.empty(L, H, **f32)
        images = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        avg = torch.empty(L, **f32)
    pre0, post0 = conv0.pre_nns[0][0], conv0.post_nns[0][0]
    ts_post = model._offset_of(conv0.post_nns[1][0].weight) - model._offset_of(post0.weight)
    bs_pre = model._offset_of(conv0.pre_nns[1][0].bias) - model._offset_of(pre0.bias)
    bs_post = model._offset_of(conv0.post_nns[1][0].bias) - model._offset_of(post0.bias)
    if plan is None or plan.filling:
        for l in range(L):
            conv, bn = model.convs[l], model.batch_norms[l].module
            _c("gnn_bn_fold_linear", conv.lin.weight, H, conv.lin.bias, bn.weight, bn.bias, bn.running_mean,
               bn.running_var, float(bn.eps), H, H, wf_all[l], bf_all[l])
        _c("gnn_tc_infer_pack", pre0.weight, tower_stride_pre, Vx_all, post0.weight, ts_post, layer_stride, wf_all, L, F,
           Fo, T, images, images.numel())
        torch.cat([c.aggr_module.avg_deg_log.reshape(1).to(torch.float32) for c in model.convs], out=avg)
    cap = int(g.num_graphs)
    tiles = torch.empty(2 * max(cap, 1), dtype=torch.int32, device=dev)
    ntiles = torch.empty(1, dtype=torch.int32, device=dev)
    _c("gnn_tc_infer_tiles", g.graph_ptr, g.rowptr, cap, g.valid[1:2], tiles, max(cap, 1), ntiles, g.err)
    x_out = torch.zeros(n, H, **f32)        # rows outside every tile (padding nodes) stay zero
    _c("gnn_tc_pna_stack_infer", x, g.rowptr, g.col_src, g.col_code, tiles, ntiles, images, tab, pre0.bias, bs_pre,
       layer_stride, post0.bias, bs_post, layer_stride, bf_all, avg, L, F, Fo, T, 1 if pp.skip_connections else 0, x_out)
    return x_out


def _backward_impl(model: PNAPCSAFT, sv, dout: torch.Tensor, grad_arena: torch.Tensor):
    """Backward of `_forward_impl`; accumulates parameter gradients into `grad_arena` (same layout
    as the parameter arena)."""
    pp = model.pna_params
    dev = dout.device
    mm = sv.mm
    H = model.hidden_dim
    F, Fo, T = H, H // TOWERS, TOWERS
    C = T * F
    L = pp.propagation_depth
    n, g = sv.n, sv.g
    f32 = dict(dtype=torch.float32, device=dev)
    arena = model._arena
    # pack-ahead continues on the forward's second stream when forward and backward are in the same capture
    pa = getattr(sv, "pa", None)
    if pa is not None and pa.capturing != torch.cuda.is_current_stream_capturing():
        pa = None

    def gview(p: torch.Tensor) -> torch.Tensor:
        off = model._offset_of(p)
        return grad_arena[off:off + p.numel()].view(p.shape)

    def bn_backward(dOut, rec_y, stats, bn, relu, use_batch, valid):
        rows, ch = rec_y.shape
        nbytes = _lib.query("gnn_bn_workspace_bytes", rows, ch)
        ws = _WS.get("bn", nbytes, dev)
        dY = torch.empty_like(rec_y)
        _c("gnn_bn_act_bwd", dOut, rec_y, stats, bn.weight, bn.bias, rows, ch, 1 if relu else 0,
           1 if use_batch else 0, dY, gview(bn.weight), gview(bn.bias), valid, ws, ws.numel())
        return dY

    # ---- head
    d = dout.contiguous()
    for rec in reversed(sv.head):
        lin = rec.lin
        if rec.bn is not None:
            d = bn_backward(d, rec.z, rec.stats, rec.bn, True, rec.bn_batch, sv.valid_graphs)
        rows = rec.inp.shape[0]
        _lin_wgrad(mm, d, lin.out_features, rec.inp, lin.in_features, rows, lin.out_features,
           lin.in_features, gview(lin.weight), lin.in_features, gview(lin.bias))
        dprev = torch.empty(rows, lin.in_features, **f32)
        _lin_dgrad(mm, d, lin.out_features, lin.weight, lin.in_features, rows, lin.out_features,
           lin.in_features, 0, dprev, lin.in_features, pa, ("head_dgrad", id(lin)))
        d = dprev
    # ---- pool
    dx = torch.empty(n, H, **f32)
    _c("gnn_segment_pool_bwd", d, g.node_graph, n, H, 0, dx)

    dtab = torch.zeros(L, NUM_BOND_CODES, C, **f32) if L > 0 else None
    sv.dVx = torch.zeros(L, T, 3, Fo, F, **f32) if (sv.fold and L > 0) else None
    for l in reversed(range(L)):
        conv, bn = model.convs[l], model.batch_norms[l].module
        rec = sv.layers[l]
        x_in = rec.x_in
        dy = bn_backward(dx, rec.y, rec.stats, bn, True, rec.bn_batch, sv.valid_nodes)
        if not pp.skip_connections:
            dx = torch.zeros(n, H, **f32)        # else: dx (the skip gradient) keeps accumulating
        # lin
        O_last = rec.os[-1]
        _lin_wgrad(mm, dy, H, O_last, H, n, H, H, gview(conv.lin.weight), H, gview(conv.lin.bias))
        dO = torch.empty(n, H, **f32)
        _lin_dgrad(mm, dy, H, conv.lin.weight, H, n, H, H, 0, dO, H, pa, ("lin_dgrad", l))
        # extra post layers
        for k in reversed(range(1, pp.post_layers)):
            O_in = rec.os[k - 1]
            dprev = torch.empty(n, H, **f32)
            for t in range(T):
                lin = conv.post_nns[t][2 * k]
                _lin_wgrad(mm, dO[:, t * Fo:], H, O_in[:, t * Fo:], H, n, Fo, Fo, gview(lin.weight), Fo,
                   gview(lin.bias))
                _lin_dgrad(mm, dO[:, t * Fo:], H, lin.weight, Fo, n, Fo, Fo, 0, dprev[:, t * Fo:], H, pa,
                           ("postk_dgrad", l, k, t))
            _c("gnn_relu_bwd", dprev, O_in, n * H)
            dO = dprev
        ts_post, bs_post = rec.post_strides
        w0, b0 = conv.post_nns[0][0].weight, conv.post_nns[0][0].bias
        if sv.fold:
            _backward_layer_folded(model, sv, rec, l, dO, dx, dtab, gview, pa)
            continue
        _post_wgrad(mm, dO, x_in, rec.A, rec.scal, n, F, Fo, T, gview(w0), ts_post, gview(b0), bs_post)
        gA = torch.empty(n, T * 4 * F, **f32)
        fused_dx = mm >= 0        # tensor-core path: every node-feature gradient of the layer in ONE GEMM (below)
        _post_dgrad(mm, dO, rec.scal, w0, ts_post, n, F, Fo, T, None if fused_dx else dx, gA, pa=pa,
                    key=("post_dgrad", l))
        # aggregation + message
        dP = torch.empty(n, 2 * C, **f32)
        if pp.pre_layers == 1:
            # default: per-slot gradients dq, summed per source node by gnn_scatter_src_bwd; dq = None makes the
            # kernel accumulate them atomically into dP[:, C:] instead
            dq = torch.empty(max(g.ep, 1), C, **f32) if DETERMINISTIC_SCATTER else None
            nb = _lib.query("gnn_pna_aggregate_bwd_workspace_bytes", T, F)
            zws = _WS.get_zeroed("aggbwd", nb, dev)         # kept all-zero by the kernel pair itself
            _c("gnn_pna_aggregate_bwd", gA, rec.P, sv.tab[l], rec.A, g.rowptr, g.col_src, g.col_code, n, T, F, dP,
               dq, dtab[l], zws, zws.numel())
        else:
            hs = rec.hs
            dm = torch.empty(max(g.ep, 1), C, **f32)
            _c("gnn_pna_aggregate_msg_bwd", gA, hs[-1], rec.A, g.rowptr, n, T, F, dm)
            for k in reversed(range(1, pp.pre_layers)):
                h_in = hs[k - 1]
                dprev = torch.empty(max(g.ep, 1), C, **f32)
                for t in range(T):
                    lin = conv.pre_nns[t][2 * k]
                    _lin_wgrad(mm, dm[:, t * F:], C, h_in[:, t * F:], C, g.ep, F, F, gview(lin.weight), F,
                       gview(lin.bias))
                    _lin_dgrad(mm, dm[:, t * F:], C, lin.weight, F, g.ep, F, F, 0, dprev[:, t * F:], C)
                _c("gnn_relu_bwd", dprev, h_in, g.ep * C)
                dm = dprev
            dq = dm
            _c("gnn_edge_hidden_bwd", dq, g.rowptr, g.col_code, n, C, dP, dtab[l])
        if dq is not None:
            _c("gnn_scatter_src_bwd", dq, g.srcptr, g.src_slot, n, C, dP)
        # projections
        if fused_dx:
            pre0 = conv.pre_nns[0][0]
            ts_pre, bs_pre = rec.pre_strides
            _c("gnn_tc_pna_proj_wgrad", dP, x_in, n, F, T, gview(pre0.weight), ts_pre, gview(pre0.bias), bs_pre, mm)
            _gemm_site(None, pa, ("dx_dgrad", l), F, 2 * C + T * Fo, dev, "gnn_tc_pna_dx_dgrad", 6,
                       dP, dO, pre0.weight, ts_pre, w0, ts_post, n, F, Fo, T, 1, dx, mm)
        else:
            for t in range(T):
                lin = conv.pre_nns[t][0]
                gw = gview(lin.weight)
                _lin_wgrad(mm, dP[:, t * F:], 2 * C, x_in, H, n, F, F, gw, 3 * F, gview(lin.bias))
                _lin_wgrad(mm, dP[:, C + t * F:], 2 * C, x_in, H, n, F, F, gw[:, F:], 3 * F, None)
                _lin_dgrad(mm, dP[:, t * F:], 2 * C, lin.weight, 3 * F, n, F, F, 1, dx, H)
                _lin_dgrad(mm, dP[:, C + t * F:], 2 * C, lin.weight[:, F:], 3 * F, n, F, F, 1, dx, H)

    # ---- encoders: three small, latency-bound jobs that write disjoint gradient regions (atom tables | post weights +
    # destination third of the pre weights | bond tables, edge encoders, bond third of the pre weights).  None of them
    # fills the GPU, so the last two run on side streams beside the AtomEncoder backward (parallel branches of a
    # captured step); with per-call timing on they stay in line.
    tails = []
    if L > 0:
        conv0 = model.convs[0]
        if sv.fold:     # dVx of all layers -> post_nns / pre_nns (destination third) weight gradients, one launch
            w0, wp0 = conv0.post_nns[0][0].weight, conv0.pre_nns[0][0].weight
            tsp = model._offset_of(conv0.post_nns[1][0].weight) - model._offset_of(w0)
            tails.append(lambda: _c("gnn_pna_fold_bwd", sv.dVx, w0, tsp, wp0, sv.tower_stride_pre, sv.layer_stride, L, F,
                                    Fo, T, gview(w0), gview(wp0)))
        denc = torch.empty(L, NUM_BOND_CODES, F, **f32)
        tails.append(lambda: _c(
            "gnn_bond_table_bwd", dtab, model.edge_embed.bond_embedding_list[0].weight, conv0.edge_encoder.weight,
            conv0.pre_nns[0][0].weight, sv.tower_stride_pre, sv.layer_stride, L, sv.enc, H, T, denc,
            gview(model.edge_embed.bond_embedding_list[0].weight), gview(conv0.edge_encoder.weight),
            gview(conv0.edge_encoder.bias), gview(conv0.pre_nns[0][0].weight)))
    joins = []
    if dev.type == "cuda" and _PROFILER is None and tails:
        fork = torch.cuda.Event()
        fork.record()
        for i, job in enumerate(tails):
            st = _side_stream(dev, i)
            st.wait_event(fork)
            with torch.cuda.stream(st):
                job()
                ev = torch.cuda.Event()
                ev.record()
            joins.append(ev)
        tails = []
    atom_tab = model.node_embed.atom_embedding_list[0].weight
    _c("gnn_atom_embed_u8_bwd" if sv.compact else "gnn_atom_embed_bwd", sv.x_idx, n, dx, model._consts(dev),
       sum(ATOM_FEATURE_DIMS), H, gview(atom_tab))
    for job in tails:
        job()
    for ev in joins:
        torch.cuda.current_stream().wait_event(ev)
    del arena


def _backward_layer_folded(model, sv, rec, l, dO, dx, dtab, gview, pa=None):
    """Message / aggregate / post_nns[.][0] part of one conv layer's backward with the destination projection
    folded into the post weights (csrc/fold.cu): accumulates into dx and the parameter gradients."""
    conv = model.convs[l]
    mm, g, n = sv.mm, sv.g, sv.n
    dev = dO.device
    H = model.hidden_dim
    F, Fo, T = H, H // TOWERS, TOWERS
    C = T * F
    f32 = dict(dtype=torch.float32, device=dev)
    ts_post, bs_post = rec.post_strides
    ts_pre, bs_pre = rec.pre_strides
    w0, b0 = conv.post_nns[0][0].weight, conv.post_nns[0][0].bias
    pre0 = conv.pre_nns[0][0]
    dVx = sv.dVx[l]
    _c("gnn_tc_pna_postf_wgrad", dO, rec.x_in, rec.A, rec.scal, n, F, Fo, T, gview(w0), ts_post, gview(b0), bs_post,
       dVx, mm)
    gA = torch.empty(n, T * 4 * F, **f32)
    _post_dgrad(mm, dO, rec.scal, w0, ts_post, n, F, Fo, T, None, gA, pa=pa, key=("post_dgrad", l))
    dPs = torch.empty(n, C, **f32)
    dq = torch.empty(max(g.ep, 1), C, **f32) if DETERMINISTIC_SCATTER else None
    nb = _lib.query("gnn_pna_aggregate_bwd_workspace_bytes", T, F)
    zws = _WS.get_zeroed("aggbwd", nb, dev)
    _c("gnn_pna_aggregate_src_bwd", gA, rec.P, sv.tab[l], rec.A, g.rowptr, g.col_src, g.col_code, n, T, F, dPs, dq,
       dtab[l], zws, zws.numel())
    if dq is not None:
        _c("gnn_slot_sum_by_source", dq, g.srcptr, g.src_slot, n, C, dPs, C)
    _c("gnn_tc_pna_projsrc_wgrad", dPs, rec.x_in, n, F, T, gview(pre0.weight), ts_pre, gview(pre0.bias), bs_pre, mm)
    _gemm_site(None, pa, ("dxf_dgrad", l), F, C + 3 * T * Fo, dev, "gnn_tc_pna_dxf_dgrad", 6,
               dPs, dO, rec.scal, pre0.weight, ts_pre, rec.Vx, n, F, Fo, T, 1, dx, mm)


class _PNAFunction(torch.autograd.Function):
    """One autograd node for the whole network: forward/backward are the CUDA kernel chains."""

    @staticmethod
    def forward(ctx, model, x_idx, edge_index, edge_attr, batch, num_graphs, *params):
        out, sv = _forward_impl(model, x_idx, edge_index, edge_attr, batch, num_graphs, save=True)
        ctx.model, ctx.sv = model, sv
        ctx.n_params = len(params)
        return out

    @staticmethod
    def backward(ctx, dout):
        model, sv = ctx.model, ctx.sv
        if sv is None:
            raise RuntimeError("gnnepcsaft_b200: backward was already run for this forward (the saved kernel state "
                               "is freed after the first backward; retain_graph is not supported)")
        if not model._arena_valid():
            raise RuntimeError("gnnepcsaft_b200: parameters were moved between forward and backward")
        with _device_guard(dout.device):
            grad_arena = torch.zeros_like(model._arena)
            _backward_impl(model, sv, dout, grad_arena)
        grads = []
        for p, off in zip(model._param_list(), model._arena_offsets):
            grads.append(grad_arena[off:off + p.numel()].view(p.shape) if p.requires_grad else None)
        ctx.sv = None
        return (None, None, None, None, None, None, *grads)


def mape(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """torchmetrics.functional.mean_absolute_percentage_error (reference models.py:16,194)."""
    return torch.sum((pred - target).abs() / target.abs().clamp(min=1.17e-06)) / target.numel()


try:  # the reference's wrapper is a LightningModule (models.py:138); subclass it whenever Lightning is importable
    import lightning as _L  # type: ignore
    _LightningBase = _L.LightningModule
except ImportError:  # pragma: no cover - depends on the environment
    try:
        import pytorch_lightning as _L  # type: ignore
        _LightningBase = _L.LightningModule
    except ImportError:
        _L = None

        class _LightningBase(nn.Module):
            """Lightning is not installed: same attribute surface (`log`, `log_dict`, hooks the reference's
            train.py never overrides) on a plain module, so `TrainEngine` / a hand-written loop can drive it."""

            def log(self, *args, **kwargs):
                return None

            def log_dict(self, *args, **kwargs):
                return None


class PNApcsaftL(_LightningBase):
    """Graph neural network to predict ePCSAFT parameters with pytorch lightning (reference models.py:138-251).

    Holds `self.model = PNAPCSAFT(...)` (checkpoint keys carry the `model.` prefix), builds the same optimizer /
    per-step cosine-restart schedule and computes the same training loss, so `L.Trainer.fit(model, loader)` of
    reference train.py:129-201 drives the CUDA kernels through autograd.  `validation_step` / `test_step` hand the
    predicted parameters to `property_metrics` - the ePC-SAFT density / vapour-pressure evaluation (feos, reference
    models.py:204-248) is CPU thermodynamics outside this path; plug the reference's functions in there."""

    def __init__(self, pna_params: PnaconvsParams, mlp_params: ReadoutMLPParams, config):
        super().__init__()
        self.config = config
        self.model = PNAPCSAFT(config.hidden_dim, pna_params=pna_params, mlp_params=mlp_params)
        self.property_metrics = None        # optional callable(pred_para float64 cpu [P], graphs) -> dict of floats

    def forward(self, data):
        """Forward pass of the model"""
        return self.model(data)

    def configure_optimizers(self):
        if self.config.optimizer == "adam":
            opt = torch.optim.AdamW(self.parameters(), lr=self.config.learning_rate,
                                    weight_decay=self.config.weight_decay, amsgrad=True, eps=1e-5)
        elif self.config.optimizer == "sgd":
            opt = torch.optim.SGD(self.parameters(), lr=self.config.learning_rate, momentum=self.config.momentum,
                                  weight_decay=self.config.weight_decay, nesterov=True)
        else:
            raise ValueError(f"Unsupported optimizer: {self.config.optimizer}.")
        sched = torch.optim.lr_scheduler.CosineAnnealingWarmRestarts(opt, self.config.warmup_steps)
        return {"optimizer": opt, "lr_scheduler": {"scheduler": sched, "interval": "step", "frequency": 1}}

    def training_step(self, graphs, batch_idx=0):
        target = graphs.para.view(-1, self.config.num_para)
        pred = self(graphs)
        loss_mape = mape(pred, target)
        self.log("train_mape", loss_mape, on_step=True, batch_size=target.shape[0], sync_dist=True)
        return loss_mape

    def validation_step(self, graphs, batch_idx=0):
        pred_para = self(graphs).squeeze().to(torch.float64).cpu()
        metrics = {}
        if self.property_metrics is not None:
            metrics = dict(self.property_metrics(pred_para, graphs))
        if metrics:
            self.log_dict(metrics, on_step=True, batch_size=1, sync_dist=True)
        return metrics

    def test_step(self, graphs, batch_idx=0):
        return self.validation_step(graphs, batch_idx)
