"""ORACLE (test infrastructure, not product code) — per-entry-point CPU restatement of the
C ABI in include/gnnepcsaft_b200.h, written with plain torch ops.

Two uses, both in tests/ only:
  * kernel-level checker: every `gnn_*` CUDA entry point is compared with the function of the
    same name here on identical inputs (`-m gpu` tests);
  * host-logic checker: `install(monkeypatch)` swaps the ctypes dispatcher for this module so
    the Python orchestration (forward/backward chains, strides, gradient accumulation) can be
    compared with oracle/pna_oracle.py on a machine without a GPU (`-m "not gpu"` tests).
The product never imports this file; without the CUDA library the product raises.

Arguments mirror the C signatures: tensors stand for the pointers (a sliced view stands for
`ptr + offset`, with the leading dimension passed explicitly as in C), the trailing `stream`
argument is ignored.  PARITY UNPINNED — see oracle/pna_oracle.py.
"""

from __future__ import annotations

import math

import torch

NUM_CODES = 60


def _flat(t: torch.Tensor, numel: int, extra_offset: int = 0) -> torch.Tensor:
    """1-D window of `numel` elements starting at t's first element (+offset) in its storage."""
    return t.as_strided((numel,), (1,), t.storage_offset() + extra_offset)


def _mat(t: torch.Tensor, rows: int, cols: int, ld: int, extra_offset: int = 0) -> torch.Tensor:
    """rows x cols matrix with leading dimension ld starting at t's first element (+offset)."""
    return t.as_strided((rows, cols), (ld, 1), t.storage_offset() + extra_offset)


# ------------------------------------------------------------------ graph structure
def gnn_build_csr_workspace_bytes(n, e):
    return 256


def gnn_build_csr(edge_index, edge_attr, n, e, self_loops, rowptr, col_src, col_code, col_eid, srcptr, src_slot,
                  err_flag, workspace, workspace_bytes, stream=None):
    src = edge_index[0].tolist() if e else []
    dst = edge_index[1].tolist() if e else []
    codes = [int(a[0]) * 12 + int(a[1]) * 2 + int(a[2]) for a in edge_attr.tolist()] if e else []
    if self_loops:
        src += list(range(n))
        dst += list(range(n))
        codes += [0] * n
    order = sorted(range(len(src)), key=lambda k: (dst[k], k))
    counts = [0] * (n + 1)
    for d in dst:
        counts[d + 1] += 1
    rp = torch.tensor(counts).cumsum(0).to(torch.int32)
    rowptr[: n + 1] = rp
    slot_of_edge = {}
    for s, k in enumerate(order):
        col_src[s] = src[k]
        col_code[s] = codes[k]
        col_eid[s] = k
        slot_of_edge[k] = s
    out_counts = [0] * (n + 1)
    for s_ in src:
        out_counts[s_ + 1] += 1
    sp = torch.tensor(out_counts).cumsum(0).to(torch.int32)
    srcptr[: n + 1] = sp
    by_src = sorted(range(len(src)), key=lambda k: (src[k], slot_of_edge[k]))
    for pos, k in enumerate(by_src):
        src_slot[pos] = slot_of_edge[k]
    if err_flag is not None:
        err_flag.zero_()


def gnn_graph_ptr_from_batch(batch, n, b, node_graph, graph_ptr, err_flag=None, stream=None):
    if err_flag is not None and n > 0:
        bv = batch[:n]
        if bool(((bv < 0) | (bv >= b)).any()):
            err_flag |= 4
        elif n > 1 and bool((bv[1:] < bv[:-1]).any()):
            err_flag |= 8
    batch = batch[:n].clamp(0, b - 1)
    counts = torch.bincount(batch[:n], minlength=b)
    graph_ptr[0] = 0
    graph_ptr[1: b + 1] = counts.cumsum(0).to(torch.int32)
    node_graph[:n] = batch[:n].to(torch.int32)


def gnn_graph_ptr_single(n, graph_ptr, stream=None):
    graph_ptr[0] = 0
    graph_ptr[1] = n


def gnn_collate_workspace_bytes(nb):
    return 256


def gnn_collate_offsets(ids, nb, node_slices, edge_slices, out_ptr, out_eptr, workspace, workspace_bytes,
                        stream=None):
    ncnt = node_slices[ids + 1] - node_slices[ids]
    ecnt = edge_slices[ids + 1] - edge_slices[ids]
    out_ptr[0] = 0
    out_eptr[0] = 0
    out_ptr[1: nb + 1] = ncnt.cumsum(0)
    out_eptr[1: nb + 1] = ecnt.cumsum(0)


def gnn_collate_gather(ids, nb, node_slices, edge_slices, sx, sei, store_e, sea, optr, oeptr, tot_n, tot_e, x, ei,
                       ea, batch, stream=None):
    for b in range(nb):
        g = int(ids[b])
        n0, n1 = int(node_slices[g]), int(node_slices[g + 1])
        e0, e1 = int(edge_slices[g]), int(edge_slices[g + 1])
        on, oe = int(optr[b]), int(oeptr[b])
        x[on: on + n1 - n0] = sx[n0:n1]
        batch[on: on + n1 - n0] = b
        ea[oe: oe + e1 - e0] = sea[e0:e1]
        ei[:, oe: oe + e1 - e0] = sei[:, e0:e1] + on


def gnn_compact_collate_workspace_bytes(nb):
    return 256


def gnn_compact_collate(ids, nb, cap, num_store_graphs, node_off, slot_off, sx, s_rowptr, s_col_src, s_col_code, s_para,
                        num_para, pad_nodes, pad_slots, own_surplus, x, rowptr, col_src, col_code, node_graph, graph_ptr, para,
                        valid,
                        err_flag=None, workspace=None, workspace_bytes=0, stream=None):
    if err_flag is not None:
        err_flag.zero_()
    on, os_ = 0, 0
    sizes = [(int(node_off[g + 1] - node_off[g]), int(slot_off[g + 1] - slot_off[g])) for g in ids[:nb].tolist()]
    tot_n, tot_s = sum(a for a, _ in sizes), sum(b for _, b in sizes)
    if tot_n > pad_nodes or tot_s > pad_slots:
        if err_flag is not None:
            err_flag |= 1
        return
    for i, g in enumerate(ids[:nb].tolist()):
        n0, s0 = int(node_off[g]), int(slot_off[g])
        cn, cs = sizes[i]
        x[on:on + cn] = sx[n0:n0 + cn]
        rowptr[on:on + cn] = s_rowptr[n0:n0 + cn] - s0 + os_
        node_graph[on:on + cn] = i
        col_src[os_:os_ + cs] = s_col_src[s0:s0 + cs] - n0 + on
        col_code[os_:os_ + cs] = s_col_code[s0:s0 + cs]
        graph_ptr[i] = on
        if para is not None:
            para[i * num_para:(i + 1) * num_para] = s_para.reshape(-1)[g * num_para:(g + 1) * num_para]
        on += cn
        os_ += cs
    nd, d = pad_nodes - on, pad_slots - os_
    rowptr[pad_nodes] = pad_slots if own_surplus else os_
    graph_ptr[nb:cap + 1] = on          # graphs nb .. cap-1 are empty, graph `cap` is the padding graph
    graph_ptr[cap + 1] = pad_nodes
    valid[0], valid[1], valid[2], valid[3] = on, nb, os_, 0
    if para is not None:
        para[nb * num_para:(cap + 1) * num_para] = 1.0
    if nd == 0:
        if own_surplus and d != 0 and err_flag is not None:
            err_flag |= 1
        return
    base, rem = (d // nd, d % nd) if own_surplus else (0, 0)
    for k in range(nd):
        node, start, cnt = on + k, os_ + k * base + min(k, rem), base + (1 if k < rem else 0)
        x[node] = 0
        rowptr[node] = start
        node_graph[node] = cap
        col_src[start:start + cnt] = node
        col_code[start:start + cnt] = 0


def gnn_atom_features_to_u8(x, n, out, err_flag=None, stream=None):
    v = x[:n]
    if err_flag is not None and bool(((v < 0) | (v > 255)).any()):
        err_flag |= 16
    out[:n] = v.clamp(0, 255).to(torch.uint8)


# ------------------------------------------------------------------ encoders
def gnn_atom_embed_fwd(x, n, tables, row_off, total_rows, hidden, out, err_flag=None, stream=None):
    tab = _mat(tables, total_rows, hidden, hidden)
    acc = torch.zeros(n, hidden, dtype=tab.dtype)
    for k in range(9):
        hi = (int(row_off[k + 1]) if k + 1 < 9 else total_rows) - int(row_off[k])
        v = x[:n, k]
        if err_flag is not None and bool(((v < 0) | (v >= hi)).any()):
            err_flag |= 16
        acc = acc + tab[row_off[k].long() + v.clamp(0, hi - 1)]
    out[:n] = acc


def gnn_atom_embed_u8_fwd(x, n, tables, row_off, total_rows, hidden, out, err_flag=None, stream=None):
    gnn_atom_embed_fwd(x.long(), n, tables, row_off, total_rows, hidden, out, err_flag)


def gnn_atom_embed_u8_bwd(x, n, dout, row_off, total_rows, hidden, dtables, stream=None):
    gnn_atom_embed_bwd(x.long(), n, dout, row_off, total_rows, hidden, dtables)


def gnn_atom_embed_bwd(x, n, dout, row_off, total_rows, hidden, dtables, stream=None):
    dt = _mat(dtables, total_rows, hidden, hidden)
    for k in range(9):
        dt.index_add_(0, row_off[k].long() + x[:n, k], dout[:n])


def _bond_rows(code):
    return code // 12, 5 + (code % 12) // 2, 11 + (code & 1)


def _bond_emb(bond_tables, hidden):
    tab = _mat(bond_tables, 13, hidden, hidden)
    return torch.stack([tab[r0] + tab[r1] + tab[r2] for r0, r1, r2 in map(_bond_rows, range(NUM_CODES))])


def gnn_bond_table_fwd(bond_tables, wee, bee, wpre, wpre_tower_stride, layer_stride, num_layers, hidden, towers,
                       enc, tab, stream=None):
    f = hidden
    emb = _bond_emb(bond_tables, hidden)
    for l in range(num_layers):
        w = _mat(wee, f, hidden, hidden, l * layer_stride)
        b = _flat(bee, f, l * layer_stride)
        e = emb @ w.t() + b
        enc[l] = e
        for t in range(towers):
            wp = _mat(wpre, f, 3 * f, 3 * f, l * layer_stride + t * wpre_tower_stride)
            tab[l, :, t * f:(t + 1) * f] = e @ wp[:, 2 * f:].t()


def gnn_bond_table_bwd(dtab, bond_tables, wee, wpre, wpre_tower_stride, layer_stride, num_layers, enc, hidden,
                       towers, denc_ws, dbond_tables, dwee, dbee, dwpre, stream=None):
    f = hidden
    emb = _bond_emb(bond_tables, hidden)
    dbt = _mat(dbond_tables, 13, hidden, hidden)
    for l in range(num_layers):
        denc = torch.zeros(NUM_CODES, f, dtype=dtab.dtype)
        for t in range(towers):
            off = l * layer_stride + t * wpre_tower_stride
            wp = _mat(wpre, f, 3 * f, 3 * f, off)
            d = dtab[l, :, t * f:(t + 1) * f]
            denc += d @ wp[:, 2 * f:]
            _mat(dwpre, f, 3 * f, 3 * f, off)[:, 2 * f:] += d.t() @ enc[l]
        denc_ws[l] = denc
        w = _mat(wee, f, hidden, hidden, l * layer_stride)
        _mat(dwee, f, hidden, hidden, l * layer_stride).add_(denc.t() @ emb)
        _flat(dbee, f, l * layer_stride).add_(denc.sum(0))
        demb = denc @ w
        for c in range(NUM_CODES):
            for r in _bond_rows(c):
                dbt[r] += demb[c]


# ------------------------------------------------------------------ aggregation
def _std_from_var(var):
    s = var.clamp(min=1e-5).sqrt()
    return s.masked_fill(s <= math.sqrt(1e-5), 0.0)


def _segments(rowptr, n):
    rp = rowptr[: n + 1].tolist()
    return [(rp[i], rp[i + 1]) for i in range(n)]


def _aggregate(q_rows, shift):
    """q_rows [d, C] messages (minus shift), shift [C] added to mean/min/max."""
    mean = q_rows.mean(0)
    var = ((q_rows - mean) ** 2).mean(0)
    return mean + shift, q_rows.min(0).values + shift, q_rows.max(0).values + shift, _std_from_var(var)


def _store_A(A, i, towers, f, mean, mn, mx, sd):
    row = A[i].view(towers, 4, f)
    row[:, 0] = mean.view(towers, f)
    row[:, 1] = mn.view(towers, f)
    row[:, 2] = mx.view(towers, f)
    row[:, 3] = sd.view(towers, f)


def gnn_pna_aggregate_fwd(P, tab, rowptr, col_src, col_code, n, towers, f, A, stream=None):
    c = towers * f
    for i, (b, e) in enumerate(_segments(rowptr, n)):
        if e == b:
            A[i] = 0
            continue
        q = P[col_src[b:e].long(), c:] + tab[col_code[b:e].long()]
        _store_A(A, i, towers, f, *_aggregate(q, P[i, :c]))


def _agg_grad(q, ga_row, sd, towers, f):
    """q [d,C] -> dq [d,C] and the (g_mean+g_min+g_max) row."""
    d = q.shape[0]
    g = ga_row.view(towers, 4, f)
    g_mean, g_min, g_max, g_std = (g[:, k].reshape(-1) for k in range(4))
    mean = q.mean(0)
    is_min = (q == q.min(0).values).to(q.dtype)
    is_max = (q == q.max(0).values).to(q.dtype)
    dq = (g_mean / d).expand_as(q).clone()
    dq += is_min * (g_min / is_min.sum(0))
    dq += is_max * (g_max / is_max.sum(0))
    coef = torch.where(sd > 0, g_std / (d * sd.clamp(min=1e-30)), torch.zeros_like(sd))
    dq += coef * (q - mean)
    return dq, g_mean + g_min + g_max


def gnn_pna_aggregate_bwd_workspace_bytes(towers, f):
    return 256


def gnn_pna_aggregate_bwd(gA, P, tab, A, rowptr, col_src, col_code, n, towers, f, dP, dq, dtab_accum,
                          workspace=None, workspace_bytes=0, stream=None):
    c = towers * f
    if dq is None:
        dP[:n, c:2 * c] = 0
    for i, (b, e) in enumerate(_segments(rowptr, n)):
        if e == b:
            dP[i, :c] = 0
            continue
        q = P[col_src[b:e].long(), c:] + tab[col_code[b:e].long()]
        sd = A[i].view(towers, 4, f)[:, 3].reshape(-1)
        d, dpd = _agg_grad(q, gA[i], sd, towers, f)
        if dq is None:
            dP[:, c:2 * c].index_add_(0, col_src[b:e].long(), d)
        else:
            dq[b:e] = d
        dP[i, :c] = dpd
        dtab_accum.index_add_(0, col_code[b:e].long(), d)


def gnn_scatter_src_bwd(dq, srcptr, src_slot, n, c, dP, stream=None):
    for j, (b, e) in enumerate(_segments(srcptr, n)):
        dP[j, c:2 * c] = dq[src_slot[b:e].long()].sum(0) if e > b else 0


# ---- folded destination projection (gnn_pna_fold_fwd and friends)
def _wsum(wp, f, s):
    """(Wmean + Wmin + Wmax) of scaler group s: [F_out, F]."""
    base = f + s * 4 * f
    return wp[:, base:base + f] + wp[:, base + f:base + 2 * f] + wp[:, base + 2 * f:base + 3 * f]


def gnn_pna_fold_fwd(wpost, wpost_ts, wpre, wpre_ts, layer_stride, num_layers, f, f_out, towers, Vx, stream=None):
    v = Vx.view(num_layers, towers, 3, f_out, f)
    for l in range(num_layers):
        for t in range(towers):
            wp = _mat(wpost, f_out, 13 * f, 13 * f, l * layer_stride + t * wpost_ts)
            wd = _mat(wpre, f, 3 * f, 3 * f, l * layer_stride + t * wpre_ts)[:, :f]
            for s in range(3):
                v[l, t, s] = _wsum(wp, f, s) @ wd + (wp[:, :f] if s == 0 else 0)


def gnn_pna_fold_bwd(dVx, wpost, wpost_ts, wpre, wpre_ts, layer_stride, num_layers, f, f_out, towers, dwpost, dwpre,
                     stream=None):
    dv = dVx.view(num_layers, towers, 3, f_out, f)
    for l in range(num_layers):
        for t in range(towers):
            wp = _mat(wpost, f_out, 13 * f, 13 * f, l * layer_stride + t * wpost_ts)
            wd = _mat(wpre, f, 3 * f, 3 * f, l * layer_stride + t * wpre_ts)[:, :f]
            dwp = _mat(dwpost, f_out, 13 * f, 13 * f, l * layer_stride + t * wpost_ts)
            dwd = _mat(dwpre, f, 3 * f, 3 * f, l * layer_stride + t * wpre_ts)
            dwp[:, :f] += dv[l, t, 0]
            for s in range(3):
                g = dv[l, t, s] @ wd.t()
                base = f + s * 4 * f
                for a in range(3):
                    dwp[:, base + a * f:base + (a + 1) * f] += g
                dwd[:, :f] += _wsum(wp, f, s).t() @ dv[l, t, s]


def gnn_pna_aggregate_src_fwd(Psrc, tab, rowptr, col_src, col_code, n, towers, f, A, stream=None):
    for i, (b, e) in enumerate(_segments(rowptr, n)):
        if e == b:
            A[i] = 0
            continue
        q = Psrc[col_src[b:e].long()] + tab[col_code[b:e].long()]
        _store_A(A, i, towers, f, *_aggregate(q, torch.zeros_like(q[0])))


def gnn_pna_aggregate_src_bwd(gA, Psrc, tab, A, rowptr, col_src, col_code, n, towers, f, dPsrc, dq, dtab_accum,
                              workspace=None, workspace_bytes=0, stream=None):
    if dq is None:
        dPsrc[:n] = 0
    for i, (b, e) in enumerate(_segments(rowptr, n)):
        if e == b:
            continue
        q = Psrc[col_src[b:e].long()] + tab[col_code[b:e].long()]
        sd = A[i].view(towers, 4, f)[:, 3].reshape(-1)
        d, _ = _agg_grad(q, gA[i], sd, towers, f)
        if dq is None:
            dPsrc.index_add_(0, col_src[b:e].long(), d)
        else:
            dq[b:e] = d
        dtab_accum.index_add_(0, col_code[b:e].long(), d)


def gnn_slot_sum_by_source(dq, srcptr, src_slot, n, c, out, ldo, stream=None):
    o = _mat(out, n, c, ldo)
    for j, (b, e) in enumerate(_segments(srcptr, n)):
        o[j] = dq[src_slot[b:e].long()].sum(0) if e > b else 0


def gnn_edge_hidden_fwd(P, tab, rowptr, col_src, col_code, n, c, relu, hidden_out, stream=None):
    for i, (b, e) in enumerate(_segments(rowptr, n)):
        if e > b:
            h = P[i, :c] + P[col_src[b:e].long(), c:] + tab[col_code[b:e].long()]
            hidden_out[b:e] = h.clamp(min=0) if relu else h


def gnn_edge_hidden_bwd(dh, rowptr, col_code, n, c, dP, dtab_accum, stream=None):
    for i, (b, e) in enumerate(_segments(rowptr, n)):
        dP[i, :c] = dh[b:e].sum(0) if e > b else 0
        if e > b:
            dtab_accum.index_add_(0, col_code[b:e].long(), dh[b:e])


def gnn_pna_aggregate_msg_fwd(m, rowptr, n, towers, f, A, stream=None):
    for i, (b, e) in enumerate(_segments(rowptr, n)):
        if e == b:
            A[i] = 0
            continue
        _store_A(A, i, towers, f, *_aggregate(m[b:e], torch.zeros_like(m[0])))


def gnn_pna_aggregate_msg_bwd(gA, m, A, rowptr, n, towers, f, dm, stream=None):
    for i, (b, e) in enumerate(_segments(rowptr, n)):
        if e > b:
            sd = A[i].view(towers, 4, f)[:, 3].reshape(-1)
            dm[b:e] = _agg_grad(m[b:e], gA[i], sd, towers, f)[0]


# ------------------------------------------------------------------ dense stages
def gnn_linear_fwd(X, ldx, W, ldw, bias, M, N, K, act, Y, ldy, stream=None):
    y = _mat(X, M, K, ldx) @ _mat(W, N, K, ldw).t()
    if bias is not None:
        y = y + _flat(bias, N)
    _mat(Y, M, N, ldy).copy_(y.clamp(min=0) if act else y)


def gnn_linear_dgrad(dY, lddy, W, ldw, M, N, K, accumulate, dX, lddx, stream=None):
    d = _mat(dY, M, N, lddy) @ _mat(W, N, K, ldw)
    out = _mat(dX, M, K, lddx)
    out.copy_(out + d if accumulate else d)


def gnn_linear_wgrad(dY, lddy, X, ldx, M, N, K, dW, lddw, db, stream=None):
    dy = _mat(dY, M, N, lddy)
    _mat(dW, N, K, lddw).add_(dy.t() @ _mat(X, M, K, ldx))
    if db is not None:
        _flat(db, N).add_(dy.sum(0))


def gnn_relu_bwd(dY, Y_act, numel, stream=None):
    d, y = _flat(dY, numel), _flat(Y_act, numel)
    d.copy_(torch.where(y > 0, d, torch.zeros_like(d)))


def gnn_pna_scalers(rowptr, avg_deg_log, n, scal, stream=None):
    deg = (rowptr[1: n + 1] - rowptr[:n]).to(scal.dtype)
    delta = avg_deg_log.reshape(-1)[0].to(scal.dtype)
    scal[:n, 0] = torch.log(deg + 1) / delta
    scal[:n, 1] = delta / torch.log(deg.clamp(min=1) + 1)


def _post_cat(x, A, scal, n, f, towers, t):
    a = A[:n].view(n, towers, 4 * f)[:, t]
    return torch.cat([x[:n, :f], a, a * scal[:n, 0:1], a * scal[:n, 1:2]], dim=1)


def gnn_pna_post_fwd(x, A, scal, wpost, wts, bpost, bts, n, f, f_out, towers, act, O, stream=None):
    for t in range(towers):
        w = _mat(wpost, f_out, 13 * f, 13 * f, t * wts)
        b = _flat(bpost, f_out, t * bts)
        o = _post_cat(x, A, scal, n, f, towers, t) @ w.t() + b
        O[:n, t * f_out:(t + 1) * f_out] = o.clamp(min=0) if act else o


def gnn_pna_post_dgrad(dO, scal, wpost, wts, n, f, f_out, towers, dx, gA, stream=None):
    for t in range(towers):
        w = _mat(wpost, f_out, 13 * f, 13 * f, t * wts)
        ds = dO[:n, t * f_out:(t + 1) * f_out] @ w           # [n, 13f]
        dx[:n] += ds[:, :f]
        ga = ds[:, f:5 * f] + scal[:n, 0:1] * ds[:, 5 * f:9 * f] + scal[:n, 1:2] * ds[:, 9 * f:]
        gA[:n].view(n, towers, 4 * f)[:, t] = ga


def gnn_pna_post_wgrad(dO, x, A, scal, n, f, f_out, towers, dwpost, wts, dbpost, bts, stream=None):
    for t in range(towers):
        d = dO[:n, t * f_out:(t + 1) * f_out]
        _mat(dwpost, f_out, 13 * f, 13 * f, t * wts).add_(d.t() @ _post_cat(x, A, scal, n, f, towers, t))
        _flat(dbpost, f_out, t * bts).add_(d.sum(0))


# tensor-core entry points: same semantics (precision only changes the rounding of the operands)
def gnn_tc_linear_fwd(X, ldx, W, ldw, bias, M, N, K, act, Y, ldy, precision, pack_ws=None, pack_ws_bytes=0,
                      stream=None):
    gnn_linear_fwd(X, ldx, W, ldw, bias, M, N, K, act, Y, ldy)


def gnn_tc_linear_dgrad(dY, lddy, W, ldw, M, N, K, accumulate, dX, lddx, precision, pack_ws=None, pack_ws_bytes=0,
                        stream=None):
    gnn_linear_dgrad(dY, lddy, W, ldw, M, N, K, accumulate, dX, lddx)


def gnn_tc_linear_wgrad(dY, lddy, X, ldx, M, N, K, dW, lddw, db, precision, stream=None):
    gnn_linear_wgrad(dY, lddy, X, ldx, M, N, K, dW, lddw, db)


def gnn_tc_pna_post_fwd(x, A, scal, wpost, wts, bpost, bts, n, f, f_out, towers, act, O, precision, pack_ws=None,
                        pack_ws_bytes=0, stream=None):
    gnn_pna_post_fwd(x, A, scal, wpost, wts, bpost, bts, n, f, f_out, towers, act, O)


def gnn_tc_pna_post_dgrad(dO, scal, wpost, wts, n, f, f_out, towers, dx, gA, precision, pack_ws=None,
                          pack_ws_bytes=0, stream=None):
    scratch = dx if dx is not None else torch.zeros(n, f, dtype=gA.dtype)
    gnn_pna_post_dgrad(dO, scal, wpost, wts, n, f, f_out, towers, scratch, gA)


def gnn_tc_pna_proj_fwd(x, wpre, wts, bpre, bts, n, f, towers, P, precision, pack_ws=None, pack_ws_bytes=0,
                        stream=None):
    c = towers * f
    for t in range(towers):
        w = _mat(wpre, f, 3 * f, 3 * f, t * wts)
        b = _flat(bpre, f, t * bts)
        P[:n, t * f:(t + 1) * f] = x[:n] @ w[:, :f].t() + b
        P[:n, c + t * f:c + (t + 1) * f] = x[:n] @ w[:, f:2 * f].t()


def gnn_tc_pna_dx_dgrad(dP, dO, wpre, wpre_ts, wpost, wpost_ts, n, f, f_out, towers, accumulate, dx, precision,
                        pack_ws=None, pack_ws_bytes=0, stream=None):
    c = towers * f
    acc = dx[:n].clone() if accumulate else torch.zeros_like(dx[:n])
    for t in range(towers):
        w = _mat(wpre, f, 3 * f, 3 * f, t * wpre_ts)
        acc += dP[:n, t * f:(t + 1) * f] @ w[:, :f] + dP[:n, c + t * f:c + (t + 1) * f] @ w[:, f:2 * f]
        wp = _mat(wpost, f_out, 13 * f, 13 * f, t * wpost_ts)
        acc += dO[:n, t * f_out:(t + 1) * f_out] @ wp[:, :f]
    dx[:n] = acc


def gnn_tc_pna_proj_wgrad(dP, x, n, f, towers, dwpre, wts, dbpre, bts, precision, stream=None):
    c = towers * f
    for t in range(towers):
        dw = _mat(dwpre, f, 3 * f, 3 * f, t * wts)
        dw[:, :f] += dP[:n, t * f:(t + 1) * f].t() @ x[:n]
        dw[:, f:2 * f] += dP[:n, c + t * f:c + (t + 1) * f].t() @ x[:n]
        if dbpre is not None:
            _flat(dbpre, f, t * bts).add_(dP[:n, t * f:(t + 1) * f].sum(0))


def gnn_tc_pna_post_wgrad(dO, x, A, scal, n, f, f_out, towers, dwpost, wts, dbpost, bts, precision, stream=None):
    gnn_pna_post_wgrad(dO, x, A, scal, n, f, f_out, towers, dwpost, wts, dbpost, bts)


def gnn_tc_pna_projsrc_fwd(x, wpre, wts, bpre, bts, n, f, towers, Psrc, precision, pack_ws=None, pack_ws_bytes=0,
                           stream=None):
    for t in range(towers):
        w = _mat(wpre, f, 3 * f, 3 * f, t * wts)
        Psrc[:n, t * f:(t + 1) * f] = x[:n] @ w[:, f:2 * f].t() + _flat(bpre, f, t * bts)


def _postf_cat(x, A, scal, n, f, towers, t):
    """[x | amp x | att x | A | amp A | att A] and the matching folded weight [F_out, 15F] builder."""
    a = A[:n].view(n, towers, 4 * f)[:, t]
    amp, att = scal[:n, 0:1], scal[:n, 1:2]
    return torch.cat([x[:n, :f], x[:n, :f] * amp, x[:n, :f] * att, a, a * amp, a * att], dim=1)


def gnn_tc_pna_postf_fwd(x, A, scal, Vx, wpost, wts, bpost, bts, n, f, f_out, towers, act, O, precision,
                         pack_ws=None, pack_ws_bytes=0, stream=None):
    v = Vx.view(towers, 3, f_out, f)
    for t in range(towers):
        wp = _mat(wpost, f_out, 13 * f, 13 * f, t * wts)
        w = torch.cat([v[t, 0], v[t, 1], v[t, 2], wp[:, f:]], dim=1)
        o = _postf_cat(x, A, scal, n, f, towers, t) @ w.t() + _flat(bpost, f_out, t * bts)
        O[:n, t * f_out:(t + 1) * f_out] = o.clamp(min=0) if act else o


def gnn_tc_pna_postf_agg_fwd(x, Psrc, tab, rowptr, col_src, col_code, scal, Vx, wpost, wts, bpost, bts, n, f, f_out, towers,
                             act, O, A_out, precision, pack_ws=None, pack_ws_bytes=0, stream=None):
    """aggregate (gnn_pna_aggregate_src_fwd) + folded post forward (gnn_tc_pna_postf_fwd) in one call."""
    A = A_out if A_out is not None else torch.zeros(n, towers * 4 * f, dtype=x.dtype)
    gnn_pna_aggregate_src_fwd(Psrc, tab, rowptr, col_src, col_code, n, towers, f, A)
    gnn_tc_pna_postf_fwd(x, A, scal, Vx, wpost, wts, bpost, bts, n, f, f_out, towers, act, O, precision)


def gnn_tc_pna_postf_wgrad(dO, x, A, scal, n, f, f_out, towers, dwpost, wts, dbpost, bts, dVx, precision,
                           stream=None):
    dv = dVx.view(towers, 3, f_out, f)
    for t in range(towers):
        d = dO[:n, t * f_out:(t + 1) * f_out]
        g = d.t() @ _postf_cat(x, A, scal, n, f, towers, t)          # [F_out, 15F]
        for s in range(3):
            dv[t, s] += g[:, s * f:(s + 1) * f]
        _mat(dwpost, f_out, 13 * f, 13 * f, t * wts)[:, f:] += g[:, 3 * f:]
        _flat(dbpost, f_out, t * bts).add_(d.sum(0))


def gnn_tc_pna_dxf_dgrad(dPsrc, dO, scal, wpre, wpre_ts, Vx, n, f, f_out, towers, accumulate, dx, precision,
                         pack_ws=None, pack_ws_bytes=0, stream=None):
    v = Vx.view(towers, 3, f_out, f)
    acc = dx[:n].clone() if accumulate else torch.zeros_like(dx[:n])
    sc = [torch.ones_like(scal[:n, 0:1]), scal[:n, 0:1], scal[:n, 1:2]]
    for t in range(towers):
        w = _mat(wpre, f, 3 * f, 3 * f, t * wpre_ts)
        acc += dPsrc[:n, t * f:(t + 1) * f] @ w[:, f:2 * f]
        d = dO[:n, t * f_out:(t + 1) * f_out]
        for s in range(3):
            acc += (d * sc[s]) @ v[t, s]
    dx[:n] = acc


def gnn_tc_pna_projsrc_wgrad(dPsrc, x, n, f, towers, dwpre, wts, dbpre, bts, precision, stream=None):
    for t in range(towers):
        dw = _mat(dwpre, f, 3 * f, 3 * f, t * wts)
        dw[:, f:2 * f] += dPsrc[:n, t * f:(t + 1) * f].t() @ x[:n]
        if dbpre is not None:
            _flat(dbpre, f, t * bts).add_(dPsrc[:n, t * f:(t + 1) * f].sum(0))


# ------------------------------------------------------------------ inference: whole conv stack per tile
_INFER_IMAGES = {}      # the CPU restatement keeps the weight references instead of packed tensor-core images


def gnn_tc_infer_image_bytes(num_layers):
    return 256


def gnn_tc_infer_pack(wpre, wpre_ts, Vx, wpost, wpost_ts, layer_stride, wlin_folded, num_layers, f, f_out, towers, images,
                      image_bytes, stream=None):
    _INFER_IMAGES[images.data_ptr()] = (wpre, wpre_ts, Vx, wpost, wpost_ts, layer_stride, wlin_folded)


def gnn_tc_infer_tiles(graph_ptr, rowptr, num_graphs, valid_graphs, tiles, max_tiles, num_tiles, err_flag=None,
                       stream=None, range_graphs=2048):
    """Greedy packing of consecutive graphs into tiles of <= 128 nodes and <= 1024 CSR slots, restarted every
    `range_graphs` graphs (one CTA per range in the CUDA kernel, which appends its tiles in no particular order)."""
    g_total = int(num_graphs) if valid_graphs is None else min(int(num_graphs), int(valid_graphs.reshape(-1)[0]))
    gp = graph_ptr.tolist()
    rp = rowptr.tolist()
    out = []
    for g0 in range(0, g_total, range_graphs):
        g, end = g0, min(g_total, g0 + range_graphs)
        while g < end:
            e = g
            while e < end and gp[e + 1] - gp[g] <= 128 and rp[gp[e + 1]] - rp[gp[g]] <= 1024:
                e += 1
            if e == g:
                if err_flag is not None:
                    err_flag |= 1
                e = g + 1
            elif gp[e] > gp[g]:
                out.append((gp[g], gp[e]))
            g = e
    num_tiles[0] = len(out)
    t = tiles.view(-1, 2)
    for i, (a, b) in enumerate(out[:max_tiles]):
        t[i, 0], t[i, 1] = a, b


def gnn_tc_pna_stack_infer(x0, rowptr, col_src, col_code, tiles, num_tiles, images, tab, bpre, bpre_ts, bpre_ls, bpost,
                           bpost_ts, bpost_ls, blin, avg_deg_log, num_layers, f, f_out, towers, skip, x_out, stream=None):
    wpre, wpre_ts, Vx, wpost, wpost_ts, ls, wlin = _INFER_IMAGES[images.data_ptr()]
    n, c = x0.shape[0], towers * f
    x = x0.clone()
    scal = torch.zeros(n, 2, dtype=x.dtype)
    for l in range(num_layers):
        psrc = torch.zeros(n, c, dtype=x.dtype)
        for t in range(towers):
            w = _mat(wpre, f, 3 * f, 3 * f, l * ls + t * wpre_ts)
            psrc[:, t * f:(t + 1) * f] = x @ w[:, f:2 * f].t() + _flat(bpre, f, l * bpre_ls + t * bpre_ts)
        A = torch.zeros(n, towers * 4 * f, dtype=x.dtype)
        gnn_pna_aggregate_src_fwd(psrc, tab[l], rowptr, col_src, col_code, n, towers, f, A)
        gnn_pna_scalers(rowptr, avg_deg_log[l:l + 1], n, scal)
        v = Vx[l].reshape(towers, 3, f_out, f)
        O = torch.zeros(n, towers * f_out, dtype=x.dtype)
        for t in range(towers):
            wp = _mat(wpost, f_out, 13 * f, 13 * f, l * ls + t * wpost_ts)
            w = torch.cat([v[t, 0], v[t, 1], v[t, 2], wp[:, f:]], dim=1)
            O[:, t * f_out:(t + 1) * f_out] = (_postf_cat(x, A, scal, n, f, towers, t) @ w.t() +
                                               _flat(bpost, f_out, l * bpost_ls + t * bpost_ts))
        y = (O @ wlin[l].t() + blin[l]).clamp(min=0)
        x = y + x if skip else y
    for a, b in tiles.view(-1, 2)[: int(num_tiles[0])].tolist():
        x_out[a:b] = x[a:b]


# ------------------------------------------------------------------ BatchNorm / pool / loss / optimiser
def gnn_bn_workspace_bytes(rows, c):
    return 256


def gnn_bn_stats(Y, rows, c, training, eps, momentum, running_mean, running_var, nbt, stats, valid_rows, workspace,
                 workspace_bytes, stream=None):
    if valid_rows is not None:          # padded batch: only the first valid_rows rows are real
        rows = max(min(rows, int(valid_rows)), 0)
    if training:
        y = Y[:rows].double()
        mean = y.mean(0)
        var = (y * y).mean(0) - mean * mean
        stats[:c] = mean.to(stats.dtype)
        stats[c:2 * c] = (1.0 / torch.sqrt(var + eps)).to(stats.dtype)
        if running_mean is not None:
            unbiased = var * rows / (rows - 1) if rows > 1 else var
            running_mean.copy_(((1 - momentum) * running_mean.double() + momentum * mean).to(running_mean.dtype))
            running_var.copy_(((1 - momentum) * running_var.double() + momentum * unbiased).to(running_var.dtype))
        if nbt is not None:
            nbt += 1
    else:
        stats[:c] = running_mean
        stats[c:2 * c] = (1.0 / torch.sqrt(running_var.double() + eps)).to(stats.dtype)


def gnn_bn_act_fwd(Y, stats, gamma, beta, skip, rows, c, relu, out, stream=None):
    v = _flat(gamma, c) * ((Y[:rows] - stats[:c]) * stats[c:2 * c]) + _flat(beta, c)
    if relu:
        v = v.clamp(min=0)
    if skip is not None:
        v = v + skip[:rows]
    out[:rows] = v


def gnn_bn_act_bwd(dOut, Y, stats, gamma, beta, rows, c, relu, training, dY, dgamma, dbeta, valid_rows, workspace,
                   workspace_bytes, stream=None):
    if valid_rows is not None:
        dY[:rows] = 0
        rows = max(min(rows, int(valid_rows)), 0)
    ga, be = _flat(gamma, c), _flat(beta, c)
    yh = (Y[:rows] - stats[:c]) * stats[c:2 * c]
    dz = dOut[:rows].clone()
    if relu:
        dz = torch.where(ga * yh + be > 0, dz, torch.zeros_like(dz))
    s, sy = dz.sum(0), (dz * yh).sum(0)
    if dbeta is not None:
        _flat(dbeta, c).add_(s)
    if dgamma is not None:
        _flat(dgamma, c).add_(sy)
    v = dz - s / rows - yh * sy / rows if training else dz
    dY[:rows] = ga * stats[c:2 * c] * v


def gnn_bn_fold_linear(W, ldw, bias, gamma, beta, running_mean, running_var, eps, N, K, Wf, bf, stream=None):
    sc = _flat(gamma, N) * (1.0 / torch.sqrt(running_var.double() + eps)).to(W.dtype)
    Wf.reshape(-1)[: N * K].view(N, K).copy_(_mat(W, N, K, ldw) * sc[:, None])
    b = _flat(bias, N) if bias is not None else torch.zeros(N, dtype=W.dtype)
    bf.reshape(-1)[:N] = (b - running_mean) * sc + _flat(beta, N)


def gnn_tc_linear_skip_fwd(X, ldx, W, ldw, bias, M, N, K, act, skip, ldskip, Y, ldy, precision, pack_ws=None,
                           pack_ws_bytes=0, stream=None):
    y = _mat(X, M, K, ldx) @ _mat(W, N, K, ldw).t() + _flat(bias, N)
    if act:
        y = y.clamp(min=0)
    if skip is not None:
        y = y + _mat(skip, M, N, ldskip)
    _mat(Y, M, N, ldy).copy_(y)


def gnn_segment_pool_fwd(x, graph_ptr, b, c, out, stream=None):
    gp = graph_ptr[: b + 1].tolist()
    for g in range(b):
        out[g] = x[gp[g]:gp[g + 1]].sum(0)


def gnn_segment_pool_bwd(dout, node_graph, n, c, accumulate, dx, stream=None):
    v = dout[node_graph[:n].long()] if node_graph is not None else dout[0].expand(n, c)
    dx[:n] = dx[:n] + v if accumulate else v


def gnn_mape_fwd_bwd(pred, target, numel, grad_scale, loss, dpred, valid_rows=None, row_len=1, stream=None):
    if valid_rows is not None:          # padded batch: the padding graph has no loss and a zero gradient
        if dpred is not None:
            dpred.reshape(-1)[:numel] = 0
        numel = max(min(numel, int(valid_rows) * row_len), 0)
    p, t = pred.reshape(-1)[:numel], target.reshape(-1)[:numel]
    den = t.abs().clamp(min=1.17e-06)
    if loss is not None:
        loss[0] = ((p - t).abs() / den).sum() / max(numel, 1)
    if dpred is not None:
        dpred.reshape(-1)[:numel] = torch.sign(p - t) / den / max(numel, 1) * grad_scale


def gnn_adamw_amsgrad_step(param, grad, m, v, vmax, numel, step_count, lr0, beta1, beta2, eps, wd, t0, grad_scale,
                           stream=None):
    t = int(step_count[0])
    lr = 0.5 * lr0 * (1.0 + math.cos(math.pi * (t % t0) / t0)) if t0 > 0 else lr0
    step = t + 1
    bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
    g = grad[:numel] * grad_scale
    param[:numel] *= 1 - lr * wd
    m[:numel] = m[:numel] + (g - m[:numel]) * (1 - beta1)
    v[:numel] = v[:numel] * beta2 + (1 - beta2) * g * g
    vmax[:numel] = torch.maximum(vmax[:numel], v[:numel])
    denom = vmax[:numel].sqrt() / math.sqrt(bc2) + eps
    param[:numel] -= (lr / bc1) * (m[:numel] / denom)


def gnn_step_increment(step_count, stream=None):
    step_count += 1


def gnn_fill_zero(ptr, nbytes, stream=None):
    ptr.view(torch.uint8).reshape(-1)[:nbytes] = 0


def install(monkeypatch, lib_module, models_module):
    """Route the product's ctypes dispatcher to this module (CPU host-logic tests only)."""
    here = globals()

    def call(name, *args):
        here[name](*args)

    def query(name, *args):
        return int(here[name](*args))

    monkeypatch.setattr(lib_module, "call", call)
    monkeypatch.setattr(lib_module, "query", query)
    monkeypatch.setattr(lib_module, "stream_ptr", lambda: None)
    monkeypatch.setattr(models_module, "_require_cuda", lambda t, what: None)
