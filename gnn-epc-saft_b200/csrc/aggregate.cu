// Fused edge-message + multi-aggregator (mean/min/max/std) kernels, forward and backward.
// Replaces MessagePassing._collect + PNAConv.message + MultiAggregation(mean,min,max,std)
// (PyG nn/conv/pna_conv.py, nn/aggr/{multi,basic}.py) called at reference
// gnnepcsaft/train/models.py:128.
//
// Layout: one warp per destination node, lanes across channels with 128-bit accesses
// (a warp covers 128 channels per chunk; C = towers*F channels, CHUNKS = ceil(C/128)).
// The in-edges of a node are a contiguous CSR segment, walked serially in PyG edge order, so
// no atomics are needed and the summation order is deterministic.  HBM-bound: per node it
// reads one gathered row per in-edge (L2-resident re-reads) and writes 4*C floats.
#include "common.cuh"
#include "pna_agg.cuh"

#include <float.h>

constexpr int AGG_WARPS = 8;
constexpr int AGG_THREADS = AGG_WARPS * 32;

__device__ __forceinline__ void welford_update(float& mean, float& m2, float q, float inv_k) {
    float d = q - mean;
    mean = fmaf(d, inv_k, mean);
    m2 = fmaf(d, q - mean, m2);
}
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return pna_add4(a, b); }

// std of DegreeScalerAggregation's StdAggregation: sqrt(clamp(var,1e-5)) masked to 0 at the floor
__device__ __forceinline__ float pna_std(float var) {
    float s = sqrtf(fmaxf(var, 1e-5f));
    return s <= sqrtf(1e-5f) ? 0.0f : s;
}

__device__ __forceinline__ void store_aggregates(float* __restrict__ A, int64_t node, int C, int F, int c,
                                                 float4 mean, float4 mn, float4 mx, float4 sd) {
    // A[node][t][agg][f], channel c = t*F + f  (F % 4 == 0 so a float4 never straddles towers)
    int t = c / F, f = c - t * F;
    float* base = A + node * (4 * (int64_t)C) + (int64_t)t * 4 * F + f;
    st4_stream(base, mean);
    st4_stream(base + F, mn);
    st4_stream(base + 2 * F, mx);
    st4_stream(base + 3 * F, sd);
}

// ------------------------------------------------------------------ forward, factorised message
// Persistent grid, one warp per destination node, lanes over channels (NCHUNK x 128, 128-bit accesses).  The in-edges
// of a node are a contiguous CSR segment walked in PyG edge order (no atomics, deterministic).  ncu on the first
// version (one node per warp iteration): DRAM 36 % busy, issue slots 41 %, 12.5 of 17.8 resident warps per issue
// waiting on a long scoreboard - the dependent chain rowptr -> col_src/col_code -> gathered rows -> stores costs four
// exposed memory latencies per node.  Here a warp owns nodes warp, warp + W, ... and
//   * requests the rowptr pair of node i+2 and the column indices of node i+1 before it touches node i,
//   * issues the gathers of the first PF in-edges of node i together (one exposed latency per node),
// so that only the gather latency itself is left on the critical path.  Statistics: pna_agg.cuh (single pass, sums
// shifted by the first message).  Results leave through streaming 128-bit stores.
// SRC_ONLY: P is [n, C] = source projections only and the aggregates are those of the shift-free q (the destination
// term is folded into the post GEMM's weights, gnn_pna_fold_fwd) - the kernel then reads half as much.

template <int NCHUNK, bool SRC_ONLY, int PF, int MINB>
__global__ void __launch_bounds__(AGG_THREADS, MINB)
pna_aggregate_fwd_pipe_kernel(const float* __restrict__ P, const float* __restrict__ tab,
                              const int* __restrict__ rowptr, const int* __restrict__ col_src,
                              const uint8_t* __restrict__ col_code, int n, int C, int F, float* __restrict__ A) {
    const int lane = threadIdx.x & 31;
    const int warp0 = blockIdx.x * AGG_WARPS + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * AGG_WARPS;
    const int ldp = SRC_ONLY ? C : 2 * C, lda = 4 * C, soff = SRC_ONLY ? 0 : C;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    int cj[NCHUNK], aoff[NCHUNK];
    bool live[NCHUNK];
#pragma unroll
    for (int j = 0; j < NCHUNK; ++j) {
        cj[j] = j * 128 + lane * 4;
        live[j] = cj[j] < C;
        if (!live[j]) cj[j] = 0;                          // dead lanes read column 0 (valid memory), never store
        const int t = cj[j] / F;
        aoff[j] = t * 4 * F + (cj[j] - t * F);            // A[node][t][agg][f]
    }
    int node = warp0;
    if (node >= n) return;
    // prologue: indices of the first node, rowptr pair of the second
    int beg, cnt, raw_src = 0, raw_code = 0;
    {
        const int rp = __ldg(rowptr + node + (lane & 1));
        beg = __shfl_sync(GNN_FULL_MASK, rp, 0);
        // REDUX writes a uniform register: the edge loops below are then provably warp-uniform and the compiler drops
        // the divergence bookkeeping (BSSY / WARPSYNC) around the shuffles
        cnt = __reduce_max_sync(GNN_FULL_MASK, __shfl_sync(GNN_FULL_MASK, rp, 1) - beg);
        if (lane < cnt) {
            raw_src = __ldg(col_src + beg + lane);
            raw_code = (int)__ldg(col_code + beg + lane);
        }
    }
    int node_n = node + nwarps;
    int rp_n = node_n < n ? __ldg(rowptr + node_n + (lane & 1)) : 0;
    while (true) {
        // ---- (1) next node: column indices now (its rowptr pair was requested one iteration ago), rowptr of the one after
        int beg_n = 0, cnt_n = 0, raw_src_n = 0, raw_code_n = 0;
        if (node_n < n) {
            beg_n = __shfl_sync(GNN_FULL_MASK, rp_n, 0);
            cnt_n = __reduce_max_sync(GNN_FULL_MASK, __shfl_sync(GNN_FULL_MASK, rp_n, 1) - beg_n);
            if (lane < cnt_n) {
                raw_src_n = __ldg(col_src + beg_n + lane);
                raw_code_n = (int)__ldg(col_code + beg_n + lane);
            }
        }
        const int node_nn = node_n + nwarps;
        const int rp_nn = node_nn < n ? __ldg(rowptr + node_nn + (lane & 1)) : 0;
        // ---- (2) this node
        const int my_src = raw_src * ldp + soff, my_code = raw_code * C;
        float* const arow = A + (size_t)node * lda;
        const int m = min(cnt, 32);
#pragma unroll
        for (int j = 0; j < NCHUNK; ++j) {
            float* const out = arow + aoff[j];
            if (cnt == 0) {
                if (live[j]) {
                    st4_stream(out, zero);
                    st4_stream(out + F, zero);
                    st4_stream(out + 2 * F, zero);
                    st4_stream(out + 3 * F, zero);
                }
                continue;
            }
            const float* const pbase = P + cj[j];
            const float* const tbase = tab + cj[j];
            float4 pd = zero;
            if (!SRC_ONLY) pd = ldg4(P + (size_t)node * ldp + cj[j]);
            float4 pr[PF];
            int co[PF];
#pragma unroll
            for (int k = 0; k < PF; ++k) {                 // all gathers of the first PF edges in flight together
                const int so = __shfl_sync(GNN_FULL_MASK, my_src, k);
                co[k] = __shfl_sync(GNN_FULL_MASK, my_code, k);
                if (k < m) pr[k] = ldg4(pbase + so);
            }
            PnaAcc<false> acc;
            acc.first(add4(pr[0], ldg4(tbase + co[0])));    // bond-table rows are L1 hits: fetched at use
#pragma unroll
            for (int k = 1; k < PF; ++k)
                if (k < m) acc.next(add4(pr[k], ldg4(tbase + co[k])));
            for (int k = PF; k < m; ++k) {                 // in-degree > PF: serial
                const int so = __shfl_sync(GNN_FULL_MASK, my_src, k), ck = __shfl_sync(GNN_FULL_MASK, my_code, k);
                acc.next(add4(ldg4(pbase + so), ldg4(tbase + ck)));
            }
            for (int s = beg + 32; s < beg + cnt; ++s)     // in-degree > 32 (never for molecules)
                acc.next(add4(ldg4(pbase + __ldg(col_src + s) * ldp + soff), ldg4(tbase + (int)__ldg(col_code + s) * C)));
            if (live[j]) {
                float4 mean, sd;
                acc.finish(cnt, mean, sd);
                st4_stream(out, add4(mean, pd));
                st4_stream(out + F, add4(acc.mn, pd));
                st4_stream(out + 2 * F, add4(acc.mx, pd));
                st4_stream(out + 3 * F, sd);
            }
        }
        // ---- (3) rotate
        if (node_n >= n) break;
        node = node_n; beg = beg_n; cnt = cnt_n; raw_src = raw_src_n; raw_code = raw_code_n;
        node_n = node_nn; rp_n = rp_nn;
    }
}

static int agg_grid(int64_t n) {
    int64_t blocks = gnn_ceil_div(n, AGG_WARPS);
    int64_t cap = (int64_t)gnn_sm_count() * 64;   // grid-stride beyond 64 CTAs/SM worth of work
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

template <bool SRC_ONLY>
static int aggregate_fwd_launch(const float* P, const float* tab, const int32_t* rowptr, const int32_t* col_src,
                                const uint8_t* col_code, int64_t n, int towers, int F, float* A, void* stream) {
    if (n < 0 || towers < 1 || F < 4 || (F & 3)) return GNN_ERR_BAD_ARG;
    if (n == 0) return 0;
    const int C = towers * F;
    if ((int64_t)n * 2 * C >= (1ll << 31)) return GNN_ERR_TOO_LARGE;    // 32-bit row offsets in the gather
    cudaStream_t st = gnn_stream(stream);
    const int chunks = (C + 127) / 128;
    // persistent grid: 3 resident CTAs per SM (80 registers per thread)
    int64_t pb = (int64_t)gnn_sm_count() * 3;
    if (pb > gnn_ceil_div(n, AGG_WARPS)) pb = gnn_ceil_div(n, AGG_WARPS);
#define LAUNCH_PIPE(CH) pna_aggregate_fwd_pipe_kernel<CH, SRC_ONLY, 3, 3><<<(unsigned)pb, AGG_THREADS, 0, st>>>(P, tab, rowptr, col_src, col_code, (int)n, C, F, A)
    switch (chunks) {
        case 1: LAUNCH_PIPE(1); break;
        case 2: LAUNCH_PIPE(2); break;
        case 3: case 4: LAUNCH_PIPE(4); break;
        case 5: case 6: case 7: case 8: LAUNCH_PIPE(8); break;
        default: return GNN_ERR_UNSUPPORTED;
    }
#undef LAUNCH_PIPE
    return gnn_launch_status();
}

extern "C" int gnn_pna_aggregate_fwd(const float* P, const float* tab, const int32_t* rowptr,
                                     const int32_t* col_src, const uint8_t* col_code, int64_t n, int towers,
                                     int F, float* A, void* stream) {
    return aggregate_fwd_launch<false>(P, tab, rowptr, col_src, col_code, n, towers, F, A, stream);
}
extern "C" int gnn_pna_aggregate_src_fwd(const float* Psrc, const float* tab, const int32_t* rowptr,
                                         const int32_t* col_src, const uint8_t* col_code, int64_t n, int towers,
                                         int F, float* A, void* stream) {
    return aggregate_fwd_launch<true>(Psrc, tab, rowptr, col_src, col_code, n, towers, F, A, stream);
}

// ------------------------------------------------------------------ backward, factorised message
// Per destination node: recompute the segment statistics of q (shifted by the first message exactly as in the
// forward kernel), count ties, then emit
//   dq_e = g_mean/cnt + g_min*[q_e==min]/n_min + g_max*[q_e==max]/n_max + g_std*(q_e-mean)/(cnt*std)
// per CSR slot, dPdst = (g_mean+g_min+g_max)*[cnt>0] and the per-bond-code sum of dq.  sm_100 has no native
// fp32 shared-memory atomic add (it compiles to a CAS spin loop), so the bond-code sums go to one of DTAB_PARTS
// partial tables in global memory with native 128-bit reductions (RED.E.ADD.F32x4 at the L2), summed afterwards.
constexpr int DTAB_PARTS = 64;

__device__ __forceinline__ float rcp_approx(float x) { return pna_rcp_approx(x); }

// statistics of one more message for pass 1 (shifted sum, min / max with tie counts)
__device__ __forceinline__ void bwd_stat(float q, float q0, float& s1, float& mn, float& mx, float& nmin, float& nmax) {
    s1 += q - q0;
    // branch-free: a new extremum resets the count to 1, a tie adds 1 (selects only; the nested ?: form compiled to
    // per-lane divergent branches)
    nmin = (q < mn ? 0.f : nmin) + (q <= mn ? 1.f : 0.f);
    mn = fminf(mn, q);
    nmax = (q > mx ? 0.f : nmax) + (q >= mx ? 1.f : 0.f);
    mx = fmaxf(mx, q);
}

// Measured (ncu, 4096-graph batch): 720 warp instructions per node, issue slots 56 % busy, DRAM 50 %; a persistent
// software-pipelined variant like pna_aggregate_fwd_pipe_kernel (indices prefetched across nodes, message rows kept
// in registers for pass 2) ran at the same 140-145 us with or without the stores and reductions, so it was dropped:
// this kernel is bound by its instruction stream (per-slot addressing, tie counting), not by memory latency.
template <bool SRC_ONLY>
__global__ void __launch_bounds__(AGG_THREADS, 3)
pna_aggregate_bwd_kernel(const float* __restrict__ gA, const float* __restrict__ P,
                         const float* __restrict__ tab, const float* __restrict__ A,
                         const int* __restrict__ rowptr, const int* __restrict__ col_src,
                         const uint8_t* __restrict__ col_code, int n, int C, int F,
                         float* __restrict__ dP, float* __restrict__ dq, float* __restrict__ dtab_parts) {
    const int lane = threadIdx.x & 31;
    const int warp0 = blockIdx.x * AGG_WARPS + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * AGG_WARPS;
    const int ldp = SRC_ONLY ? C : 2 * C, soff = SRC_ONLY ? 0 : C;
    float* dtab = dtab_parts + (size_t)(blockIdx.x % DTAB_PARTS) * GNN_NUM_BOND_CODES * C;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int node = warp0; node < n; node += nwarps) {
        const int rp = __ldg(rowptr + node + (lane & 1));
        const int beg = __shfl_sync(GNN_FULL_MASK, rp, 0), end = __shfl_sync(GNN_FULL_MASK, rp, 1);
        // REDUX writes a uniform register: the compiler then knows that the edge loops below are warp-uniform
        // (no divergence bookkeeping around the shuffles)
        const int cnt = __reduce_max_sync(GNN_FULL_MASK, end - beg);
        int my_src = 0, my_code = 0;
        if (lane < cnt) {
            my_src = __ldg(col_src + beg + lane) * ldp + soff;
            my_code = (int)__ldg(col_code + beg + lane) * C;
        }
        const int m = min(cnt, 32);
        for (int cb = 0; cb < C; cb += 128) {                  // warp-uniform trip count (shuffles inside)
            const int c = cb + lane * 4;
            const bool live = c < C;
            if (cnt == 0) {
                if (!SRC_ONLY && live) st4(dP + (size_t)node * ldp + c, zero);
                continue;
            }
            const float* pbase = P + (live ? c : 0);           // dead lanes read column 0 (valid memory, results unused)
            const float* tbase = tab + (live ? c : 0);
            // gradient rows of this node: requested before pass 1, used after it
            float4 g_mean = zero, g_min = zero, g_max = zero, g_std = zero, sd = zero;
            if (live) {
                const int t = c / F, f = c - t * F;
                const size_t off = (size_t)node * (4 * (size_t)C) + (size_t)t * 4 * F + f;
                g_mean = ldg4(gA + off);
                g_min = ldg4(gA + off + F);
                g_max = ldg4(gA + off + 2 * F);
                g_std = ldg4(gA + off + 3 * F);
                sd = ldg4(A + off + 3 * F);
            }
            // ---- pass 1: shifted sum, min / max with tie counts; the row of edge k+1 is in flight while k is used
            int so = __shfl_sync(GNN_FULL_MASK, my_src, 0), co = __shfl_sync(GNN_FULL_MASK, my_code, 0);
            float4 pn = ldg4(pbase + so), tn = ldg4(tbase + co);
            const float4 q0 = add4(pn, tn);
            float4 s1 = zero, mn = q0, mx = q0;
            float4 nmin = make_float4(1.f, 1.f, 1.f, 1.f), nmax = nmin;
            if (1 < m) {
                so = __shfl_sync(GNN_FULL_MASK, my_src, 1);
                co = __shfl_sync(GNN_FULL_MASK, my_code, 1);
                pn = ldg4(pbase + so);
                tn = ldg4(tbase + co);
            }
            for (int k = 1; k < m; ++k) {
                const float4 q = add4(pn, tn);
                if (k + 1 < m) {
                    so = __shfl_sync(GNN_FULL_MASK, my_src, k + 1);
                    co = __shfl_sync(GNN_FULL_MASK, my_code, k + 1);
                    pn = ldg4(pbase + so);
                    tn = ldg4(tbase + co);
                }
                bwd_stat(q.x, q0.x, s1.x, mn.x, mx.x, nmin.x, nmax.x);
                bwd_stat(q.y, q0.y, s1.y, mn.y, mx.y, nmin.y, nmax.y);
                bwd_stat(q.z, q0.z, s1.z, mn.z, mx.z, nmin.z, nmax.z);
                bwd_stat(q.w, q0.w, s1.w, mn.w, mx.w, nmin.w, nmax.w);
            }
            for (int s = beg + 32; s < end; ++s) {             // in-degree > 32 (never for molecules)
                const float4 q = add4(ldg4(pbase + __ldg(col_src + s) * ldp + soff), ldg4(tbase + (int)__ldg(col_code + s) * C));
                bwd_stat(q.x, q0.x, s1.x, mn.x, mx.x, nmin.x, nmax.x);
                bwd_stat(q.y, q0.y, s1.y, mn.y, mx.y, nmin.y, nmax.y);
                bwd_stat(q.z, q0.z, s1.z, mn.z, mx.z, nmin.z, nmax.z);
                bwd_stat(q.w, q0.w, s1.w, mn.w, mx.w, nmin.w, nmax.w);
            }
            const float inv_cnt = rcp_approx((float)cnt);
            const float4 mean = make_float4(fmaf(s1.x, inv_cnt, q0.x), fmaf(s1.y, inv_cnt, q0.y),
                                            fmaf(s1.z, inv_cnt, q0.z), fmaf(s1.w, inv_cnt, q0.w));
            if (!SRC_ONLY && live)
                st4(dP + (size_t)node * ldp + c, make_float4(g_mean.x + g_min.x + g_max.x, g_mean.y + g_min.y + g_max.y,
                                                             g_mean.z + g_min.z + g_max.z, g_mean.w + g_min.w + g_max.w));
            float4 cm, cmin, cmax, cs;
#define COEF(X)                                              \
    cm.X = g_mean.X * inv_cnt;                               \
    cmin.X = g_min.X * rcp_approx(nmin.X);                   \
    cmax.X = g_max.X * rcp_approx(nmax.X);                   \
    cs.X = sd.X > 0.f ? g_std.X * inv_cnt * rcp_approx(sd.X) : 0.f;
            COEF(x) COEF(y) COEF(z) COEF(w)
#undef COEF
            // ---- pass 2: per-slot gradient (rows are L1/L2 hits from pass 1)
            float* dq_row = dq != nullptr ? dq + (size_t)beg * C + c : nullptr;
            so = __shfl_sync(GNN_FULL_MASK, my_src, 0);
            co = __shfl_sync(GNN_FULL_MASK, my_code, 0);
            pn = ldg4(pbase + so);
            tn = ldg4(tbase + co);
            for (int k = 0; k < cnt; ++k) {
                const float4 q = add4(pn, tn);
                const int so_k = so, co_k = co;
                if (k + 1 < cnt) {
                    if (k + 1 < 32) {
                        so = __shfl_sync(GNN_FULL_MASK, my_src, k + 1);
                        co = __shfl_sync(GNN_FULL_MASK, my_code, k + 1);
                    } else {
                        so = __ldg(col_src + beg + k + 1) * ldp + soff;
                        co = (int)__ldg(col_code + beg + k + 1) * C;
                    }
                    pn = ldg4(pbase + so);
                    tn = ldg4(tbase + co);
                }
                float4 g;
#define GRAD(X) \
    g.X = cm.X + (q.X == mn.X ? cmin.X : 0.f) + (q.X == mx.X ? cmax.X : 0.f) + cs.X * (q.X - mean.X);
                GRAD(x) GRAD(y) GRAD(z) GRAD(w)
#undef GRAD
                if (live) {
                    if (dq != nullptr) st4(dq_row + k * C, g);
                    else atomicAdd(reinterpret_cast<float4*>(dP + so_k + c), g);  // dPsrc[src] += g  (so = src*ldp + soff)
                    atomicAdd(reinterpret_cast<float4*>(dtab + co_k + c), g);     // RED.E.ADD.F32x4
                }
            }
        }
    }
}

// dtab_accum[k] += sum over the partial tables; the partial tables are cleared for the next call.
// block (64 float4 columns, 4 part phases): every thread sums DTAB_PARTS / 4 partials with independent 16-byte loads
// (all in flight together), the phases meet in shared memory.  total % 4 == 0 (C % 4 == 0).
__global__ void __launch_bounds__(256)
dtab_reduce_kernel(float* __restrict__ parts, int total, float* __restrict__ dtab_accum) {
    __shared__ float4 red[4][64];
    const int i4 = blockIdx.x * 64 + threadIdx.x;          // float4 column
    const int total4 = total >> 2;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i4 < total4) {
        float4 v[DTAB_PARTS / 4];
#pragma unroll
        for (int k = 0; k < DTAB_PARTS / 4; ++k)
            v[k] = *reinterpret_cast<const float4*>(parts + (size_t)(threadIdx.y + 4 * k) * total + 4 * i4);
#pragma unroll
        for (int k = 0; k < DTAB_PARTS / 4; ++k) {
            acc = add4(acc, v[k]);
            *reinterpret_cast<float4*>(parts + (size_t)(threadIdx.y + 4 * k) * total + 4 * i4) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    red[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && i4 < total4) {
        const float4 t = add4(add4(red[0][threadIdx.x], red[1][threadIdx.x]), add4(red[2][threadIdx.x], red[3][threadIdx.x]));
        float4* dst = reinterpret_cast<float4*>(dtab_accum + 4 * i4);
        *dst = add4(*dst, t);
    }
}

extern "C" int64_t gnn_pna_aggregate_bwd_workspace_bytes(int towers, int F) {
    return (int64_t)DTAB_PARTS * GNN_NUM_BOND_CODES * towers * F * (int64_t)sizeof(float);
}

template <bool SRC_ONLY>
static int aggregate_bwd_launch(const float* gA, const float* P, const float* tab, const float* A,
                                const int32_t* rowptr, const int32_t* col_src, const uint8_t* col_code,
                                int64_t n, int towers, int F, float* dP, float* dq, float* dtab_accum,
                                void* zeroed_workspace, int64_t workspace_bytes, void* stream) {
    if (n < 0 || towers < 1 || F < 4 || (F & 3)) return GNN_ERR_BAD_ARG;
    if (n == 0) return 0;
    if (workspace_bytes < gnn_pna_aggregate_bwd_workspace_bytes(towers, F)) return GNN_ERR_WORKSPACE;
    if (SRC_ONLY && dq == nullptr && dP == nullptr) return GNN_ERR_BAD_ARG;
    const int C = towers * F;
    if ((int64_t)n * 2 * C >= (1ll << 31)) return GNN_ERR_TOO_LARGE;    // 32-bit row offsets
    cudaStream_t st = gnn_stream(stream);
    int64_t blocks = gnn_ceil_div(n, AGG_WARPS);
    int64_t cap = (int64_t)gnn_sm_count() * 32;
    if (blocks > cap) blocks = cap;
    float* parts = reinterpret_cast<float*>(zeroed_workspace);
    if (dq == nullptr) {       // source gradients accumulate atomically into the source columns of dP
        cudaError_t e = SRC_ONLY ? cudaMemsetAsync(dP, 0, (size_t)n * C * sizeof(float), st)
                                 : cudaMemset2DAsync(dP + C, 2 * (size_t)C * sizeof(float), 0, (size_t)C * sizeof(float),
                                                     (size_t)n, st);
        if (e != cudaSuccess) return (int)e;
    }
    pna_aggregate_bwd_kernel<SRC_ONLY><<<(unsigned)blocks, AGG_THREADS, 0, st>>>(gA, P, tab, A, rowptr, col_src, col_code,
                                                                                 (int)n, C, F, dP, dq, parts);
    const int total = GNN_NUM_BOND_CODES * C;
    dtab_reduce_kernel<<<(total / 4 + 63) / 64, dim3(64, 4), 0, st>>>(parts, total, dtab_accum);
    return gnn_launch_status();
}

extern "C" int gnn_pna_aggregate_bwd(const float* gA, const float* P, const float* tab, const float* A,
                                     const int32_t* rowptr, const int32_t* col_src, const uint8_t* col_code,
                                     int64_t n, int towers, int F, float* dP, float* dq, float* dtab_accum,
                                     void* zeroed_workspace, int64_t workspace_bytes, void* stream) {
    return aggregate_bwd_launch<false>(gA, P, tab, A, rowptr, col_src, col_code, n, towers, F, dP, dq,
                                       dtab_accum, zeroed_workspace, workspace_bytes, stream);
}
extern "C" int gnn_pna_aggregate_src_bwd(const float* gA, const float* Psrc, const float* tab, const float* A,
                                         const int32_t* rowptr, const int32_t* col_src,
                                         const uint8_t* col_code, int64_t n, int towers, int F, float* dPsrc, float* dq,
                                         float* dtab_accum, void* zeroed_workspace, int64_t workspace_bytes,
                                         void* stream) {
    return aggregate_bwd_launch<true>(gA, Psrc, tab, A, rowptr, col_src, col_code, n, towers, F, dPsrc, dq,
                                      dtab_accum, zeroed_workspace, workspace_bytes, stream);
}

// ------------------------------------------------------------------ gather dq by source node
__global__ void __launch_bounds__(AGG_THREADS)
scatter_src_bwd_kernel(const float* __restrict__ dq, const int* __restrict__ srcptr,
                       const int* __restrict__ src_slot, int64_t n, int C, float* __restrict__ out, int64_t ldo) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * AGG_WARPS + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * AGG_WARPS;
    for (int64_t node = warp0; node < n; node += nwarps) {
        const int beg = srcptr[node], end = srcptr[node + 1];
        for (int c = lane * 4; c < C; c += 128) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int k = beg; k < end; ++k) {
                const int slot = __ldg(src_slot + k);
                acc = add4(acc, ldg4(dq + (int64_t)slot * C + c));
            }
            st4(out + node * ldo + c, acc);
        }
    }
}

extern "C" int gnn_scatter_src_bwd(const float* dq, const int32_t* srcptr, const int32_t* src_slot, int64_t n,
                                   int C, float* dP, void* stream) {
    if (n < 0 || C < 4 || (C & 3)) return GNN_ERR_BAD_ARG;
    if (n == 0) return 0;
    scatter_src_bwd_kernel<<<agg_grid(n), AGG_THREADS, 0, gnn_stream(stream)>>>(dq, srcptr, src_slot, n, C, dP + C,
                                                                                 2 * (int64_t)C);
    return gnn_launch_status();
}
// out[j, 0:C] (row stride ldo) = sum over out-edges of j of dq[slot]
extern "C" int gnn_slot_sum_by_source(const float* dq, const int32_t* srcptr, const int32_t* src_slot, int64_t n,
                                      int C, float* out, int64_t ldo, void* stream) {
    if (n < 0 || C < 4 || (C & 3) || (ldo & 3) || ldo < C) return GNN_ERR_BAD_ARG;
    if (n == 0) return 0;
    scatter_src_bwd_kernel<<<agg_grid(n), AGG_THREADS, 0, gnn_stream(stream)>>>(dq, srcptr, src_slot, n, C, out, ldo);
    return gnn_launch_status();
}

// ------------------------------------------------------------------ general message path (pre_layers > 1)
// h[s,:] = act(Pdst[dst] + Psrc[src_s] + tab[code_s]) for every CSR slot s of dst
__global__ void __launch_bounds__(AGG_THREADS)
edge_hidden_fwd_kernel(const float* __restrict__ P, const float* __restrict__ tab, const int* __restrict__ rowptr,
                       const int* __restrict__ col_src, const uint8_t* __restrict__ col_code, int64_t n, int C,
                       int relu, float* __restrict__ h) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * AGG_WARPS + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * AGG_WARPS;
    const int64_t ldp = 2 * (int64_t)C;
    for (int64_t node = warp0; node < n; node += nwarps) {
        const int beg = rowptr[node], end = rowptr[node + 1];
        for (int c = lane * 4; c < C; c += 128) {
            float4 pd = ldg4(P + node * ldp + c);
            for (int s = beg; s < end; ++s) {
                const int src = __ldg(col_src + s);
                const int code = __ldg(col_code + s);
                float4 v = add4(pd, add4(ldg4(P + (int64_t)src * ldp + C + c), ldg4(tab + (int64_t)code * C + c)));
                if (relu) v = make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
                st4(h + (int64_t)s * C + c, v);
            }
        }
    }
}
extern "C" int gnn_edge_hidden_fwd(const float* P, const float* tab, const int32_t* rowptr, const int32_t* col_src,
                                   const uint8_t* col_code, int64_t n, int C, int relu, float* hidden_out,
                                   void* stream) {
    if (n < 0 || C < 4 || (C & 3)) return GNN_ERR_BAD_ARG;
    if (n == 0) return 0;
    edge_hidden_fwd_kernel<<<agg_grid(n), AGG_THREADS, 0, gnn_stream(stream)>>>(P, tab, rowptr, col_src, col_code, n,
                                                                                 C, relu, hidden_out);
    return gnn_launch_status();
}

// dh [E',C] -> dPdst[i] = sum over slots of i; dtab[code] += dh[s]  (dq for the source scatter is dh itself)
__global__ void __launch_bounds__(AGG_THREADS)
edge_hidden_bwd_kernel(const float* __restrict__ dh, const int* __restrict__ rowptr,
                       const uint8_t* __restrict__ col_code, int64_t n, int C, float* __restrict__ dP,
                       float* dtab) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * AGG_WARPS + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * AGG_WARPS;
    for (int64_t node = warp0; node < n; node += nwarps) {
        const int beg = rowptr[node], end = rowptr[node + 1];
        for (int c = lane * 4; c < C; c += 128) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int s = beg; s < end; ++s) {
                float4 g = ldg4(dh + (int64_t)s * C + c);
                acc = add4(acc, g);
                float* dt = dtab + (int)__ldg(col_code + s) * C + c;
                atomicAdd(dt, g.x);
                atomicAdd(dt + 1, g.y);
                atomicAdd(dt + 2, g.z);
                atomicAdd(dt + 3, g.w);
            }
            st4(dP + node * (2 * (int64_t)C) + c, acc);
        }
    }
}
extern "C" int gnn_edge_hidden_bwd(const float* dh, const int32_t* rowptr, const uint8_t* col_code, int64_t n, int C,
                                   float* dP, float* dtab_accum, void* stream) {
    if (n < 0 || C < 4 || (C & 3)) return GNN_ERR_BAD_ARG;
    if (n == 0) return 0;
    int64_t blocks = gnn_ceil_div(n, AGG_WARPS);
    int64_t cap = (int64_t)gnn_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    edge_hidden_bwd_kernel<<<(unsigned)blocks, AGG_THREADS, 0, gnn_stream(stream)>>>(dh, rowptr, col_code, n, C, dP,
                                                                                      dtab_accum);
    return gnn_launch_status();
}

// aggregate materialised messages m [E',C] (CSR slot order)
__global__ void __launch_bounds__(AGG_THREADS)
aggregate_msg_fwd_kernel(const float* __restrict__ m, const int* __restrict__ rowptr, int64_t n, int C, int F,
                         float* __restrict__ A) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * AGG_WARPS + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * AGG_WARPS;
    for (int64_t node = warp0; node < n; node += nwarps) {
        const int beg = rowptr[node], end = rowptr[node + 1];
        const int cnt = end - beg;
        for (int c = lane * 4; c < C; c += 128) {
            float4 mean = make_float4(0.f, 0.f, 0.f, 0.f), m2 = mean;
            float4 mn = make_float4(FLT_MAX, FLT_MAX, FLT_MAX, FLT_MAX);
            float4 mx = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX);
            for (int s = beg; s < end; ++s) {
                float4 q = ldg4(m + (int64_t)s * C + c);
                const float inv_k = 1.0f / (float)(s - beg + 1);
                welford_update(mean.x, m2.x, q.x, inv_k);
                welford_update(mean.y, m2.y, q.y, inv_k);
                welford_update(mean.z, m2.z, q.z, inv_k);
                welford_update(mean.w, m2.w, q.w, inv_k);
                mn = make_float4(fminf(mn.x, q.x), fminf(mn.y, q.y), fminf(mn.z, q.z), fminf(mn.w, q.w));
                mx = make_float4(fmaxf(mx.x, q.x), fmaxf(mx.y, q.y), fmaxf(mx.z, q.z), fmaxf(mx.w, q.w));
            }
            float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            if (cnt == 0) {
                store_aggregates(A, node, C, F, c, z, z, z, z);
            } else {
                const float ic = 1.0f / (float)cnt;
                store_aggregates(A, node, C, F, c, mean, mn, mx,
                                 make_float4(pna_std(m2.x * ic), pna_std(m2.y * ic), pna_std(m2.z * ic),
                                             pna_std(m2.w * ic)));
            }
        }
    }
}
extern "C" int gnn_pna_aggregate_msg_fwd(const float* m, const int32_t* rowptr, int64_t n, int towers, int F,
                                         float* A, void* stream) {
    if (n < 0 || towers < 1 || F < 4 || (F & 3)) return GNN_ERR_BAD_ARG;
    if (n == 0) return 0;
    aggregate_msg_fwd_kernel<<<agg_grid(n), AGG_THREADS, 0, gnn_stream(stream)>>>(m, rowptr, n, towers * F, F, A);
    return gnn_launch_status();
}

__global__ void __launch_bounds__(AGG_THREADS)
aggregate_msg_bwd_kernel(const float* __restrict__ gA, const float* __restrict__ m, const float* __restrict__ A,
                         const int* __restrict__ rowptr, int64_t n, int C, int F, float* __restrict__ dm) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * AGG_WARPS + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * AGG_WARPS;
    for (int64_t node = warp0; node < n; node += nwarps) {
        const int beg = rowptr[node], end = rowptr[node + 1];
        const int cnt = end - beg;
        if (cnt == 0) continue;
        const float inv_cnt = 1.0f / (float)cnt;
        for (int c = lane * 4; c < C; c += 128) {
            const int t = c / F, f = c - t * F;
            const int64_t off = node * (4 * (int64_t)C) + (int64_t)t * 4 * F + f;
            const float4 mean = ldg4(A + off), mn = ldg4(A + off + F), mx = ldg4(A + off + 2 * F),
                         sd = ldg4(A + off + 3 * F);
            const float4 g_mean = ldg4(gA + off), g_min = ldg4(gA + off + F), g_max = ldg4(gA + off + 2 * F),
                         g_std = ldg4(gA + off + 3 * F);
            float4 nmin = make_float4(0.f, 0.f, 0.f, 0.f), nmax = nmin;
            for (int s = beg; s < end; ++s) {
                float4 q = ldg4(m + (int64_t)s * C + c);
                nmin.x += q.x == mn.x; nmin.y += q.y == mn.y; nmin.z += q.z == mn.z; nmin.w += q.w == mn.w;
                nmax.x += q.x == mx.x; nmax.y += q.y == mx.y; nmax.z += q.z == mx.z; nmax.w += q.w == mx.w;
            }
            for (int s = beg; s < end; ++s) {
                float4 q = ldg4(m + (int64_t)s * C + c);
                float4 g;
#define GRAD(X)                                                                                           \
    g.X = g_mean.X * inv_cnt + (q.X == mn.X ? g_min.X / nmin.X : 0.f) + (q.X == mx.X ? g_max.X / nmax.X : 0.f) + \
          (sd.X > 0.f ? g_std.X * inv_cnt / sd.X * (q.X - mean.X) : 0.f);
                GRAD(x) GRAD(y) GRAD(z) GRAD(w)
#undef GRAD
                st4(dm + (int64_t)s * C + c, g);
            }
        }
    }
}
extern "C" int gnn_pna_aggregate_msg_bwd(const float* gA, const float* m, const float* A, const int32_t* rowptr,
                                         int64_t n, int towers, int F, float* dm, void* stream) {
    if (n < 0 || towers < 1 || F < 4 || (F & 3)) return GNN_ERR_BAD_ARG;
    if (n == 0) return 0;
    aggregate_msg_bwd_kernel<<<agg_grid(n), AGG_THREADS, 0, gnn_stream(stream)>>>(gA, m, A, rowptr, n, towers * F, F,
                                                                                   dm);
    return gnn_launch_status();
}
