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

#include <float.h>

constexpr int AGG_WARPS = 8;
constexpr int AGG_THREADS = AGG_WARPS * 32;

struct Stat4 {
    float4 mean, m2, mn, mx;
};

__device__ __forceinline__ void welford_update(float& mean, float& m2, float q, float inv_k) {
    float d = q - mean;
    mean = fmaf(d, inv_k, mean);
    m2 = fmaf(d, q - mean, m2);
}
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

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
template <int CHUNKS>
__global__ void __launch_bounds__(AGG_THREADS)
pna_aggregate_fwd_kernel(const float* __restrict__ P, const float* __restrict__ tab,
                         const int* __restrict__ rowptr, const int* __restrict__ col_src,
                         const uint8_t* __restrict__ col_code, int64_t n, int C, int F, float* __restrict__ A) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * AGG_WARPS + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * AGG_WARPS;
    const int64_t ldp = 2 * (int64_t)C;
    for (int64_t node = warp0; node < n; node += nwarps) {
        const int beg = rowptr[node], end = rowptr[node + 1];
        const int cnt = end - beg;
        Stat4 st[CHUNKS];
#pragma unroll
        for (int j = 0; j < CHUNKS; ++j) {
            st[j].mean = make_float4(0.f, 0.f, 0.f, 0.f);
            st[j].m2 = make_float4(0.f, 0.f, 0.f, 0.f);
            st[j].mn = make_float4(FLT_MAX, FLT_MAX, FLT_MAX, FLT_MAX);
            st[j].mx = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX);
        }
        for (int base = beg; base < end; base += 32) {
            const int m = min(32, end - base);
            int my_src = 0, my_code = 0;
            if (lane < m) {
                my_src = __ldg(col_src + base + lane);
                my_code = __ldg(col_code + base + lane);
            }
            for (int k = 0; k < m; ++k) {
                const int src = __shfl_sync(GNN_FULL_MASK, my_src, k);
                const int code = __shfl_sync(GNN_FULL_MASK, my_code, k);
                const float inv_k = 1.0f / (float)(base - beg + k + 1);
                const float* prow = P + (int64_t)src * ldp + C;
                const float* trow = tab + (int64_t)code * C;
#pragma unroll
                for (int j = 0; j < CHUNKS; ++j) {
                    const int c = j * 128 + lane * 4;
                    if (c < C) {
                        float4 q = add4(ldg4(prow + c), ldg4(trow + c));
                        welford_update(st[j].mean.x, st[j].m2.x, q.x, inv_k);
                        welford_update(st[j].mean.y, st[j].m2.y, q.y, inv_k);
                        welford_update(st[j].mean.z, st[j].m2.z, q.z, inv_k);
                        welford_update(st[j].mean.w, st[j].m2.w, q.w, inv_k);
                        st[j].mn = make_float4(fminf(st[j].mn.x, q.x), fminf(st[j].mn.y, q.y),
                                               fminf(st[j].mn.z, q.z), fminf(st[j].mn.w, q.w));
                        st[j].mx = make_float4(fmaxf(st[j].mx.x, q.x), fmaxf(st[j].mx.y, q.y),
                                               fmaxf(st[j].mx.z, q.z), fmaxf(st[j].mx.w, q.w));
                    }
                }
            }
        }
        const float inv_cnt = cnt > 0 ? 1.0f / (float)cnt : 0.0f;
#pragma unroll
        for (int j = 0; j < CHUNKS; ++j) {
            const int c = j * 128 + lane * 4;
            if (c < C) {
                float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                if (cnt == 0) {
                    store_aggregates(A, node, C, F, c, z, z, z, z);
                } else {
                    float4 pd = ldg4(P + node * ldp + c);
                    float4 sd = make_float4(pna_std(st[j].m2.x * inv_cnt), pna_std(st[j].m2.y * inv_cnt),
                                            pna_std(st[j].m2.z * inv_cnt), pna_std(st[j].m2.w * inv_cnt));
                    store_aggregates(A, node, C, F, c, add4(st[j].mean, pd), add4(st[j].mn, pd),
                                     add4(st[j].mx, pd), sd);
                }
            }
        }
    }
}

static int agg_grid(int64_t n) {
    int64_t blocks = gnn_ceil_div(n, AGG_WARPS);
    int64_t cap = (int64_t)gnn_sm_count() * 64;   // grid-stride beyond 64 CTAs/SM worth of work
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

extern "C" int gnn_pna_aggregate_fwd(const float* P, const float* tab, const int32_t* rowptr,
                                     const int32_t* col_src, const uint8_t* col_code, int64_t n, int towers,
                                     int F, float* A, void* stream) {
    if (n < 0 || towers < 1 || F < 4 || (F & 3)) return GNN_ERR_BAD_ARG;
    if (n == 0) return 0;
    const int C = towers * F;
    const int chunks = (C + 127) / 128;
    cudaStream_t st = gnn_stream(stream);
    const int grid = agg_grid(n);
#define LAUNCH_FWD(CH) pna_aggregate_fwd_kernel<CH><<<grid, AGG_THREADS, 0, st>>>(P, tab, rowptr, col_src, col_code, n, C, F, A)
    switch (chunks) {
        case 1: LAUNCH_FWD(1); break;
        case 2: LAUNCH_FWD(2); break;
        case 3: LAUNCH_FWD(3); break;
        case 4: LAUNCH_FWD(4); break;
        case 5: case 6: LAUNCH_FWD(6); break;
        case 7: case 8: LAUNCH_FWD(8); break;
        default: return GNN_ERR_UNSUPPORTED;
    }
#undef LAUNCH_FWD
    return gnn_launch_status();
}

// ------------------------------------------------------------------ backward, factorised message
// Per destination node: recompute the segment statistics of q (same order as forward, so the
// mean/min/max are bitwise those of the forward), count ties, then emit
//   dq_e = g_mean/cnt + g_min*[q_e==min]/n_min + g_max*[q_e==max]/n_max + g_std*(q_e-mean)/(cnt*std)
// per CSR slot, dPdst = (g_mean+g_min+g_max)*[cnt>0] and the per-bond-code sum of dq
// (CTA-private shared table when it fits, flushed once per CTA).
template <bool SMEM_TAB>
__global__ void __launch_bounds__(AGG_THREADS)
pna_aggregate_bwd_kernel(const float* __restrict__ gA, const float* __restrict__ P,
                         const float* __restrict__ tab, const float* __restrict__ A,
                         const int* __restrict__ rowptr, const int* __restrict__ col_src,
                         const uint8_t* __restrict__ col_code, int64_t n, int C, int F,
                         float* __restrict__ dP, float* __restrict__ dq, float* dtab) {
    extern __shared__ float s_dtab[];
    if (SMEM_TAB) {
        for (int k = threadIdx.x; k < GNN_NUM_BOND_CODES * C; k += blockDim.x) s_dtab[k] = 0.f;
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * AGG_WARPS + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * AGG_WARPS;
    const int64_t ldp = 2 * (int64_t)C;
    for (int64_t node = warp0; node < n; node += nwarps) {
        const int beg = rowptr[node], end = rowptr[node + 1];
        const int cnt = end - beg;
        const float inv_cnt = cnt > 0 ? 1.0f / (float)cnt : 0.f;
        // 128 channels at a time (one float4 per lane) keeps the register footprint flat in C
        for (int c = lane * 4; c < C; c += 128) {
            if (cnt == 0) {
                st4(dP + node * ldp + c, make_float4(0.f, 0.f, 0.f, 0.f));
                continue;
            }
            float4 mean = make_float4(0.f, 0.f, 0.f, 0.f);
            float4 mn = make_float4(FLT_MAX, FLT_MAX, FLT_MAX, FLT_MAX);
            float4 mx = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX);
            float4 nmin = make_float4(0.f, 0.f, 0.f, 0.f), nmax = nmin;
            // pass 1: statistics + tie counts (same update order as the forward kernel)
            for (int s = beg; s < end; ++s) {
                const int src = __ldg(col_src + s);
                const int code = __ldg(col_code + s);
                const float inv_k = 1.0f / (float)(s - beg + 1);
                float4 q = add4(ldg4(P + (int64_t)src * ldp + C + c), ldg4(tab + (int64_t)code * C + c));
#define UPD(X)                                                     \
    {                                                              \
        float d = q.X - mean.X;                                    \
        mean.X = fmaf(d, inv_k, mean.X);                           \
        if (q.X < mn.X) { mn.X = q.X; nmin.X = 1.f; }              \
        else if (q.X == mn.X) nmin.X += 1.f;                       \
        if (q.X > mx.X) { mx.X = q.X; nmax.X = 1.f; }              \
        else if (q.X == mx.X) nmax.X += 1.f;                       \
    }
                UPD(x) UPD(y) UPD(z) UPD(w)
#undef UPD
            }
            const int t = c / F, f = c - t * F;
            const int64_t off = node * (4 * (int64_t)C) + (int64_t)t * 4 * F + f;
            const float4 g_mean = ldg4(gA + off), g_min = ldg4(gA + off + F), g_max = ldg4(gA + off + 2 * F),
                         g_std = ldg4(gA + off + 3 * F);
            const float4 sd = ldg4(A + off + 3 * F);
            st4(dP + node * ldp + c, make_float4(g_mean.x + g_min.x + g_max.x, g_mean.y + g_min.y + g_max.y,
                                                 g_mean.z + g_min.z + g_max.z, g_mean.w + g_min.w + g_max.w));
            float4 cm, cmin, cmax, cs;
#define COEF(X)                                              \
    cm.X = g_mean.X * inv_cnt;                               \
    cmin.X = g_min.X / nmin.X;                               \
    cmax.X = g_max.X / nmax.X;                               \
    cs.X = sd.X > 0.f ? g_std.X * inv_cnt / sd.X : 0.f;
            COEF(x) COEF(y) COEF(z) COEF(w)
#undef COEF
            // pass 2: per-slot gradient (rows are L1/L2 hits from pass 1)
            for (int s = beg; s < end; ++s) {
                const int src = __ldg(col_src + s);
                const int code = __ldg(col_code + s);
                float4 q = add4(ldg4(P + (int64_t)src * ldp + C + c), ldg4(tab + (int64_t)code * C + c));
                float4 g;
#define GRAD(X) \
    g.X = cm.X + (q.X == mn.X ? cmin.X : 0.f) + (q.X == mx.X ? cmax.X : 0.f) + cs.X * (q.X - mean.X);
                GRAD(x) GRAD(y) GRAD(z) GRAD(w)
#undef GRAD
                st4(dq + (int64_t)s * C + c, g);
                float* dt = (SMEM_TAB ? s_dtab : dtab) + code * C + c;
                atomicAdd(dt, g.x);
                atomicAdd(dt + 1, g.y);
                atomicAdd(dt + 2, g.z);
                atomicAdd(dt + 3, g.w);
            }
        }
    }
    if (SMEM_TAB) {
        __syncthreads();
        for (int k = threadIdx.x; k < GNN_NUM_BOND_CODES * C; k += blockDim.x) {
            float v = s_dtab[k];
            if (v != 0.f) atomicAdd(dtab + k, v);
        }
    }
}

extern "C" int gnn_pna_aggregate_bwd(const float* gA, const float* P, const float* tab, const float* A,
                                     const int32_t* rowptr, const int32_t* col_src, const uint8_t* col_code,
                                     int64_t n, int towers, int F, float* dP, float* dq, float* dtab_accum,
                                     void* stream) {
    if (n < 0 || towers < 1 || F < 4 || (F & 3)) return GNN_ERR_BAD_ARG;
    if (n == 0) return 0;
    const int C = towers * F;
    cudaStream_t st = gnn_stream(stream);
    const size_t smem = (size_t)GNN_NUM_BOND_CODES * C * sizeof(float);
    int64_t blocks = gnn_ceil_div(n, AGG_WARPS);
    if (smem <= 200 * 1024) {
        auto kern = pna_aggregate_bwd_kernel<true>;
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
        }
        int per_sm = (int)((220 * 1024) / (smem + 1024));
        if (per_sm < 1) per_sm = 1;
        if (per_sm > 4) per_sm = 4;
        int64_t cap = (int64_t)gnn_sm_count() * per_sm;
        if (blocks > cap) blocks = cap;
        kern<<<(unsigned)blocks, AGG_THREADS, smem, st>>>(gA, P, tab, A, rowptr, col_src, col_code, n, C, F, dP, dq,
                                                          dtab_accum);
    } else {
        int64_t cap = (int64_t)gnn_sm_count() * 8;
        if (blocks > cap) blocks = cap;
        pna_aggregate_bwd_kernel<false><<<(unsigned)blocks, AGG_THREADS, 0, st>>>(gA, P, tab, A, rowptr, col_src,
                                                                                  col_code, n, C, F, dP, dq, dtab_accum);
    }
    return gnn_launch_status();
}

// ------------------------------------------------------------------ gather dq by source node
__global__ void __launch_bounds__(AGG_THREADS)
scatter_src_bwd_kernel(const float* __restrict__ dq, const int* __restrict__ srcptr,
                       const int* __restrict__ src_slot, int64_t n, int C, float* __restrict__ dP) {
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
            st4(dP + node * (2 * (int64_t)C) + C + c, acc);
        }
    }
}

extern "C" int gnn_scatter_src_bwd(const float* dq, const int32_t* srcptr, const int32_t* src_slot, int64_t n,
                                   int C, float* dP, void* stream) {
    if (n < 0 || C < 4 || (C & 3)) return GNN_ERR_BAD_ARG;
    if (n == 0) return 0;
    scatter_src_bwd_kernel<<<agg_grid(n), AGG_THREADS, 0, gnn_stream(stream)>>>(dq, srcptr, src_slot, n, C, dP);
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
