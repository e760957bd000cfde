// fp32 SIMT dense stages (the exact-fp32 parity path).
// Replace torch.nn.Linear / PyG Linear (addmm) inside PNAConv (pre_nns, post_nns, lin) and the
// MLP head (reference gnnepcsaft/train/models.py:69-103,128,134), plus their backward.
//
// One register-tiled kernel template: C[M,N] = sum_k Aop(m,k) * Bop(k,n), operands fetched by
// small functors so that (a) weight slices are used in place (no packing), (b) the
// [x | A | amp*A | att*A] concatenation of DegreeScalerAggregation + PNAConv.forward is
// *virtual* — the [N,T,12F] / [N,T,13F] tensors never exist.
#include "common.cuh"

// ------------------------------------------------------------------ operand functors
// element (r, c) of a row-major matrix with leading dimension ld
struct RowMajor {
    const float* p;
    int64_t ld, rows, cols;
    int64_t group_stride;
    __device__ __forceinline__ void set_group(int g) { p += (int64_t)g * group_stride; }
    __device__ __forceinline__ float get(int64_t r, int64_t c) const {
        return (r < rows && c < cols) ? __ldg(p + r * ld + c) : 0.f;
    }
};
// element (r, c) = M[c, r]
struct Transposed {
    const float* p;
    int64_t ld, rows, cols;   // rows/cols of the *transposed view*
    int64_t group_stride;
    __device__ __forceinline__ void set_group(int g) { p += (int64_t)g * group_stride; }
    __device__ __forceinline__ float get(int64_t r, int64_t c) const {
        return (r < rows && c < cols) ? __ldg(p + c * ld + r) : 0.f;
    }
};
// virtual [x | A_t | amp*A_t | att*A_t]  (node i, column k in [0,13F)); group = tower
struct PostConcat {
    const float* x;       // [N,F]
    const float* A;       // [N,T,4F]
    const float* scal;    // [N,2] amp, att
    int64_t n;
    int F, towers, tower;
    __device__ __forceinline__ void set_group(int g) { tower = g; }
    __device__ __forceinline__ float get_node_col(int64_t i, int64_t k) const {
        if (i >= n || k >= 13 * (int64_t)F) return 0.f;
        if (k < F) return __ldg(x + i * F + k);
        int kk = (int)k - F;
        int s = kk / (4 * F), r = kk - s * 4 * F;
        float v = __ldg(A + (i * towers + tower) * (4 * (int64_t)F) + r);
        if (s == 0) return v;
        return v * __ldg(scal + 2 * i + (s - 1));
    }
};
struct PostConcatA : PostConcat {   // as A operand (m = node, k = column)
    __device__ __forceinline__ float get(int64_t m, int64_t k) const { return get_node_col(m, k); }
};
struct PostConcatB : PostConcat {   // as B operand of wgrad (k = node, n = column)
    __device__ __forceinline__ float get(int64_t k, int64_t n_) const { return get_node_col(k, n_); }
};
// virtual [dO_t | amp*dO_t | att*dO_t] (node i, column k in [0,3*F_out)); group = tower
struct ScaledGrad {
    const float* dO;      // [N, T*F_out]
    const float* scal;
    int64_t n;
    int F_out, towers, tower;
    __device__ __forceinline__ void set_group(int g) { tower = g; }
    __device__ __forceinline__ float get(int64_t i, int64_t k) const {
        if (i >= n || k >= 3 * (int64_t)F_out) return 0.f;
        int s = (int)k / F_out, o = (int)k - s * F_out;
        float v = __ldg(dO + i * ((int64_t)towers * F_out) + tower * F_out + o);
        if (s == 0) return v;
        return v * __ldg(scal + 2 * i + (s - 1));
    }
};
// B operand of the post dgrad: (k = s*F_out + o, n = r) -> Wpost_t[o, F + s*4F + r]
struct PostWeightSlices {
    const float* w;       // tower 0, [F_out, 13F]
    int64_t tower_stride;
    int F, F_out;
    __device__ __forceinline__ void set_group(int g) { w += (int64_t)g * tower_stride; }
    __device__ __forceinline__ float get(int64_t k, int64_t n_) const {
        if (k >= 3 * (int64_t)F_out || n_ >= 4 * (int64_t)F) return 0.f;
        int s = (int)k / F_out, o = (int)k - s * F_out;
        return __ldg(w + (int64_t)o * 13 * F + F + (int64_t)s * 4 * F + n_);
    }
};

// ------------------------------------------------------------------ epilogues
struct EpiStore {   // Y[m, n] = act(acc + bias[n])   (group shifts the output column and bias)
    float* y;
    int64_t ldy;
    const float* bias;
    int64_t y_group_stride, bias_group_stride;
    int act, accumulate;
    __device__ __forceinline__ void set_group(int g) {
        y += (int64_t)g * y_group_stride;
        if (bias) bias += (int64_t)g * bias_group_stride;
    }
    __device__ __forceinline__ void put(int64_t m, int64_t n_, float v) const {
        if (bias) v += __ldg(bias + n_);
        if (act) v = fmaxf(v, 0.f);
        float* dst = y + m * ldy + n_;
        if (accumulate) v += *dst;
        *dst = v;
    }
};
struct EpiAtomic {  // split-K weight gradient: dW[m, n] += acc
    float* y;
    int64_t ldy;
    int64_t y_group_stride;
    __device__ __forceinline__ void set_group(int g) { y += (int64_t)g * y_group_stride; }
    __device__ __forceinline__ void put(int64_t m, int64_t n_, float v) const { atomicAdd(y + m * ldy + n_, v); }
};

// ------------------------------------------------------------------ the kernel
template <int BM, int BN, int BK, int TM, int TN, bool A_KC, bool B_KC, class AL, class BL, class EP>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
gemm_kernel(AL al, BL bl, EP ep, int64_t M, int64_t N, int64_t K, int64_t k_chunk, int groups,
            float* rowsum_out, int64_t rowsum_group_stride) {
    constexpr int NT = (BM / TM) * (BN / TN);
    constexpr int A_PER = BM * BK / NT, B_PER = BN * BK / NT;
    static_assert(BM * BK % NT == 0 && BN * BK % NT == 0, "tile/threads mismatch");
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];

    const int tid = threadIdx.x;
    const int g = blockIdx.z % groups, split = blockIdx.z / groups;
    al.set_group(g);
    bl.set_group(g);
    ep.set_group(g);
    const int64_t m0 = (int64_t)blockIdx.x * BM, n0 = (int64_t)blockIdx.y * BN;
    const int64_t k_begin = (int64_t)split * k_chunk;
    const int64_t k_end = min(K, k_begin + k_chunk);
    const int tx = tid % (BN / TN), ty = tid / (BN / TN);

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    float ra[A_PER], rb[B_PER];
    float rowsum = 0.f;   // colsum(dY) for bias gradients (wgrad, first N tile only)

    auto fetch = [&](int64_t kt) {
#pragma unroll
        for (int i = 0; i < A_PER; ++i) {
            int idx = tid + i * NT;
            int kk = A_KC ? idx % BK : idx / BM;
            int mm = A_KC ? idx / BK : idx % BM;
            int64_t k = kt + kk;
            ra[i] = k < k_end ? al.get(m0 + mm, k) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < B_PER; ++i) {
            int idx = tid + i * NT;
            int kk = B_KC ? idx % BK : idx / BN;
            int nn = B_KC ? idx / BK : idx % BN;
            int64_t k = kt + kk;
            rb[i] = k < k_end ? bl.get(k, n0 + nn) : 0.f;
        }
    };
    auto stash = [&]() {
#pragma unroll
        for (int i = 0; i < A_PER; ++i) {
            int idx = tid + i * NT;
            int kk = A_KC ? idx % BK : idx / BM;
            int mm = A_KC ? idx / BK : idx % BM;
            As[kk][mm] = ra[i];
        }
#pragma unroll
        for (int i = 0; i < B_PER; ++i) {
            int idx = tid + i * NT;
            int kk = B_KC ? idx % BK : idx / BN;
            int nn = B_KC ? idx / BK : idx % BN;
            Bs[kk][nn] = rb[i];
        }
    };

    if (k_begin < k_end) {
        fetch(k_begin);
        for (int64_t kt = k_begin; kt < k_end; kt += BK) {
            stash();
            __syncthreads();
            if (kt + BK < k_end) fetch(kt + BK);
            if (rowsum_out != nullptr && blockIdx.y == 0 && tid < BM) {
#pragma unroll
                for (int kk = 0; kk < BK; ++kk) rowsum += As[kk][tid];
            }
#pragma unroll
            for (int kk = 0; kk < BK; ++kk) {
                float a[TM], b[TN];
                static_assert(TM % 4 == 0 && TN % 2 == 0, "micro-tile must allow vector smem reads");
#pragma unroll
                for (int i = 0; i < TM; i += 4) {
                    float4 v = *reinterpret_cast<const float4*>(&As[kk][ty * TM + i]);
                    a[i] = v.x; a[i + 1] = v.y; a[i + 2] = v.z; a[i + 3] = v.w;
                }
#pragma unroll
                for (int j = 0; j < TN; j += 2) {
                    float2 v = *reinterpret_cast<const float2*>(&Bs[kk][tx * TN + j]);
                    b[j] = v.x; b[j + 1] = v.y;
                }
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int64_t m = m0 + ty * TM + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int64_t n_ = n0 + tx * TN + j;
            if (n_ < N) ep.put(m, n_, acc[i][j]);
        }
    }
    if (rowsum_out != nullptr && blockIdx.y == 0 && tid < BM && m0 + tid < M)
        atomicAdd(rowsum_out + (int64_t)g * rowsum_group_stride + m0 + tid, rowsum);
}

// forward / dgrad style launch (M large): pick BN by N
template <bool A_KC, bool B_KC, class AL, class BL, class EP>
static int launch_rowtiled(AL al, BL bl, EP ep, int64_t M, int64_t N, int64_t K, int groups, cudaStream_t st) {
    if (M <= 0 || N <= 0) return 0;
    if (N > 32) {
        dim3 grid((unsigned)gnn_ceil_div(M, 128), (unsigned)gnn_ceil_div(N, 64), (unsigned)groups);
        gemm_kernel<128, 64, 16, 8, 4, A_KC, B_KC><<<grid, 256, 0, st>>>(al, bl, ep, M, N, K, K, groups, nullptr, 0);
    } else {
        dim3 grid((unsigned)gnn_ceil_div(M, 128), 1, (unsigned)groups);
        gemm_kernel<128, 32, 16, 8, 2, A_KC, B_KC><<<grid, 256, 0, st>>>(al, bl, ep, M, N, K, K, groups, nullptr, 0);
    }
    return gnn_launch_status();
}

// weight-gradient style launch (M,N small, K = rows huge): split K over the grid, atomics
template <bool A_KC, bool B_KC, class AL, class BL>
static int launch_splitk(AL al, BL bl, EpiAtomic ep, int64_t M, int64_t N, int64_t K, int groups, float* rowsum,
                         int64_t rowsum_group_stride, cudaStream_t st) {
    if (M <= 0 || N <= 0 || K <= 0) return 0;
    const int64_t tiles = gnn_ceil_div(M, 64) * gnn_ceil_div(N, 64) * groups;
    int64_t target = (int64_t)gnn_sm_count() * 4;
    int64_t splits = gnn_ceil_div(target, tiles);
    int64_t max_splits = gnn_ceil_div(K, 256);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    if (splits * groups > 65535) splits = 65535 / groups;
    int64_t k_chunk = gnn_align_up(gnn_ceil_div(K, splits), 16);
    splits = gnn_ceil_div(K, k_chunk);
    dim3 grid((unsigned)gnn_ceil_div(M, 64), (unsigned)gnn_ceil_div(N, 64), (unsigned)(splits * groups));
    gemm_kernel<64, 64, 16, 4, 4, A_KC, B_KC><<<grid, 256, 0, st>>>(al, bl, ep, M, N, K, k_chunk, groups, rowsum,
                                                                     rowsum_group_stride);
    return gnn_launch_status();
}

// ------------------------------------------------------------------ plain Linear
extern "C" int gnn_linear_fwd(const float* X, int64_t ldx, const float* W, int64_t ldw, const float* bias, int64_t M,
                              int N, int K, int act, float* Y, int64_t ldy, void* stream) {
    if (M < 0 || N < 1 || K < 1) return GNN_ERR_BAD_ARG;
    RowMajor al{X, ldx, M, K, 0};
    Transposed bl{W, ldw, K, N, 0};
    EpiStore ep{Y, ldy, bias, 0, 0, act, 0};
    return launch_rowtiled<true, true>(al, bl, ep, M, N, K, 1, gnn_stream(stream));
}

extern "C" int gnn_linear_dgrad(const float* dY, int64_t lddy, const float* W, int64_t ldw, int64_t M, int N, int K,
                                int accumulate, float* dX, int64_t lddx, void* stream) {
    if (M < 0 || N < 1 || K < 1) return GNN_ERR_BAD_ARG;
    RowMajor al{dY, lddy, M, N, 0};
    RowMajor bl{W, ldw, N, K, 0};
    EpiStore ep{dX, lddx, nullptr, 0, 0, 0, accumulate};
    return launch_rowtiled<true, false>(al, bl, ep, M, K, N, 1, gnn_stream(stream));
}

extern "C" int gnn_linear_wgrad(const float* dY, int64_t lddy, const float* X, int64_t ldx, int64_t M, int N, int K,
                                float* dW, int64_t lddw, float* db, void* stream) {
    if (M < 0 || N < 1 || K < 1) return GNN_ERR_BAD_ARG;
    Transposed al{dY, lddy, N, M, 0};    // (o, i) -> dY[i, o]
    RowMajor bl{X, ldx, M, K, 0};        // (i, k) -> X[i, k]
    EpiAtomic ep{dW, lddw, 0};
    return launch_splitk<false, false>(al, bl, ep, N, K, M, 1, db, 0, gnn_stream(stream));
}

// ------------------------------------------------------------------ PNA post stage
__global__ void pna_scalers_kernel(const int* __restrict__ rowptr, const float* __restrict__ avg_deg_log, int64_t n,
                                   float* __restrict__ scal) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float amp, att;
    pna_scalers(rowptr[i + 1] - rowptr[i], __ldg(avg_deg_log), amp, att);
    scal[2 * i] = amp;
    scal[2 * i + 1] = att;
}
extern "C" int gnn_pna_scalers(const int32_t* rowptr, const float* avg_deg_log, int64_t n, float* scal, void* stream) {
    if (n <= 0) return n == 0 ? 0 : GNN_ERR_BAD_ARG;
    pna_scalers_kernel<<<(unsigned)gnn_ceil_div(n, 256), 256, 0, gnn_stream(stream)>>>(rowptr, avg_deg_log, n, scal);
    return gnn_launch_status();
}

extern "C" int gnn_pna_post_fwd(const float* x, const float* A, const float* scal, const float* wpost,
                                int64_t wpost_tower_stride, const float* bpost, int64_t bpost_tower_stride, int64_t n,
                                int F, int F_out, int towers, int act, float* O, void* stream) {
    if (n < 0 || F < 1 || F_out < 1 || towers < 1) return GNN_ERR_BAD_ARG;
    PostConcatA al;
    al.x = x; al.A = A; al.scal = scal; al.n = n; al.F = F; al.towers = towers; al.tower = 0;
    Transposed bl{wpost, 13 * (int64_t)F, 13 * (int64_t)F, F_out, wpost_tower_stride};
    EpiStore ep{O, (int64_t)towers * F_out, bpost, F_out, bpost_tower_stride, act, 0};
    return launch_rowtiled<true, true>(al, bl, ep, n, F_out, 13 * (int64_t)F, towers, gnn_stream(stream));
}

extern "C" int gnn_pna_post_dgrad(const float* dO, const float* scal, const float* wpost, int64_t wpost_tower_stride,
                                  int64_t n, int F, int F_out, int towers, float* dx, float* gA, void* stream) {
    if (n < 0 || F < 1 || F_out < 1 || towers < 1) return GNN_ERR_BAD_ARG;
    cudaStream_t st = gnn_stream(stream);
    // gA[:, t, :] = [dO_t | amp dO_t | att dO_t] . Wslices_t      (K = 3 F_out, N = 4F)
    ScaledGrad al{dO, scal, n, F_out, towers, 0};
    PostWeightSlices bl{wpost, wpost_tower_stride, F, F_out};
    EpiStore ep{gA, (int64_t)towers * 4 * F, nullptr, 4 * (int64_t)F, 0, 0, 0};
    int rc = launch_rowtiled<true, false>(al, bl, ep, n, 4 * (int64_t)F, 3 * (int64_t)F_out, towers, st);
    if (rc) return rc;
    // dx += sum_t dO_t . Wpost_t[:, :F]   (towers accumulate into the same dx: sequential launches)
    for (int t = 0; t < towers; ++t) {
        RowMajor a2{dO + t * F_out, (int64_t)towers * F_out, n, F_out, 0};
        RowMajor b2{wpost + t * wpost_tower_stride, 13 * (int64_t)F, F_out, F, 0};
        EpiStore e2{dx, F, nullptr, 0, 0, 0, 1};
        rc = launch_rowtiled<true, false>(a2, b2, e2, n, F, F_out, 1, st);
        if (rc) return rc;
    }
    return 0;
}

extern "C" int gnn_pna_post_wgrad(const float* dO, const float* x, const float* A, const float* scal, int64_t n, int F,
                                  int F_out, int towers, float* dwpost, int64_t wpost_tower_stride, float* dbpost,
                                  int64_t bpost_tower_stride, void* stream) {
    if (n < 0 || F < 1 || F_out < 1 || towers < 1) return GNN_ERR_BAD_ARG;
    Transposed al{dO, (int64_t)towers * F_out, F_out, n, F_out};   // (o, i) -> dO[i, t*F_out + o]
    PostConcatB bl;
    bl.x = x; bl.A = A; bl.scal = scal; bl.n = n; bl.F = F; bl.towers = towers; bl.tower = 0;
    EpiAtomic ep{dwpost, 13 * (int64_t)F, wpost_tower_stride};
    return launch_splitk<false, false>(al, bl, ep, F_out, 13 * (int64_t)F, n, towers, dbpost, bpost_tower_stride,
                                       gnn_stream(stream));
}
