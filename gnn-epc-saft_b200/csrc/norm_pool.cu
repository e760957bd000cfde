// BatchNorm(+ReLU+skip), segment pooling, MAPE loss and the fused AdamW-amsgrad step.
// Replace PyG BatchNorm / torch BatchNorm1d, F.relu, the skip add, global_add_pool, torchmetrics
// MAPE and torch.optim.AdamW + CosineAnnealingWarmRestarts
// (reference gnnepcsaft/train/models.py:82,128-133,162-188,194).  All HBM-bound elementwise /
// column-reduction work.
#include "common.cuh"

#include <math.h>

constexpr int RED_TX = 32, RED_TY = 8;          // column-reduction block: 32 columns x 8 row phases
constexpr int RED_ROWS_PER_CTA = 128;

static int red_ctas(int64_t rows) {
    int64_t c = gnn_ceil_div(rows, RED_ROWS_PER_CTA);
    // two CTAs per SM: the finalize kernels read every partial, so few, fat parts (the finalize pass with 400-800
    // parts cost as much as the reduction itself: 14 us of a 28 us BatchNorm statistics call)
    int64_t cap = (int64_t)gnn_sm_count() * 2;
    if (c > cap) c = cap;
    if (c < 1) c = 1;
    return (int)c;
}

extern "C" int64_t gnn_bn_workspace_bytes(int64_t rows, int C) {
    return gnn_align_up((int64_t)red_ctas(rows) * 2 * C * (int64_t)sizeof(double) + 2 * C * (int64_t)sizeof(float), 256);
}

__device__ __forceinline__ int64_t ceil_div_dev(int64_t a, int64_t b) { return (a + b - 1) / b; }
// Padded batches (shape buckets of the CUDA-graph-captured step): only the first *valid_rows rows are real; the
// statistics, the backward sums and their 1/n use that count, and the gradient of the padding rows is zero.  The count
// lives in device memory so that one captured graph serves every batch of a bucket.
__device__ __forceinline__ int64_t bn_valid(int64_t rows, const int* valid_rows) {
    if (valid_rows == nullptr) return rows;
    const int64_t v = (int64_t)__ldg(valid_rows);
    return v < rows ? (v < 0 ? 0 : v) : rows;
}

// mode 0: (y, y*y)            -> batch statistics
// mode 1: (dz, dz*yhat)       -> BN backward sums, dz = dOut * [relu ? z>0 : 1]
template <int MODE>
__global__ void __launch_bounds__(RED_TX * RED_TY)
bn_partial_kernel(const float* __restrict__ Y, const float* __restrict__ dOut, const float* __restrict__ stats,
                  const float* __restrict__ gamma, const float* __restrict__ beta, int64_t rows, int C, int relu,
                  double* __restrict__ partial, const int* __restrict__ valid_rows) {
    __shared__ double s0[RED_TY][RED_TX + 1], s1[RED_TY][RED_TX + 1];
    const int64_t chunk = ceil_div_dev(rows, gridDim.x);
    const int64_t r0 = (int64_t)blockIdx.x * chunk;
    const int64_t r1 = min(bn_valid(rows, valid_rows), r0 + chunk);
    for (int cb = 0; cb < C; cb += RED_TX) {
        const int c = cb + threadIdx.x;
        double a0 = 0.0, a1 = 0.0;
        if (c < C) {
            float mu = 0.f, rs = 0.f, ga = 0.f, be = 0.f;
            if (MODE == 1) {
                mu = __ldg(stats + c);
                rs = __ldg(stats + C + c);
                ga = __ldg(gamma + c);
                be = __ldg(beta + c);
            }
            for (int64_t r = r0 + threadIdx.y; r < r1; r += RED_TY) {
                float y = __ldg(Y + r * C + c);
                if (MODE == 0) {
                    a0 += (double)y;
                    a1 += (double)y * (double)y;
                } else {
                    float yh = (y - mu) * rs;
                    float dz = __ldg(dOut + r * C + c);
                    if (relu && !(fmaf(ga, yh, be) > 0.f)) dz = 0.f;
                    a0 += (double)dz;
                    a1 += (double)dz * (double)yh;
                }
            }
        }
        s0[threadIdx.y][threadIdx.x] = a0;
        s1[threadIdx.y][threadIdx.x] = a1;
        __syncthreads();
        if (threadIdx.y == 0 && c < C) {
            double t0 = 0.0, t1 = 0.0;
#pragma unroll
            for (int k = 0; k < RED_TY; ++k) {
                t0 += s0[k][threadIdx.x];
                t1 += s1[k][threadIdx.x];
            }
            partial[((int64_t)blockIdx.x * 2 + 0) * C + c] = t0;
            partial[((int64_t)blockIdx.x * 2 + 1) * C + c] = t1;
        }
        __syncthreads();
    }
}

// vectorised variant for C in {4, 8, ..., 1024} with 256 % (C/4) == 0: 16-byte loads, TX = C/4 threads per row
constexpr int BN4_THREADS = 512;
template <int MODE>
__global__ void __launch_bounds__(BN4_THREADS)
bn_partial4_kernel(const float* __restrict__ Y, const float* __restrict__ dOut, const float* __restrict__ stats,
                   const float* __restrict__ gamma, const float* __restrict__ beta, int64_t rows, int C, int relu,
                   double* __restrict__ partial, const int* __restrict__ valid_rows) {
    extern __shared__ double sred[];                 // [BN4_THREADS][8]
    const int TX = C >> 2, TY = BN4_THREADS / TX;
    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    const int c = tx * 4;
    const int64_t chunk = ceil_div_dev(rows, gridDim.x);
    const int64_t r0 = (int64_t)blockIdx.x * chunk, r1 = min(bn_valid(rows, valid_rows), r0 + chunk);
    double a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    float4 mu = make_float4(0.f, 0.f, 0.f, 0.f), rs = mu, ga = mu, be = mu;
    if (MODE == 1) {
        mu = ldg4(stats + c);
        rs = ldg4(stats + C + c);
        ga = ldg4(gamma + c);
        be = ldg4(beta + c);
    }
    // four rows per trip: the (up to eight) 16-byte loads are independent and issued together, their contributions are
    // summed in fp32 (4 terms) and only then promoted to the fp64 accumulators
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t r = r0 + ty; r < r1; r += 4 * TY) {
        float4 y[4], d[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t ru = r + (int64_t)u * TY;
            const bool ok = ru < r1;
            y[u] = ok ? ldg4(Y + ru * C + c) : (MODE == 0 ? z4 : mu);        // padding rows contribute zero
            if (MODE == 1) d[u] = ok ? ldg4(dOut + ru * C + c) : z4;
        }
        float4 p = z4, q = z4;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (MODE == 0) {
                p = make_float4(p.x + y[u].x, p.y + y[u].y, p.z + y[u].z, p.w + y[u].w);
                q = make_float4(fmaf(y[u].x, y[u].x, q.x), fmaf(y[u].y, y[u].y, q.y), fmaf(y[u].z, y[u].z, q.z),
                                fmaf(y[u].w, y[u].w, q.w));
            } else {
                float4 dz = d[u];
                const float4 yh = make_float4((y[u].x - mu.x) * rs.x, (y[u].y - mu.y) * rs.y, (y[u].z - mu.z) * rs.z,
                                              (y[u].w - mu.w) * rs.w);
                if (relu) {
                    if (!(fmaf(ga.x, yh.x, be.x) > 0.f)) dz.x = 0.f;
                    if (!(fmaf(ga.y, yh.y, be.y) > 0.f)) dz.y = 0.f;
                    if (!(fmaf(ga.z, yh.z, be.z) > 0.f)) dz.z = 0.f;
                    if (!(fmaf(ga.w, yh.w, be.w) > 0.f)) dz.w = 0.f;
                }
                p = make_float4(p.x + dz.x, p.y + dz.y, p.z + dz.z, p.w + dz.w);
                q = make_float4(fmaf(dz.x, yh.x, q.x), fmaf(dz.y, yh.y, q.y), fmaf(dz.z, yh.z, q.z), fmaf(dz.w, yh.w, q.w));
            }
        }
        a[0] += p.x; a[1] += p.y; a[2] += p.z; a[3] += p.w;
        a[4] += q.x; a[5] += q.y; a[6] += q.z; a[7] += q.w;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) sred[threadIdx.x * 8 + k] = a[k];
    __syncthreads();
    for (int idx = threadIdx.x; idx < TX * 8; idx += BN4_THREADS) {     // one thread per (column quad, sum) pair
        const int tx2 = idx >> 3, k = idx & 7;
        double t = 0.0;
        for (int j = 0; j < TY; ++j) t += sred[(j * TX + tx2) * 8 + k];
        partial[((int64_t)blockIdx.x * 2 + (k >> 2)) * C + tx2 * 4 + (k & 3)] = t;
    }
}

static bool bn_vec_ok(int C) { return (C & 3) == 0 && C >= 4 && C <= 1024 && (256 % (C >> 2)) == 0; }

// Sum the per-CTA partials of one column; launched with dim3(32, FIN_TY): threadIdx.y strides the parts.  Returns
// the totals in thread y == 0 (other threads get garbage and must not write).
constexpr int FIN_TY = 32;
__device__ __forceinline__ void reduce_parts(const double* __restrict__ partial, int nparts, int C, int c, double& s,
                                             double& ss) {
    __shared__ double r0[FIN_TY][33], r1[FIN_TY][33];
    double a = 0.0, b = 0.0;
    if (c < C)
#pragma unroll 10
        for (int p = threadIdx.y; p < nparts; p += FIN_TY) {
            a += partial[((int64_t)p * 2 + 0) * C + c];
            b += partial[((int64_t)p * 2 + 1) * C + c];
        }
    r0[threadIdx.y][threadIdx.x] = a;
    r1[threadIdx.y][threadIdx.x] = b;
    __syncthreads();
    s = ss = 0.0;
    if (threadIdx.y == 0)
#pragma unroll
        for (int k = 0; k < FIN_TY; ++k) {
            s += r0[k][threadIdx.x];
            ss += r1[k][threadIdx.x];
        }
}

__global__ void bn_stats_finalize_kernel(const double* __restrict__ partial, int nparts, int64_t rows, int C,
                                         int training, float eps, float momentum, float* running_mean,
                                         float* running_var, int64_t* num_batches_tracked, float* stats,
                                         const int* __restrict__ valid_rows) {
    rows = bn_valid(rows, valid_rows);
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    double s = 0.0, ss = 0.0;
    if (training) reduce_parts(partial, nparts, C, c, s, ss);
    if (threadIdx.y != 0) return;
    if (c == 0 && training && num_batches_tracked) *num_batches_tracked += 1;
    if (c >= C) return;
    if (!training) {
        stats[c] = running_mean[c];
        stats[C + c] = (float)(1.0 / sqrt((double)running_var[c] + (double)eps));
        return;
    }
    double mean = s / (double)(rows > 0 ? rows : 1);
    double var = ss / (double)(rows > 0 ? rows : 1) - mean * mean;
    if (var < 0.0) var = 0.0;
    stats[c] = (float)mean;
    stats[C + c] = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean) {
        double unbiased = rows > 1 ? var * (double)rows / (double)(rows - 1) : var;
        running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * mean);
        running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unbiased);
    }
}

extern "C" int gnn_bn_stats(const float* Y, int64_t rows, int C, int training, float eps, float momentum,
                            float* running_mean, float* running_var, int64_t* num_batches_tracked, float* stats,
                            const int32_t* valid_rows, void* workspace, int64_t workspace_bytes, void* stream) {
    if (rows < 0 || C < 1) return GNN_ERR_BAD_ARG;
    if (!training && (!running_mean || !running_var)) return GNN_ERR_BAD_ARG;
    if (training && workspace_bytes < gnn_bn_workspace_bytes(rows, C)) return GNN_ERR_WORKSPACE;
    if (training && rows < 1) return GNN_ERR_BAD_ARG;
    cudaStream_t st = gnn_stream(stream);
    double* partial = reinterpret_cast<double*>(workspace);
    int parts = red_ctas(rows);
    if (training) {
        if (bn_vec_ok(C))
            bn_partial4_kernel<0><<<parts, BN4_THREADS, BN4_THREADS * 8 * sizeof(double), st>>>(Y, nullptr, nullptr, nullptr, nullptr,
                                                                                rows, C, 0, partial, valid_rows);
        else
            bn_partial_kernel<0><<<parts, dim3(RED_TX, RED_TY), 0, st>>>(Y, nullptr, nullptr, nullptr, nullptr, rows, C,
                                                                          0, partial, valid_rows);
    }
    bn_stats_finalize_kernel<<<(C + 31) / 32, dim3(32, FIN_TY), 0, st>>>(partial, parts, rows, C, training, eps, momentum,
                                                              running_mean, running_var, num_batches_tracked, stats,
                                                              valid_rows);
    return gnn_launch_status();
}

__global__ void bn_act_fwd_kernel(const float* __restrict__ Y, const float* __restrict__ stats,
                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                  const float* __restrict__ skip, int64_t total, int C, int relu,
                                  float* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        int c = (int)(i % C);
        float v = fmaf(__ldg(gamma + c), (__ldg(Y + i) - __ldg(stats + c)) * __ldg(stats + C + c), __ldg(beta + c));
        if (relu) v = fmaxf(v, 0.f);
        if (skip) v += __ldg(skip + i);
        out[i] = v;
    }
}
// float4 variant (C % 4 == 0)
__global__ void bn_act_fwd4_kernel(const float* __restrict__ Y, const float* __restrict__ stats,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   const float* __restrict__ skip, int64_t total4, int C, int relu,
                                   float* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int C4 = C >> 2;
    for (; i < total4; i += stride) {
        int c = (int)(i % C4) * 4;
        float4 y = ldg4(Y + 4 * i), mu = ldg4(stats + c), rs = ldg4(stats + C + c), ga = ldg4(gamma + c),
               be = ldg4(beta + c);
        float4 v = make_float4(fmaf(ga.x, (y.x - mu.x) * rs.x, be.x), fmaf(ga.y, (y.y - mu.y) * rs.y, be.y),
                               fmaf(ga.z, (y.z - mu.z) * rs.z, be.z), fmaf(ga.w, (y.w - mu.w) * rs.w, be.w));
        if (relu) v = make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
        if (skip) {
            float4 s = ldg4(skip + 4 * i);
            v = make_float4(v.x + s.x, v.y + s.y, v.z + s.z, v.w + s.w);
        }
        st4(out + 4 * i, v);
    }
}

static int ew_grid(int64_t work, int threads) {
    int64_t b = gnn_ceil_div(work, threads);
    int64_t cap = (int64_t)gnn_sm_count() * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

extern "C" int gnn_bn_act_fwd(const float* Y, const float* stats, const float* gamma, const float* beta,
                              const float* skip, int64_t rows, int C, int relu, float* out, void* stream) {
    if (rows < 0 || C < 1) return GNN_ERR_BAD_ARG;
    if (rows == 0) return 0;
    cudaStream_t st = gnn_stream(stream);
    int64_t total = rows * C;
    if ((C & 3) == 0)
        bn_act_fwd4_kernel<<<ew_grid(total / 4, 256), 256, 0, st>>>(Y, stats, gamma, beta, skip, total / 4, C, relu, out);
    else
        bn_act_fwd_kernel<<<ew_grid(total, 256), 256, 0, st>>>(Y, stats, gamma, beta, skip, total, C, relu, out);
    return gnn_launch_status();
}

__global__ void bn_bwd_finalize_kernel(const double* __restrict__ partial, int nparts, int C, float* sums,
                                       float* dgamma, float* dbeta) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    double s, sy;
    reduce_parts(partial, nparts, C, c, s, sy);
    if (threadIdx.y != 0 || c >= C) return;
    sums[c] = (float)s;
    sums[C + c] = (float)sy;
    if (dbeta) dbeta[c] += (float)s;
    if (dgamma) dgamma[c] += (float)sy;
}

__global__ void bn_act_bwd_apply_kernel(const float* __restrict__ dOut, const float* __restrict__ Y,
                                        const float* __restrict__ stats, const float* __restrict__ gamma,
                                        const float* __restrict__ beta, const float* __restrict__ sums,
                                        int64_t rows, int C, int relu, int training, float* __restrict__ dY,
                                        const int* __restrict__ valid_rows) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t total = rows * C;
    const int64_t vrows = bn_valid(rows, valid_rows), vtotal = vrows * C;
    const float inv_n = 1.0f / (float)(vrows > 0 ? vrows : 1);
    for (; i < total; i += stride) {
        if (i >= vtotal) {
            dY[i] = 0.f;
            continue;
        }
        int c = (int)(i % C);
        float mu = __ldg(stats + c), rs = __ldg(stats + C + c), ga = __ldg(gamma + c);
        float yh = (__ldg(Y + i) - mu) * rs;
        float dz = __ldg(dOut + i);
        if (relu && !(fmaf(ga, yh, __ldg(beta + c)) > 0.f)) dz = 0.f;
        float v = training ? (dz - __ldg(sums + c) * inv_n - yh * __ldg(sums + C + c) * inv_n) : dz;
        dY[i] = ga * rs * v;
    }
}

__global__ void bn_act_bwd_apply4_kernel(const float* __restrict__ dOut, const float* __restrict__ Y,
                                         const float* __restrict__ stats, const float* __restrict__ gamma,
                                         const float* __restrict__ beta, const float* __restrict__ sums,
                                         int64_t rows, int C, int relu, int training, float* __restrict__ dY,
                                         const int* __restrict__ valid_rows) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t total4 = rows * C / 4;
    const int C4 = C >> 2;
    const int64_t vrows = bn_valid(rows, valid_rows), vtotal4 = vrows * C4;
    const float inv_n = 1.0f / (float)(vrows > 0 ? vrows : 1);
    for (; i < total4; i += stride) {
        if (i >= vtotal4) {
            st4(dY + 4 * i, make_float4(0.f, 0.f, 0.f, 0.f));
            continue;
        }
        const int c = (int)(i % C4) * 4;
        const float4 mu = ldg4(stats + c), rs = ldg4(stats + C + c), ga = ldg4(gamma + c), be = ldg4(beta + c);
        const float4 s0 = ldg4(sums + c), s1 = ldg4(sums + C + c);
        const float4 y = ldg4(Y + 4 * i);
        float4 dz = ldg4(dOut + 4 * i);
        float4 o;
#define BWD(X)                                                                      \
    {                                                                               \
        const float yh = (y.X - mu.X) * rs.X;                                       \
        if (relu && !(fmaf(ga.X, yh, be.X) > 0.f)) dz.X = 0.f;                      \
        const float v = training ? (dz.X - s0.X * inv_n - yh * s1.X * inv_n) : dz.X; \
        o.X = ga.X * rs.X * v;                                                      \
    }
        BWD(x) BWD(y) BWD(z) BWD(w)
#undef BWD
        st4(dY + 4 * i, o);
    }
}

extern "C" int gnn_bn_act_bwd(const float* dOut, const float* Y, const float* stats, const float* gamma,
                              const float* beta, int64_t rows, int C, int relu, int training, float* dY,
                              float* dgamma, float* dbeta, const int32_t* valid_rows, void* workspace,
                              int64_t workspace_bytes, void* stream) {
    if (rows < 0 || C < 1) return GNN_ERR_BAD_ARG;
    if (rows == 0) return 0;
    if (workspace_bytes < gnn_bn_workspace_bytes(rows, C)) return GNN_ERR_WORKSPACE;
    cudaStream_t st = gnn_stream(stream);
    int parts = red_ctas(rows);
    double* partial = reinterpret_cast<double*>(workspace);
    float* sums = reinterpret_cast<float*>(partial + (int64_t)parts * 2 * C);
    if (bn_vec_ok(C))
        bn_partial4_kernel<1><<<parts, BN4_THREADS, BN4_THREADS * 8 * sizeof(double), st>>>(Y, dOut, stats, gamma, beta, rows, C, relu,
                                                                            partial, valid_rows);
    else
        bn_partial_kernel<1><<<parts, dim3(RED_TX, RED_TY), 0, st>>>(Y, dOut, stats, gamma, beta, rows, C, relu, partial,
                                                                     valid_rows);
    bn_bwd_finalize_kernel<<<(C + 31) / 32, dim3(32, FIN_TY), 0, st>>>(partial, parts, C, sums, dgamma, dbeta);
    if ((C & 3) == 0)
        bn_act_bwd_apply4_kernel<<<ew_grid(rows * C / 4, 256), 256, 0, st>>>(dOut, Y, stats, gamma, beta, sums, rows, C,
                                                                             relu, training, dY, valid_rows);
    else
        bn_act_bwd_apply_kernel<<<ew_grid(rows * C, 256), 256, 0, st>>>(dOut, Y, stats, gamma, beta, sums, rows, C, relu,
                                                                        training, dY, valid_rows);
    return gnn_launch_status();
}

// Eval-mode BatchNorm folded into the preceding Linear:  BN(x W^T + b) = x Wf^T + bf  with
//   Wf[n, :] = W[n, :] * gamma[n] * rstd[n],   bf[n] = (b[n] - mean[n]) * gamma[n] * rstd[n] + beta[n]
__global__ void bn_fold_linear_kernel(const float* __restrict__ W, int64_t ldw, const float* __restrict__ b,
                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                      const float* __restrict__ mean, const float* __restrict__ var, float eps, int N,
                                      int K, float* __restrict__ Wf, float* __restrict__ bf) {
    const int n = blockIdx.x;
    const float sc = gamma[n] * (float)(1.0 / sqrt((double)var[n] + (double)eps));
    for (int k = threadIdx.x; k < K; k += blockDim.x) Wf[(int64_t)n * K + k] = W[(int64_t)n * ldw + k] * sc;
    if (threadIdx.x == 0) bf[n] = fmaf((b ? b[n] : 0.f) - mean[n], sc, beta[n]);
}
extern "C" int gnn_bn_fold_linear(const float* W, int64_t ldw, const float* bias, const float* gamma, const float* beta,
                                  const float* running_mean, const float* running_var, float eps, int N, int K,
                                  float* Wf, float* bf, void* stream) {
    if (N < 1 || K < 1 || !W || !gamma || !beta || !running_mean || !running_var) return GNN_ERR_BAD_ARG;
    bn_fold_linear_kernel<<<N, 64, 0, gnn_stream(stream)>>>(W, ldw, bias, gamma, beta, running_mean, running_var, eps, N,
                                                           K, Wf, bf);
    return gnn_launch_status();
}

// ------------------------------------------------------------------ segment pooling
__global__ void segment_pool_fwd_kernel(const float* __restrict__ x, const int* __restrict__ graph_ptr, int64_t b,
                                        int C, float* __restrict__ out) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= b * C) return;
    int64_t g = idx / C;
    int c = (int)(idx - g * C);
    int r0 = graph_ptr[g], r1 = graph_ptr[g + 1];
    float acc = 0.f;
    for (int r = r0; r < r1; ++r) acc += __ldg(x + (int64_t)r * C + c);
    out[idx] = acc;
}
extern "C" int gnn_segment_pool_fwd(const float* x, const int32_t* graph_ptr, int64_t b, int C, float* out,
                                    void* stream) {
    if (b < 0 || C < 1) return GNN_ERR_BAD_ARG;
    if (b == 0) return 0;
    segment_pool_fwd_kernel<<<(unsigned)gnn_ceil_div(b * C, 256), 256, 0, gnn_stream(stream)>>>(x, graph_ptr, b, C, out);
    return gnn_launch_status();
}

__global__ void segment_pool_bwd_kernel(const float* __restrict__ dout, const int* __restrict__ node_graph,
                                        int64_t total, int C, int accumulate, float* __restrict__ dx) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        int64_t r = i / C;
        int c = (int)(i - r * C);
        int g = node_graph ? node_graph[r] : 0;
        float v = __ldg(dout + (int64_t)g * C + c);
        dx[i] = accumulate ? dx[i] + v : v;
    }
}
// float4 variant (C % 4 == 0, 16-byte aligned rows): thread = (node, column quad), 32-bit index arithmetic
__global__ void segment_pool_bwd4_kernel(const float* __restrict__ dout, const int* __restrict__ node_graph, int n,
                                         int C4, int accumulate, float* __restrict__ dx) {
    const int total = n * C4;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int r = i / C4, c = i - r * C4;
        const int g = node_graph ? __ldg(node_graph + r) : 0;
        float4 v = ldg4(dout + ((size_t)g * C4 + c) * 4);
        float4* dst = reinterpret_cast<float4*>(dx) + i;
        if (accumulate) {
            const float4 o = *dst;
            v = make_float4(v.x + o.x, v.y + o.y, v.z + o.z, v.w + o.w);
        }
        *dst = v;
    }
}
extern "C" int gnn_segment_pool_bwd(const float* dout, const int32_t* node_graph, int64_t n, int C, int accumulate,
                                    float* dx, void* stream) {
    if (n < 0 || C < 1) return GNN_ERR_BAD_ARG;
    if (n == 0) return 0;
    if ((C & 3) == 0 && n * (int64_t)(C >> 2) < (1ll << 31) && ((reinterpret_cast<uintptr_t>(dout) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0) {
        segment_pool_bwd4_kernel<<<ew_grid(n * (C >> 2), 256), 256, 0, gnn_stream(stream)>>>(dout, node_graph, (int)n, C >> 2,
                                                                                           accumulate, dx);
        return gnn_launch_status();
    }
    segment_pool_bwd_kernel<<<ew_grid(n * C, 256), 256, 0, gnn_stream(stream)>>>(dout, node_graph, n * C, C, accumulate,
                                                                                  dx);
    return gnn_launch_status();
}

// ------------------------------------------------------------------ MAPE (single CTA, deterministic)
__global__ void __launch_bounds__(1024)
mape_kernel(const float* __restrict__ pred, const float* __restrict__ target, int64_t total, float grad_scale,
            float* loss, float* dpred, const int* __restrict__ valid_rows, int row_len) {
    __shared__ double sw[32];
    double acc = 0.0;
    // padded batches: only the first *valid_rows rows (graphs) are real; the padding graph gets a zero gradient
    int64_t numel = total;
    if (valid_rows != nullptr) {
        const int64_t v = (int64_t)__ldg(valid_rows) * row_len;
        numel = v < total ? (v < 0 ? 0 : v) : total;
    }
    const float inv = 1.0f / (float)(numel > 0 ? numel : 1);
    if (dpred)
        for (int64_t i = numel + threadIdx.x; i < total; i += blockDim.x) dpred[i] = 0.f;
    // four elements per trip: the eight loads are independent and in flight together (one CTA walks the whole
    // [B, P] matrix, so the loop is latency-bound otherwise); the terms are summed in the order of the plain loop
    for (int64_t i0 = threadIdx.x; i0 < numel; i0 += 4 * (int64_t)blockDim.x) {
        float p[4], t[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t i = i0 + (int64_t)u * blockDim.x;
            const bool ok = i < numel;
            p[u] = ok ? pred[i] : 0.f;
            t[u] = ok ? target[i] : 1.f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t i = i0 + (int64_t)u * blockDim.x;
            if (i < numel) {
                const float den = fmaxf(fabsf(t[u]), 1.17e-06f);
                const float d = p[u] - t[u];
                acc += (double)(fabsf(d) / den);
                if (dpred) dpred[i] = (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) / den * inv * grad_scale;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(GNN_FULL_MASK, acc, o);
    if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < (blockDim.x >> 5) ? sw[threadIdx.x] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(GNN_FULL_MASK, v, o);
        if (threadIdx.x == 0 && loss) loss[0] = (float)(v / (double)(numel > 0 ? numel : 1));
    }
}
extern "C" int gnn_mape_fwd_bwd(const float* pred, const float* target, int64_t numel, float grad_scale, float* loss,
                                float* dpred, const int32_t* valid_rows, int row_len, void* stream) {
    if (numel < 1 || (valid_rows != nullptr && row_len < 1)) return GNN_ERR_BAD_ARG;
    mape_kernel<<<1, 1024, 0, gnn_stream(stream)>>>(pred, target, numel, grad_scale, loss, dpred, valid_rows, row_len);
    return gnn_launch_status();
}

// ------------------------------------------------------------------ AdamW(amsgrad) + cosine warm restarts
__global__ void __launch_bounds__(256)
adamw_amsgrad_kernel(float* __restrict__ param, const float* __restrict__ grad, float* __restrict__ m,
                     float* __restrict__ v, float* __restrict__ vmax, int64_t numel,
                     const int64_t* __restrict__ step_count, float lr0, float beta1, float beta2, float eps, float wd,
                     int64_t t0, float grad_scale) {
    // The schedule / bias-correction scalars need double-precision cos / pow: one thread per CTA computes them (every
    // thread doing so made this 2 MB kernel take 19 us).  step t = 0-based count of completed steps; torch's `step`
    // state is t + 1.
    __shared__ float sc[3];
    if (threadIdx.x == 0) {
        const int64_t t = *step_count;
        const double kPi = 3.14159265358979323846;
        double lr = lr0;
        if (t0 > 0) lr = 0.5 * (double)lr0 * (1.0 + cos(kPi * (double)(t % t0) / (double)t0));
        const double step = (double)(t + 1);
        const double bc1 = 1.0 - pow((double)beta1, step);
        const double bc2 = 1.0 - pow((double)beta2, step);
        sc[0] = (float)(lr / bc1);
        sc[1] = (float)sqrt(bc2);
        sc[2] = (float)(1.0 - lr * (double)wd);
    }
    __syncthreads();
    const float step_size = sc[0], bc2_sqrt = sc[1], decay = sc[2];
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < numel; i += stride) {
        float g = grad[i] * grad_scale;
        float p = param[i] * decay;
        float mi = m[i] + (g - m[i]) * (1.0f - beta1);          // lerp
        float vi = v[i] * beta2 + (1.0f - beta2) * g * g;
        float vm = fmaxf(vmax[i], vi);
        float denom = sqrtf(vm) / bc2_sqrt + eps;
        p -= step_size * (mi / denom);
        param[i] = p;
        m[i] = mi;
        v[i] = vi;
        vmax[i] = vm;
    }
}
extern "C" int gnn_adamw_amsgrad_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                                      float* max_exp_avg_sq, int64_t numel, const int64_t* step_count, float lr0,
                                      float beta1, float beta2, float eps, float weight_decay, int64_t t0_restart,
                                      float grad_scale, void* stream) {
    if (numel < 0) return GNN_ERR_BAD_ARG;
    if (numel == 0) return 0;
    adamw_amsgrad_kernel<<<ew_grid(numel, 256), 256, 0, gnn_stream(stream)>>>(
        param, grad, exp_avg, exp_avg_sq, max_exp_avg_sq, numel, step_count, lr0, beta1, beta2, eps, weight_decay,
        t0_restart, grad_scale);
    return gnn_launch_status();
}

__global__ void step_increment_kernel(int64_t* s) { *s += 1; }
extern "C" int gnn_step_increment(int64_t* step_count, void* stream) {
    step_increment_kernel<<<1, 1, 0, gnn_stream(stream)>>>(step_count);
    return gnn_launch_status();
}

extern "C" int gnn_fill_zero(void* ptr, int64_t bytes, void* stream) {
    if (bytes <= 0) return 0;
    cudaError_t e = cudaMemsetAsync(ptr, 0, (size_t)bytes, gnn_stream(stream));
    return e == cudaSuccess ? 0 : (int)e;
}

__global__ void relu_bwd_kernel(float* __restrict__ dY, const float* __restrict__ Yact, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride)
        if (!(__ldg(Yact + i) > 0.f)) dY[i] = 0.f;
}
extern "C" int gnn_relu_bwd(float* dY, const float* Y_act, int64_t numel, void* stream) {
    if (numel <= 0) return 0;
    relu_bwd_kernel<<<ew_grid(numel, 256), 256, 0, gnn_stream(stream)>>>(dY, Y_act, numel);
    return gnn_launch_status();
}
