// Weight-space folding of the destination projection of PNAConv into post_nns[t][0].
//
// With pre_layers == 1 the message is m_e = Wdst x_dst + q_e (q_e = Wsrc x_src + bond term + b), so
//   mean(m) = Wdst x + mean(q),  min(m) = Wdst x + min(q),  max(m) = Wdst x + max(q),  std(m) = std(q)
// for every node with at least one in-edge (always true with self-loops, reference models.py:118-121).
// post_nns[t][0] (PyG pna_conv.py; called from reference models.py:128) is linear in its 13F inputs
// [x | id(mean,min,max,std) | amp(...) | att(...)], hence the Wdst x term can be moved into the x columns:
//   O_t = sum_s scal_s(i) * ( x . Vx_{t,s}^T + A^q_t . Wagg_{t,s}^T ) + b
//   Vx_{t,s} = [s == 0] W0_t + (Wmean_{t,s} + Wmin_{t,s} + Wmax_{t,s}) . Wdst_t          [F_out, F]
// where A^q are the aggregates of q alone.  The aggregate kernels then never touch a destination projection
// (half the gathered tensor), the projection GEMM halves, and the post GEMM does the same work as before
// (the x columns of the amplification / attenuation groups were zero padding).
#include "common.cuh"

// All conv layers in one launch: layer l's parameters sit l*layer_stride floats after layer 0's (flat arena),
// Vx / dVx are [L][T][3][F_out][F].
// Vx[l][t][s][o][k];  one CTA per (t, s, o) (blockIdx.x) and layer (blockIdx.y), threads over k
__global__ void __launch_bounds__(256)
pna_fold_fwd_kernel(const float* __restrict__ wpost, int64_t wpost_ts, const float* __restrict__ wpre, int64_t wpre_ts,
                    int64_t layer_stride, int F, int F_out, float* __restrict__ Vx) {
    extern __shared__ float wsum[];          // [F]
    const int o = blockIdx.x % F_out, s = (blockIdx.x / F_out) % 3, t = blockIdx.x / (3 * F_out);
    const int64_t loff = (int64_t)blockIdx.y * layer_stride;
    const float* row = wpost + loff + (int64_t)t * wpost_ts + (int64_t)o * 13 * F;
    const float* agg = row + F + (int64_t)s * 4 * F;
    for (int f = threadIdx.x; f < F; f += blockDim.x) wsum[f] = agg[f] + agg[F + f] + agg[2 * F + f];
    __syncthreads();
    const float* wd = wpre + loff + (int64_t)t * wpre_ts;           // Wdst_t[f][k] = wpre_t[f*3F + k]
    float* out = Vx + ((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * F;
    for (int k = threadIdx.x; k < F; k += blockDim.x) {
        float acc = s == 0 ? row[k] : 0.f;
#pragma unroll 8
        for (int f = 0; f < F; ++f) acc = fmaf(wsum[f], __ldg(wd + (int64_t)f * 3 * F + k), acc);
        out[k] = acc;
    }
}

extern "C" int gnn_pna_fold_fwd(const float* wpost, int64_t wpost_tower_stride, const float* wpre,
                                int64_t wpre_tower_stride, int64_t layer_stride, int num_layers, int F, int F_out,
                                int towers, float* Vx, void* stream) {
    if (F < 1 || F_out < 1 || towers < 1 || num_layers < 0) return GNN_ERR_BAD_ARG;
    if (num_layers == 0) return 0;
    const int threads = F >= 256 ? 256 : (F + 31) / 32 * 32;
    pna_fold_fwd_kernel<<<dim3(towers * 3 * F_out, num_layers), threads, F * sizeof(float), gnn_stream(stream)>>>(
        wpost, wpost_tower_stride, wpre, wpre_tower_stride, layer_stride, F, F_out, Vx);
    return gnn_launch_status();
}

// Backward of the fold given dVx [L][T][3][F_out][F]:
//   dW0_t += dVx[t][0]
//   dW{mean,min,max}_{t,s}[o][f] += sum_k dVx[t][s][o][k] Wdst_t[f][k]
//   dWdst_t[f][k] += sum_{s,o} Wsum_{t,s}[o][f] dVx[t][s][o][k]
// blocks [0, T*3*F_out): one (t, s, o) row of the first two (a warp per f, lanes over k: coalesced rows of Wdst);
// blocks after that: one (t, f) row of dWdst (Wsum column staged in shared memory, threads over k).
__global__ void __launch_bounds__(256)
pna_fold_bwd_kernel(const float* __restrict__ dVx, const float* __restrict__ wpost, int64_t wpost_ts,
                    const float* __restrict__ wpre, int64_t wpre_ts, int64_t layer_stride, int F, int F_out, int towers,
                    float* __restrict__ dwpost, float* __restrict__ dwpre) {
    extern __shared__ float sm[];            // [max(F, 3*F_out)]
    const int rows_a = towers * 3 * F_out;
    const int64_t loff = (int64_t)blockIdx.y * layer_stride;
    const float* dv_layer = dVx + (int64_t)blockIdx.y * rows_a * F;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    if ((int)blockIdx.x < rows_a) {
        const int o = blockIdx.x % F_out, s = (blockIdx.x / F_out) % 3, t = blockIdx.x / (3 * F_out);
        const float* dv = dv_layer + (int64_t)blockIdx.x * F;
        for (int k = threadIdx.x; k < F; k += blockDim.x) sm[k] = dv[k];
        __syncthreads();
        float* drow = dwpost + loff + (int64_t)t * wpost_ts + (int64_t)o * 13 * F;
        if (s == 0)
            for (int k = threadIdx.x; k < F; k += blockDim.x) drow[k] += sm[k];
        const float* wd = wpre + loff + (int64_t)t * wpre_ts;
        float* dagg = drow + F + (int64_t)s * 4 * F;
        for (int f = warp; f < F; f += nwarps) {
            const float* wrow = wd + (int64_t)f * 3 * F;
            float acc = 0.f;
            for (int k = lane; k < F; k += 32) acc = fmaf(sm[k], __ldg(wrow + k), acc);
            acc = warp_sum(acc);
            if (lane == 0) {
                dagg[f] += acc;
                dagg[F + f] += acc;
                dagg[2 * F + f] += acc;
            }
        }
    } else {
        const int r = blockIdx.x - rows_a, f = r % F, t = r / F;
        const float* wp = wpost + loff + (int64_t)t * wpost_ts;
        for (int so = threadIdx.x; so < 3 * F_out; so += blockDim.x) {
            const int s = so / F_out, o = so - s * F_out;
            const float* agg = wp + (int64_t)o * 13 * F + F + (int64_t)s * 4 * F + f;
            sm[so] = __ldg(agg) + __ldg(agg + F) + __ldg(agg + 2 * F);
        }
        __syncthreads();
        const float* dvt = dv_layer + (int64_t)t * 3 * F_out * F;
        float* drow = dwpre + loff + (int64_t)t * wpre_ts + (int64_t)f * 3 * F;
        for (int k = threadIdx.x; k < F; k += blockDim.x) {
            float acc = 0.f;
#pragma unroll 8
            for (int so = 0; so < 3 * F_out; ++so) acc = fmaf(sm[so], __ldg(dvt + (int64_t)so * F + k), acc);
            drow[k] += acc;
        }
    }
}

extern "C" int gnn_pna_fold_bwd(const float* dVx, const float* wpost, int64_t wpost_tower_stride, const float* wpre,
                                int64_t wpre_tower_stride, int64_t layer_stride, int num_layers, int F, int F_out,
                                int towers, float* dwpost, float* dwpre, void* stream) {
    if (F < 1 || F_out < 1 || towers < 1 || num_layers < 0) return GNN_ERR_BAD_ARG;
    if (num_layers == 0) return 0;
    const int threads = F >= 128 ? 256 : 128;
    const int smf = F > 3 * F_out ? F : 3 * F_out;
    pna_fold_bwd_kernel<<<dim3(towers * 3 * F_out + towers * F, num_layers), threads, smf * sizeof(float),
                          gnn_stream(stream)>>>(dVx, wpost, wpost_tower_stride, wpre, wpre_tower_stride, layer_stride, F,
                                                F_out, towers, dwpost, dwpre);
    return gnn_launch_status();
}
