// Shared helpers for the sm_100a kernels of the PNAPCSAFT path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gnnepcsaft_b200.h"

#define GNN_FULL_MASK 0xffffffffu

static inline int gnn_launch_status() {
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
}

static inline cudaStream_t gnn_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// number of SMs of the current device (cached per device ordinal; 148 on B200)
int gnn_sm_count();

static inline int64_t gnn_ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t gnn_align_up(int64_t a, int64_t b) { return gnn_ceil_div(a, b) * b; }

__device__ __forceinline__ float4 ldg4(const float* p) {
    return __ldg(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
// streaming store: data written once and not re-read by this kernel (keeps L1/L2 for gathers)
__device__ __forceinline__ void st4_stream(float* p, float4 v) {
    __stcs(reinterpret_cast<float4*>(p), v);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(GNN_FULL_MASK, v, o);
    return v;
}

// degree scalers of DegreeScalerAggregation (PyG nn/aggr/scaler.py): amplification uses the raw
// degree, attenuation clamps it to >= 1.
__device__ __forceinline__ void pna_scalers(int deg, float avg_deg_log, float& amp, float& att) {
    float d = (float)deg;
    amp = logf(d + 1.0f) / avg_deg_log;
    att = avg_deg_log / logf(fmaxf(d, 1.0f) + 1.0f);
}
