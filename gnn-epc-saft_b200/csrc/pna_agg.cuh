// Per-segment statistics of the PNA multi-aggregator (mean / min / max / std of the messages of one destination
// node), shared by the stand-alone aggregate kernels (aggregate.cu) and by the post-GEMM producers that compute the
// aggregates on the fly (gemm_tc.cu), so that both produce bit-identical values.
// Replaces PyG MultiAggregation(mean, min, max, std) (nn/aggr/{multi,basic}.py) called from PNAConv at reference
// gnnepcsaft/train/models.py:128.
#pragma once
#include "common.cuh"

// One-MUFU sqrt / reciprocal (max rel. error 2^-23); the IEEE versions expand to slow-path calls.
__device__ __forceinline__ float pna_sqrt_approx(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float pna_rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// StdAggregation: sqrt(clamp(var, 1e-5)) masked to 0 at the floor  ==  var > 1e-5 ? sqrt(var) : 0
__device__ __forceinline__ float pna_std_of_var(float var) { return var > 1e-5f ? pna_sqrt_approx(var) : 0.0f; }

// Running statistics of four channels.  Single pass on data shifted by the segment's first message q0:
//   s1 = sum(q - q0), s2 = sum((q - q0)^2)  ->  mean = q0 + s1/n, var = s2/n - (s1/n)^2
// (the shift removes the cancellation of the textbook E[q^2] - E[q]^2; unlike Welford there is no division per
// message).  TIES additionally counts how many messages equal the running min / max (for the even gradient split of
// torch's scatter_reduce amin / amax backward).
template <bool TIES>
struct PnaAcc {
    float4 q0, s1, s2, mn, mx;
    float4 nmin, nmax;       // only maintained when TIES
    __device__ __forceinline__ void first(float4 q) {
        q0 = q;
        mn = q;
        mx = q;
        s1 = make_float4(0.f, 0.f, 0.f, 0.f);
        s2 = s1;
        if (TIES) {
            nmin = make_float4(1.f, 1.f, 1.f, 1.f);
            nmax = nmin;
        }
    }
    __device__ __forceinline__ void one(float q, float q0c, float& s1c, float& s2c, float& mnc, float& mxc, float& nminc,
                                        float& nmaxc) {
        const float d = q - q0c;
        s1c += d;
        s2c = fmaf(d, d, s2c);
        if (TIES) {     // branch-free: a new extremum resets the count to 1, a tie adds 1
            nminc = (q < mnc ? 0.f : nminc) + (q <= mnc ? 1.f : 0.f);
            nmaxc = (q > mxc ? 0.f : nmaxc) + (q >= mxc ? 1.f : 0.f);
        }
        mnc = fminf(mnc, q);
        mxc = fmaxf(mxc, q);
    }
    __device__ __forceinline__ void next(float4 q) {
        one(q.x, q0.x, s1.x, s2.x, mn.x, mx.x, nmin.x, nmax.x);
        one(q.y, q0.y, s1.y, s2.y, mn.y, mx.y, nmin.y, nmax.y);
        one(q.z, q0.z, s1.z, s2.z, mn.z, mx.z, nmin.z, nmax.z);
        one(q.w, q0.w, s1.w, s2.w, mn.w, mx.w, nmin.w, nmax.w);
    }
    // mean and std of the segment (cnt >= 1 messages)
    __device__ __forceinline__ void finish(int cnt, float4& mean, float4& sd) const {
        const float inv = __frcp_rn((float)cnt);
        const float4 m1 = make_float4(s1.x * inv, s1.y * inv, s1.z * inv, s1.w * inv);
        mean = make_float4(q0.x + m1.x, q0.y + m1.y, q0.z + m1.z, q0.w + m1.w);
        sd = make_float4(pna_std_of_var(fmaf(-m1.x, m1.x, s2.x * inv)), pna_std_of_var(fmaf(-m1.y, m1.y, s2.y * inv)),
                         pna_std_of_var(fmaf(-m1.z, m1.z, s2.z * inv)), pna_std_of_var(fmaf(-m1.w, m1.w, s2.w * inv)));
    }
    // tie counts of the four channels as 4 x u16 = (nmin | nmax << 8), saturated at 255 (the backward recounts the
    // rare segments with more than 255 messages)
    __device__ __forceinline__ uint2 packed_ties() const {
        auto pk = [](float a, float b) -> uint32_t {
            return (uint32_t)fminf(a, 255.f) | ((uint32_t)fminf(b, 255.f) << 8);
        };
        return make_uint2(pk(nmin.x, nmax.x) | (pk(nmin.y, nmax.y) << 16), pk(nmin.z, nmax.z) | (pk(nmin.w, nmax.w) << 16));
    }
};

__device__ __forceinline__ float4 pna_add4(float4 a, float4 b) {
    return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
