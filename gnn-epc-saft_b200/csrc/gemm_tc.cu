// Dense stages on the 5th-generation tensor cores: tcgen05.mma with accumulators in TMEM.
// Replace the same torch.nn.Linear / PyG Linear call sites as gemm_simt.cu (reference
// gnnepcsaft/train/models.py:69-103,128,134 -> PyG nn/conv/pna_conv.py pre_nns/post_nns/lin).
//
// Shape of the problem: M = nodes (large), N and K small (32..832), operands that do not exist in
// memory as matrices (weight slices, the virtual [x | A] concatenation, degree-scaled gradients).
// So operands are *produced* by all threads straight into the UMMA canonical K-major shared-memory
// layout (no swizzle: 8x16-byte core matrices, LBO between the 16-byte K chunks, SBO between 8-row
// groups), one elected thread issues tcgen05.mma, completion is tracked with tcgen05.commit ->
// mbarrier, and the epilogue reads the accumulator with tcgen05.ld.
//
// Precision modes
//   TF32X3: every fp32 operand is split into tf32 hi + lo parts and three MMAs are issued
//           (hi*hi + hi*lo + lo*hi), which restores ~fp32 accuracy (the exact-parity path);
//   BF16  : operands rounded to bf16, fp32 accumulate (what the reference's
//           torch.set_float32_matmul_precision("medium"), train.py:319, allows).
#include "common.cuh"

#include <cuda_bf16.h>

namespace tc {

constexpr int BM = 128;          // UMMA M (cta_group::1), one TMEM lane per row
constexpr int THREADS = 256;     // 8 warps: all produce operands, all drain the accumulator
constexpr int ROW_BYTES = 128;   // bytes of K per row per k-block (32 tf32 or 64 bf16)
constexpr int CHUNKS = ROW_BYTES / 16;
enum { MODE_TF32X3 = 0, MODE_BF16 = 1 };

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    const uint32_t addr = smem_u32(bar);
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
                 "r"(ncols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols));
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
template <int MODE>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    if (MODE == MODE_TF32X3)
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
            : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
            : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// shared-memory matrix descriptor, K-major, SWIZZLE_NONE (cute::UMMA::SmemDescriptor bit layout):
// [0,14) start>>4 | [16,30) leading byte offset>>4 | [32,46) stride byte offset>>4 | [46,48) version=1
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c=F32, a/b format, K-major both, N>>3, M>>4
__host__ __device__ inline uint32_t instr_desc(int mode, int n) {
    uint32_t fmt = mode == MODE_TF32X3 ? 2u : 1u;   // TF32 = 2, BF16 = 1
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// ------------------------------------------------------------------ operand functors
// get4(row, k): elements (row, k..k+3) of the K-major operand view, zero outside [rows) x [cols)
struct VRowMajor {          // element (r, k) = p[r*ld + k]
    const float* p;
    int64_t ld, rows, cols, group_stride;
    int vec;                // 1 if 16-byte vector loads are legal (ld % 4 == 0, base aligned)
    __device__ __forceinline__ void set_group(int g) { p += (int64_t)g * group_stride; }
    __device__ __forceinline__ float4 get4(int64_t r, int64_t k) const {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r >= rows || k >= cols) return v;
        const float* q = p + r * ld + k;
        if (vec && k + 3 < cols) return ldg4(q);
        v.x = __ldg(q);
        if (k + 1 < cols) v.y = __ldg(q + 1);
        if (k + 2 < cols) v.z = __ldg(q + 2);
        if (k + 3 < cols) v.w = __ldg(q + 3);
        return v;
    }
};
struct VTransposed {        // element (r, k) = p[k*ld + r]
    const float* p;
    int64_t ld, rows, cols, group_stride;
    __device__ __forceinline__ void set_group(int g) { p += (int64_t)g * group_stride; }
    __device__ __forceinline__ float4 get4(int64_t r, int64_t k) const {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r >= rows || k >= cols) return v;
        const float* q = p + k * ld + r;
        v.x = __ldg(q);
        if (k + 1 < cols) v.y = __ldg(q + ld);
        if (k + 2 < cols) v.z = __ldg(q + 2 * ld);
        if (k + 3 < cols) v.w = __ldg(q + 3 * ld);
        return v;
    }
};
// [x | A_t] (node i, k in [0,5F)); group = tower.  F % 4 == 0.
struct VCat5 {
    const float* x;
    const float* A;
    int64_t n;
    int F, towers, tower;
    __device__ __forceinline__ void set_group(int g) { tower = g; }
    __device__ __forceinline__ float4 get4(int64_t i, int64_t k) const {
        if (i >= n || k >= 5 * (int64_t)F) return make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < F) return ldg4(x + i * F + k);
        return ldg4(A + (i * towers + tower) * (4 * (int64_t)F) + (k - F));
    }
};
// transposed view of the same: row = column of [x | A_t] restricted to [col0, col0+rows), k = node
struct VCat5T {
    const float* x;
    const float* A;
    int64_t n;
    int F, towers, tower, col0, rows;
    __device__ __forceinline__ void set_group(int g) { tower = g; }
    __device__ __forceinline__ float elem(int64_t i, int c) const {
        if (i >= n) return 0.f;
        return c < F ? __ldg(x + i * F + c) : __ldg(A + (i * towers + tower) * (4 * (int64_t)F) + (c - F));
    }
    __device__ __forceinline__ float4 get4(int64_t r, int64_t k) const {
        if (r >= rows) return make_float4(0.f, 0.f, 0.f, 0.f);
        int c = col0 + (int)r;
        return make_float4(elem(k, c), elem(k + 1, c), elem(k + 2, c), elem(k + 3, c));
    }
};
// post forward weights.  Output features are tiled FoP at a time (FoP multiple of 16); tile j holds
// rows [j*3*FoP, (j+1)*3*FoP) = three scaler groups of FoP rows: row r -> (tile, s, oo), o = tile*FoP + oo.
//   k <  F : s == 0 ? Wpost_t[o, k] : 0          k >= F : Wpost_t[o, F + s*4F + (k-F)]
struct VPostWeight {
    const float* w;
    int64_t tower_stride;
    int F, F_out, FoP;
    __device__ __forceinline__ void set_group(int g) { w += (int64_t)g * tower_stride; }
    __device__ __forceinline__ float4 get4(int64_t r, int64_t k) const {
        float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        int tile = (int)r / (3 * FoP), rem = (int)r - tile * 3 * FoP;
        int s = rem / FoP, o = tile * FoP + (rem - s * FoP);
        if (o >= F_out || k >= 5 * (int64_t)F) return z;
        const float* row = w + (int64_t)o * 13 * F;
        if (k < F) {
            if (s != 0) return z;
            return make_float4(__ldg(row + k), __ldg(row + k + 1), __ldg(row + k + 2), __ldg(row + k + 3));
        }
        const float* q = row + F + (int64_t)s * 4 * F + (k - F);
        return make_float4(__ldg(q), __ldg(q + 1), __ldg(q + 2), __ldg(q + 3));
    }
};
// [dO_t | amp dO_t | att dO_t] as K-major A operand: (node i, k = s*F_out + o)
struct VScaledGrad {
    const float* dO;
    const float* scal;
    int64_t n;
    int F_out, towers, tower;
    __device__ __forceinline__ void set_group(int g) { tower = g; }
    __device__ __forceinline__ float elem(int64_t i, int64_t k) const {
        if (k >= 3 * (int64_t)F_out) return 0.f;
        int s = (int)k / F_out, o = (int)k - s * F_out;
        float v = __ldg(dO + i * ((int64_t)towers * F_out) + tower * F_out + o);
        return s == 0 ? v : v * __ldg(scal + 2 * i + (s - 1));
    }
    __device__ __forceinline__ float4 get4(int64_t i, int64_t k) const {
        if (i >= n) return make_float4(0.f, 0.f, 0.f, 0.f);
        return make_float4(elem(i, k), elem(i, k + 1), elem(i, k + 2), elem(i, k + 3));
    }
};
// the same, transposed: row r = s*FoP + o, k = node (wgrad A operand)
struct VScaledGradT {
    const float* dO;
    const float* scal;
    int64_t n;
    int F_out, FoP, towers, tower;
    __device__ __forceinline__ void set_group(int g) { tower = g; }
    __device__ __forceinline__ float4 get4(int64_t r, int64_t k) const {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        int s = (int)r / FoP, o = (int)r - s * FoP;
        if (s >= 3 || o >= F_out) return v;
        float* pv = &v.x;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int64_t i = k + j;
            if (i < n) {
                float d = __ldg(dO + i * ((int64_t)towers * F_out) + tower * F_out + o);
                pv[j] = s == 0 ? d : d * __ldg(scal + 2 * i + (s - 1));
            }
        }
        return v;
    }
};
// post dgrad B operand: row = r in [0,4F) (column of A_t), k = s*F_out + o -> Wpost_t[o, F + s*4F + r]
struct VPostWeightT {
    const float* w;
    int64_t tower_stride;
    int F, F_out;
    __device__ __forceinline__ void set_group(int g) { w += (int64_t)g * tower_stride; }
    __device__ __forceinline__ float elem(int64_t r, int64_t k) const {
        if (k >= 3 * (int64_t)F_out) return 0.f;
        int s = (int)k / F_out, o = (int)k - s * F_out;
        return __ldg(w + (int64_t)o * 13 * F + F + (int64_t)s * 4 * F + r);
    }
    __device__ __forceinline__ float4 get4(int64_t r, int64_t k) const {
        if (r >= 4 * (int64_t)F) return make_float4(0.f, 0.f, 0.f, 0.f);
        return make_float4(elem(r, k), elem(r, k + 1), elem(r, k + 2), elem(r, k + 3));
    }
};

// ------------------------------------------------------------------ epilogues: 16 consecutive columns of one row
struct EStore {     // Y[m, n] = act(v + bias[n]) (+ Y)
    float* y;
    int64_t ldy, y_group_stride;
    const float* bias;
    int64_t bias_group_stride;
    int act, accumulate;
    static constexpr int kCombine3 = 0;
    __device__ __forceinline__ void set_group(int g) {
        y += (int64_t)g * y_group_stride;
        if (bias) bias += (int64_t)g * bias_group_stride;
    }
    __device__ __forceinline__ void store16(int64_t m, int n0, int N, const float* v) const {
        float* dst = y + m * ldy + n0;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (n0 + j < N) {
                float t = v[j];
                if (bias) t += __ldg(bias + n0 + j);
                if (act) t = fmaxf(t, 0.f);
                if (accumulate) t += dst[j];
                dst[j] = t;
            }
        }
    }
};
struct EAtomic {    // dW[m, n] += v   (m, n both bounded)
    float* y;
    int64_t ldy, y_group_stride;
    static constexpr int kCombine3 = 0;
    __device__ __forceinline__ void set_group(int g) { y += (int64_t)g * y_group_stride; }
    __device__ __forceinline__ void store16(int64_t m, int n0, int N, const float* v) const {
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if (n0 + j < N) atomicAdd(y + m * ldy + n0 + j, v[j]);
    }
};
// post forward: O[i, t*F_out + o] = act(D[i, o] + amp_i D[i, FoP+o] + att_i D[i, 2FoP+o] + b_t[o])
struct EPostCombine {
    float* O;
    const float* scal;
    const float* bias;
    int64_t bias_tower_stride;
    int F_out, FoP, towers, tower, act;
    static constexpr int kCombine3 = 1;
    __device__ __forceinline__ void set_group(int g) { tower = g; bias += (int64_t)g * bias_tower_stride; }
    __device__ __forceinline__ void store16x3(int64_t i, int o0, const float* d0, const float* d1, const float* d2) const {
        const float amp = __ldg(scal + 2 * i), att = __ldg(scal + 2 * i + 1);
        float* dst = O + i * ((int64_t)towers * F_out) + tower * F_out + o0;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (o0 + j < F_out) {
                float t = d0[j] + amp * d1[j] + att * d2[j] + __ldg(bias + o0 + j);
                dst[j] = act ? fmaxf(t, 0.f) : t;
            }
        }
    }
};
// post wgrad (A part): D[(s,o), r] -> dWpost_t[o, F + s*4F + r]  (rows are s*FoP + o)
struct EPostWgrad {
    float* dw;
    int64_t tower_stride;
    int F, F_out, FoP;
    static constexpr int kCombine3 = 0;
    __device__ __forceinline__ void set_group(int g) { dw += (int64_t)g * tower_stride; }
    __device__ __forceinline__ void store16(int64_t m, int n0, int N, const float* v) const {
        int s = (int)m / FoP, o = (int)m - s * FoP;
        if (s >= 3 || o >= F_out) return;
        float* dst = dw + (int64_t)o * 13 * F + F + (int64_t)s * 4 * F + n0;
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if (n0 + j < N) atomicAdd(dst + j, v[j]);
    }
};

// ------------------------------------------------------------------ the kernel
// grid: x = M tiles, y = N tiles (of BN columns), z = groups * k-splits
template <int MODE, class AL, class BL, class EP>
__global__ void __launch_bounds__(THREADS)
gemm_tc_kernel(AL al, BL bl, EP ep, int64_t M, int N, int BN, int64_t K, int64_t k_chunk, int groups, uint32_t tmem_cols) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t mbar_stage[2];
    __shared__ __align__(8) uint64_t mbar_acc;
    __shared__ uint32_t tmem_base_s;

    constexpr int KB_ELEMS = MODE == MODE_TF32X3 ? 32 : 64;       // K elements per k-block
    constexpr int PARTS = MODE == MODE_TF32X3 ? 2 : 1;            // hi + lo
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = blockIdx.z % groups, split = blockIdx.z / groups;
    al.set_group(g);
    bl.set_group(g);
    ep.set_group(g);
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int64_t k_begin = (int64_t)split * k_chunk;
    const int64_t k_end = min(K, k_begin + k_chunk);
    const int num_kb = k_begin < k_end ? (int)((k_end - k_begin + KB_ELEMS - 1) / KB_ELEMS) : 0;

    // shared-memory carve-up (128-byte aligned): per stage [A parts][B parts]
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    const uint32_t a_bytes = BM * ROW_BYTES, b_bytes = (uint32_t)BN * ROW_BYTES;
    const uint32_t stage_bytes = PARTS * (a_bytes + b_bytes);

    if (warp == 0) tmem_alloc(&tmem_base_s, tmem_cols);
    if (tid == 32) {
        mbar_init(&mbar_stage[0], 1);
        mbar_init(&mbar_stage[1], 1);
        mbar_init(&mbar_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = tmem_base_s;
    const uint32_t idesc = instr_desc(MODE, BN);

    for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb & 1;
        uint8_t* stage = smem + (size_t)s * stage_bytes;
        if (kb >= 2) mbar_wait(&mbar_stage[s], (uint32_t)(((kb >> 1) - 1) & 1));   // MMAs that read this stage retired
        const int64_t kbase = k_begin + (int64_t)kb * KB_ELEMS;

        // ---- produce the A tile: thread -> (row, 4 consecutive 16-byte chunks)
        {
            const int row = tid & (BM - 1);
            const int c0 = (tid >> 7) * (CHUNKS / 2);
#pragma unroll
            for (int c = 0; c < CHUNKS / 2; ++c) {
                const int chunk = c0 + c;
                uint8_t* dst = stage + (size_t)chunk * (BM * 16) + row * 16;
                if (MODE == MODE_TF32X3) {
                    int64_t k = kbase + chunk * 4;
                    float4 v = k < k_end ? al.get4(m0 + row, k) : make_float4(0.f, 0.f, 0.f, 0.f);
                    if (k + 3 >= k_end) {   // split-K boundary inside the chunk
                        if (k + 1 >= k_end) v.y = 0.f;
                        if (k + 2 >= k_end) v.z = 0.f;
                        if (k + 3 >= k_end) v.w = 0.f;
                    }
                    float4 hi = make_float4(__uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u),
                                            __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u),
                                            __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u),
                                            __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u));
                    float4 lo = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
                    *reinterpret_cast<float4*>(dst) = hi;
                    *reinterpret_cast<float4*>(dst + a_bytes) = lo;
                } else {
                    int64_t k = kbase + chunk * 8;
                    float4 v0 = k < k_end ? al.get4(m0 + row, k) : make_float4(0.f, 0.f, 0.f, 0.f);
                    float4 v1 = k + 4 < k_end ? al.get4(m0 + row, k + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                    float e[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
                    uint32_t pk[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float lo_ = k + 2 * j < k_end ? e[2 * j] : 0.f, hi_ = k + 2 * j + 1 < k_end ? e[2 * j + 1] : 0.f;
                        __nv_bfloat162 b2 = __floats2bfloat162_rn(lo_, hi_);
                        pk[j] = *reinterpret_cast<uint32_t*>(&b2);
                    }
                    *reinterpret_cast<uint4*>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                }
            }
        }
        // ---- produce the B tile: BN rows x 8 chunks
        {
            uint8_t* bst = stage + PARTS * a_bytes;
            for (int idx = tid; idx < BN * CHUNKS; idx += THREADS) {
                const int row = idx % BN, chunk = idx / BN;
                uint8_t* dst = bst + (size_t)chunk * (BN * 16) + row * 16;
                if (MODE == MODE_TF32X3) {
                    int64_t k = kbase + chunk * 4;
                    float4 v = k < k_end ? bl.get4(n0 + row, k) : make_float4(0.f, 0.f, 0.f, 0.f);
                    if (k + 3 >= k_end) {
                        if (k + 1 >= k_end) v.y = 0.f;
                        if (k + 2 >= k_end) v.z = 0.f;
                        if (k + 3 >= k_end) v.w = 0.f;
                    }
                    float4 hi = make_float4(__uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u),
                                            __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u),
                                            __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u),
                                            __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u));
                    float4 lo = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
                    *reinterpret_cast<float4*>(dst) = hi;
                    *reinterpret_cast<float4*>(dst + b_bytes) = lo;
                } else {
                    int64_t k = kbase + chunk * 8;
                    float4 v0 = k < k_end ? bl.get4(n0 + row, k) : make_float4(0.f, 0.f, 0.f, 0.f);
                    float4 v1 = k + 4 < k_end ? bl.get4(n0 + row, k + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                    float e[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
                    uint32_t pk[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float lo_ = k + 2 * j < k_end ? e[2 * j] : 0.f, hi_ = k + 2 * j + 1 < k_end ? e[2 * j + 1] : 0.f;
                        __nv_bfloat162 b2 = __floats2bfloat162_rn(lo_, hi_);
                        pk[j] = *reinterpret_cast<uint32_t*>(&b2);
                    }
                    *reinterpret_cast<uint4*>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                }
            }
        }
        fence_async_smem();          // generic-proxy smem writes -> visible to the tensor-core (async) proxy
        __syncthreads();

        // ---- one thread issues the MMAs of this k-block
        if (tid == 0) {
            tc_fence_after();
            const uint32_t a_hi = smem_u32(stage), a_lo = a_hi + a_bytes;
            const uint32_t b_hi = smem_u32(stage + PARTS * a_bytes), b_lo = b_hi + b_bytes;
            const uint32_t a_lbo = BM * 16, b_lbo = (uint32_t)BN * 16;
#pragma unroll
            for (int j = 0; j < CHUNKS / 2; ++j) {     // each MMA consumes 32 bytes of K = 2 chunks
                const uint32_t aoff = (uint32_t)(2 * j) * a_lbo, boff = (uint32_t)(2 * j) * b_lbo;
                const uint32_t first = (kb == 0 && j == 0) ? 0u : 1u;
                if (MODE == MODE_TF32X3) {
                    umma<MODE>(tmem_d, smem_desc(a_lo + aoff, a_lbo, 128), smem_desc(b_hi + boff, b_lbo, 128), idesc, first);
                    umma<MODE>(tmem_d, smem_desc(a_hi + aoff, a_lbo, 128), smem_desc(b_lo + boff, b_lbo, 128), idesc, 1u);
                    umma<MODE>(tmem_d, smem_desc(a_hi + aoff, a_lbo, 128), smem_desc(b_hi + boff, b_lbo, 128), idesc, 1u);
                } else {
                    umma<MODE>(tmem_d, smem_desc(a_hi + aoff, a_lbo, 128), smem_desc(b_hi + boff, b_lbo, 128), idesc, first);
                }
            }
            umma_commit(&mbar_stage[s]);                 // frees this smem stage when the MMAs retire
            if (kb == num_kb - 1) umma_commit(&mbar_acc);   // accumulator complete
        }
    }

    // ---- epilogue: warp w owns TMEM lanes [32*(w%4), +32); warps w and w+4 split the column chunks
    if (num_kb > 0) {
        mbar_wait(&mbar_acc, 0);
        tc_fence_after();
        const int q = warp & 3, half = warp >> 2;
        const int64_t m = m0 + q * 32 + lane;
        const uint32_t lane_addr = tmem_d + ((uint32_t)(q * 32) << 16);
        if constexpr (EP::kCombine3) {
            // BN = 3*FoP: columns [0,FoP) identity, [FoP,2FoP) amplification, [2FoP,3FoP) attenuation
            const int FoP = BN / 3;
            for (int c = half; c < FoP / 16; c += 2) {
                float d0[16], d1[16], d2[16];
                tmem_ld16(lane_addr + c * 16, d0);
                tmem_ld16(lane_addr + FoP + c * 16, d1);
                tmem_ld16(lane_addr + 2 * FoP + c * 16, d2);
                if (m < M) ep.store16x3(m, (int)blockIdx.y * FoP + c * 16, d0, d1, d2);
            }
        } else {
            for (int c = half; c < BN / 16; c += 2) {
                float v[16];
                tmem_ld16(lane_addr + c * 16, v);
                if (m < M && n0 + c * 16 < N) ep.store16(m, n0 + c * 16, N, v);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d, tmem_cols);
}


static uint32_t tmem_cols_for(int bn) {
    uint32_t c = 32;
    while ((int)c < bn) c <<= 1;
    return c;
}

template <int MODE, class AL, class BL, class EP>
static int launch(AL al, BL bl, EP ep, int64_t M, int N, int BN, int64_t K, int64_t k_chunk, int splits, int groups,
                  cudaStream_t st) {
    if (M <= 0 || N <= 0 || K <= 0) return 0;
    if (BN % 16 != 0 || BN < 16 || BN > 256) return GNN_ERR_UNSUPPORTED;
    const int parts = MODE == MODE_TF32X3 ? 2 : 1;
    const size_t smem = 2 * (size_t)parts * (BM + BN) * ROW_BYTES + 256;
    auto kern = gemm_tc_kernel<MODE, AL, BL, EP>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    dim3 grid((unsigned)gnn_ceil_div(M, BM), (unsigned)gnn_ceil_div(N, BN), (unsigned)(groups * splits));
    kern<<<grid, THREADS, smem, st>>>(al, bl, ep, M, N, BN, K, k_chunk, groups, tmem_cols_for(BN));
    return gnn_launch_status();
}

static int pick_bn(int n) {
    int bn = (n + 15) / 16 * 16;
    return bn > 256 ? 256 : bn;
}
static int is_vec(const float* p, int64_t ld) { return (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(p) & 15) == 0); }

// split-K geometry for weight gradients (K = rows): ~2 CTAs per SM in total
static void splitk(int64_t K, int64_t tiles, int kb_elems, int64_t& k_chunk, int& splits) {
    int64_t target = (int64_t)gnn_sm_count() * 2;
    int64_t s = gnn_ceil_div(target, tiles);
    int64_t max_s = gnn_ceil_div(K, 4 * (int64_t)kb_elems);
    if (s > max_s) s = max_s;
    if (s < 1) s = 1;
    if (s > 4096) s = 4096;
    k_chunk = gnn_align_up(gnn_ceil_div(K, s), kb_elems);
    splits = (int)gnn_ceil_div(K, k_chunk);
}

}  // namespace tc

using namespace tc;

#define DISPATCH_MODE(precision, CALL)                         \
    do {                                                       \
        if ((precision) == 0) { constexpr int MODE = MODE_TF32X3; CALL; } \
        else if ((precision) == 2) { constexpr int MODE = MODE_BF16; CALL; } \
        else return GNN_ERR_BAD_ARG;                           \
    } while (0)

extern "C" int gnn_tc_linear_fwd(const float* X, int64_t ldx, const float* W, int64_t ldw, const float* bias, int64_t M,
                                 int N, int K, int act, float* Y, int64_t ldy, int precision, void* stream) {
    if (M < 0 || N < 1 || K < 1) return GNN_ERR_BAD_ARG;
    VRowMajor al{X, ldx, M, K, 0, is_vec(X, ldx)};
    VRowMajor bl{W, ldw, N, K, 0, is_vec(W, ldw)};
    EStore ep{Y, ldy, 0, bias, 0, act, 0};
    const int BN = pick_bn(N);
    DISPATCH_MODE(precision, return (launch<MODE>(al, bl, ep, M, N, BN, K, K, 1, 1, gnn_stream(stream))));
}

extern "C" int gnn_tc_linear_dgrad(const float* dY, int64_t lddy, const float* W, int64_t ldw, int64_t M, int N, int K,
                                   int accumulate, float* dX, int64_t lddx, int precision, void* stream) {
    if (M < 0 || N < 1 || K < 1) return GNN_ERR_BAD_ARG;
    VRowMajor al{dY, lddy, M, N, 0, is_vec(dY, lddy)};     // (m, n) reduction over n
    VTransposed bl{W, ldw, K, N, 0};                       // row kk, k = n -> W[n, kk]
    EStore ep{dX, lddx, 0, nullptr, 0, 0, accumulate};
    const int BN = pick_bn(K);
    DISPATCH_MODE(precision, return (launch<MODE>(al, bl, ep, M, K, BN, N, N, 1, 1, gnn_stream(stream))));
}

extern "C" int gnn_tc_linear_wgrad(const float* dY, int64_t lddy, const float* X, int64_t ldx, int64_t M, int N, int K,
                                   float* dW, int64_t lddw, int precision, void* stream) {
    if (M < 0 || N < 1 || K < 1) return GNN_ERR_BAD_ARG;
    VTransposed al{dY, lddy, N, M, 0};                     // row o, k = i -> dY[i, o]
    VTransposed bl{X, ldx, K, M, 0};                       // row kk, k = i -> X[i, kk]
    EAtomic ep{dW, lddw, 0};
    const int BN = pick_bn(K);
    const int64_t tiles = gnn_ceil_div(N, BM) * gnn_ceil_div(K, BN);
    int64_t k_chunk;
    int splits;
    DISPATCH_MODE(precision, {
        splitk(M, tiles, MODE == MODE_TF32X3 ? 32 : 64, k_chunk, splits);
        return (launch<MODE>(al, bl, ep, N, K, BN, M, k_chunk, splits, 1, gnn_stream(stream)));
    });
}

// column sums (bias gradients): db[n] += sum_i dY[i, n]
__global__ void colsum_kernel(const float* __restrict__ dY, int64_t lddy, int64_t M, int N, float* db) {
    __shared__ float sm[8][33];
    const int n = blockIdx.x * 32 + threadIdx.x;
    const int64_t chunk = (M + gridDim.y - 1) / gridDim.y;
    const int64_t r0 = (int64_t)blockIdx.y * chunk, r1 = min(M, r0 + chunk);
    float acc = 0.f;
    if (n < N)
        for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) acc += __ldg(dY + r * lddy + n);
    sm[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && n < N) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += sm[k][threadIdx.x];
        atomicAdd(db + n, t);
    }
}
extern "C" int gnn_colsum_accum(const float* dY, int64_t lddy, int64_t M, int N, float* db, void* stream) {
    if (M < 0 || N < 1) return GNN_ERR_BAD_ARG;
    if (M == 0) return 0;
    int64_t ysplit = gnn_ceil_div(M, 2048);
    if (ysplit > 1024) ysplit = 1024;
    dim3 grid((unsigned)gnn_ceil_div(N, 32), (unsigned)ysplit);
    colsum_kernel<<<grid, dim3(32, 8), 0, gnn_stream(stream)>>>(dY, lddy, M, N, db);
    return gnn_launch_status();
}

// ------------------------------------------------------------------ PNA post stage on tensor cores
extern "C" int gnn_tc_pna_post_fwd(const float* x, const float* A, const float* scal, const float* wpost,
                                   int64_t wpost_tower_stride, const float* bpost, int64_t bpost_tower_stride,
                                   int64_t n, int F, int F_out, int towers, int act, float* O, int precision,
                                   void* stream) {
    if (n < 0 || F < 4 || (F & 3) || F_out < 1 || towers < 1) return GNN_ERR_BAD_ARG;
    int FoP = (F_out + 15) / 16 * 16;      // output-feature tile: three scaler groups of FoP columns per CTA
    if (FoP > 64) FoP = 64;
    const int tiles = (F_out + FoP - 1) / FoP;
    VCat5 al{x, A, n, F, towers, 0};
    VPostWeight bl{wpost, wpost_tower_stride, F, F_out, FoP};
    EPostCombine ep{O, scal, bpost, bpost_tower_stride, F_out, FoP, towers, 0, act};
    DISPATCH_MODE(precision, return (launch<MODE>(al, bl, ep, n, tiles * 3 * FoP, 3 * FoP, 5 * (int64_t)F,
                                                  5 * (int64_t)F, 1, towers, gnn_stream(stream))));
}

extern "C" int gnn_tc_pna_post_dgrad(const float* dO, const float* scal, const float* wpost, int64_t wpost_tower_stride,
                                     int64_t n, int F, int F_out, int towers, float* dx, float* gA, int precision,
                                     void* stream) {
    if (n < 0 || F < 4 || (F & 3) || F_out < 1 || towers < 1) return GNN_ERR_BAD_ARG;
    cudaStream_t st = gnn_stream(stream);
    // gA[:, t, :] = [dO_t | amp dO_t | att dO_t] . Wslices_t          (M = n, N = 4F, K = 3 F_out)
    {
        VScaledGrad al{dO, scal, n, F_out, towers, 0};
        VPostWeightT bl{wpost, wpost_tower_stride, F, F_out};
        EStore ep{gA, (int64_t)towers * 4 * F, 4 * (int64_t)F, nullptr, 0, 0, 0};
        const int BN = pick_bn(4 * F);
        int rc;
        DISPATCH_MODE(precision, rc = (launch<MODE>(al, bl, ep, n, 4 * F, BN, 3 * (int64_t)F_out, 3 * (int64_t)F_out, 1,
                                                    towers, st)));
        if (rc) return rc;
    }
    // dx += sum_t dO_t . Wpost_t[:, :F]     (towers accumulate into the same dx -> sequential launches)
    for (int t = 0; t < towers; ++t) {
        const float* d = dO + t * F_out;
        const float* w = wpost + t * wpost_tower_stride;
        VRowMajor al{d, (int64_t)towers * F_out, n, F_out, 0, is_vec(d, (int64_t)towers * F_out)};
        VTransposed bl{w, 13 * (int64_t)F, F, F_out, 0};
        EStore ep{dx, F, 0, nullptr, 0, 0, 1};
        const int BN = pick_bn(F);
        int rc;
        DISPATCH_MODE(precision, rc = (launch<MODE>(al, bl, ep, n, F, BN, (int64_t)F_out, (int64_t)F_out, 1, 1, st)));
        if (rc) return rc;
    }
    return 0;
}

extern "C" int gnn_tc_pna_post_wgrad(const float* dO, const float* x, const float* A, const float* scal, int64_t n,
                                     int F, int F_out, int towers, float* dwpost, int64_t wpost_tower_stride,
                                     float* dbpost, int64_t bpost_tower_stride, int precision, void* stream) {
    if (n < 0 || F < 4 || (F & 3) || F_out < 1 || towers < 1) return GNN_ERR_BAD_ARG;
    if (n == 0) return 0;
    cudaStream_t st = gnn_stream(stream);
    const int FoP = (F_out + 15) / 16 * 16;
    // (1) aggregate part: D[(s,o), r] = sum_i scale_s(i) dO[i,o] A_t[i,r]   (M = 3 FoP, N = 4F, K = n)
    {
        VScaledGradT al{dO, scal, n, F_out, FoP, towers, 0};
        VCat5T bl{x, A, n, F, towers, 0, F, 4 * F};
        EPostWgrad ep{dwpost, wpost_tower_stride, F, F_out, FoP};
        const int BN = pick_bn(4 * F);
        const int64_t tiles = gnn_ceil_div(4 * F, BN) * gnn_ceil_div(3 * FoP, BM) * towers;
        int64_t k_chunk;
        int splits, rc;
        DISPATCH_MODE(precision, {
            splitk(n, tiles, MODE == MODE_TF32X3 ? 32 : 64, k_chunk, splits);
            rc = (launch<MODE>(al, bl, ep, 3 * FoP, 4 * F, BN, n, k_chunk, splits, towers, st));
        });
        if (rc) return rc;
    }
    // (2) x part: dWpost_t[o, :F] += dO_t^T x
    for (int t = 0; t < towers; ++t) {
        const float* d = dO + t * F_out;
        VTransposed al{d, (int64_t)towers * F_out, F_out, n, 0};
        VTransposed bl{x, F, F, n, 0};
        EAtomic ep{dwpost + t * wpost_tower_stride, 13 * (int64_t)F, 0};
        const int BN = pick_bn(F);
        const int64_t tiles = gnn_ceil_div(F, BN);
        int64_t k_chunk;
        int splits, rc;
        DISPATCH_MODE(precision, {
            splitk(n, tiles, MODE == MODE_TF32X3 ? 32 : 64, k_chunk, splits);
            rc = (launch<MODE>(al, bl, ep, F_out, F, BN, n, k_chunk, splits, 1, st));
        });
        if (rc) return rc;
        rc = gnn_colsum_accum(d, (int64_t)towers * F_out, n, F_out, dbpost + t * bpost_tower_stride, stream);
        if (rc) return rc;
    }
    return 0;
}
