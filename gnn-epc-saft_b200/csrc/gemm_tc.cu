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
constexpr int ROW_BYTES = 128;   // bytes of K per row per k-block (32 tf32 or 64 bf16) = one SWIZZLE_128B row
constexpr int CHUNKS = ROW_BYTES / 16;
// warp roles of the persistent CTA
constexpr int EPI_WARPS = 4;                 // warps 0-3: TMEM -> registers -> smem -> coalesced global
constexpr int MMA_WARP = 4;                  // warp 4: one elected lane issues tcgen05.mma
constexpr int PROD_GROUPS = 4;               // warps 5-20: four producer groups take k-blocks round-robin
constexpr int PROD_GROUP_THREADS = 128;
constexpr int THREADS = (EPI_WARPS + 1) * 32 + PROD_GROUPS * PROD_GROUP_THREADS;   // 416
constexpr int STG_COLS = 32, STG_LD = 36;    // epilogue staging block per warp: 32 rows x 32 columns (+pad)
enum { MODE_TF32X3 = 0, MODE_BF16 = 1 };

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// TMA bulk copy global -> shared, completion (bytes) signalled on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    const uint32_t addr = smem_u32(bar);
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x989680;\n\t"
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

// fp32 -> tf32 hi + tf32 lo, both rounded to nearest (cvt.rna): v = hi + lo + O(2^-22 |v|)
__device__ __forceinline__ float tf32_rn(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ void split_tf32(float4 v, float4& hi, float4& lo) {
    hi = make_float4(tf32_rn(v.x), tf32_rn(v.y), tf32_rn(v.z), tf32_rn(v.w));
    lo = make_float4(tf32_rn(v.x - hi.x), tf32_rn(v.y - hi.y), tf32_rn(v.z - hi.z), tf32_rn(v.w - hi.w));
}

// shared-memory matrix descriptor, K-major, SWIZZLE_128B (cute::UMMA::SmemDescriptor bit layout):
// [0,14) start>>4 | [16,30) leading byte offset>>4 (unused with swizzle, 1) | [32,46) stride byte offset>>4
// (8 rows x 128 B = 1024) | [46,48) version = 1 | [61,64) layout type = 2 (SWIZZLE_128B).
// Tiles are [rows][128 B] with the 16-byte chunk index XOR-ed with (row & 7); 1024-byte aligned.
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint32_t sw128_offset(int row, int chunk) {
    return (uint32_t)row * ROW_BYTES + (uint32_t)((chunk ^ (row & 7)) << 4);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c=F32, a/b format, K-major both, N>>3, M>>4
__host__ __device__ inline uint32_t instr_desc(int mode, int n) {
    uint32_t fmt = mode == MODE_TF32X3 ? 2u : 1u;   // TF32 = 2, BF16 = 1
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// ------------------------------------------------------------------ operand functors
// get4(row, k): elements (row, k..k+3) of the K-major operand view, zero outside [rows) x [cols)
struct VRowMajor {          // element (r, k) = p[r*ld + k]
    static constexpr bool kTransposed = false;   // lanes of a producer warp run along K
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
    static constexpr bool kTransposed = true;   // lanes of a producer warp run along rows
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
    static constexpr bool kTransposed = false;   // lanes of a producer warp run along K
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
    static constexpr bool kTransposed = true;   // lanes of a producer warp run along rows
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
    static constexpr bool kTransposed = false;   // lanes of a producer warp run along K
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
    static constexpr bool kTransposed = false;   // lanes of a producer warp run along K
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
        if (i >= n || k >= 3 * (int64_t)F_out) return make_float4(0.f, 0.f, 0.f, 0.f);
        if ((F_out & 3) == 0) {                                    // four columns of one scaler group
            const int s = (int)k / F_out, o = (int)k - s * F_out;
            float4 v = ldg4(dO + i * ((int64_t)towers * F_out) + tower * F_out + o);
            if (s == 0) return v;
            const float f = __ldg(scal + 2 * i + (s - 1));
            return make_float4(v.x * f, v.y * f, v.z * f, v.w * f);
        }
        return make_float4(elem(i, k), elem(i, k + 1), elem(i, k + 2), elem(i, k + 3));
    }
};
// the same, transposed: row r = s*FoP + o, k = node (wgrad A operand)
struct VScaledGradT {
    static constexpr bool kTransposed = true;   // lanes of a producer warp run along rows
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
    static constexpr bool kTransposed = true;   // lanes of a producer warp run along rows
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

// ------------------------------------------------------------------ epilogues
// emit4(m, n, N, v): four consecutive output columns n..n+3 of row m (n % 4 == 0); lanes of a warp
// cover consecutive columns, so the vector paths are fully coalesced.
struct EStore {     // Y[m, n] = act(v + bias[n]) (+ Y)
    float* y;
    int64_t ldy, y_group_stride;
    const float* bias;
    int64_t bias_group_stride;
    int act, accumulate, vec;
    static constexpr int kCombine3 = 0;
    __device__ __forceinline__ void set_group(int g) {
        y += (int64_t)g * y_group_stride;
        if (bias) bias += (int64_t)g * bias_group_stride;
    }
    __device__ __forceinline__ void emit4(int64_t m, int n, int N, float4 v) const {
        float* dst = y + m * ldy + n;
        if (vec && n + 3 < N) {
            if (bias) {
                float4 b = ldg4(bias + n);
                v = make_float4(v.x + b.x, v.y + b.y, v.z + b.z, v.w + b.w);
            }
            if (act) v = make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
            if (accumulate) {
                float4 o = *reinterpret_cast<const float4*>(dst);
                v = make_float4(v.x + o.x, v.y + o.y, v.z + o.z, v.w + o.w);
            }
            *reinterpret_cast<float4*>(dst) = v;
            return;
        }
        const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (n + j < N) {
                float t = e[j];
                if (bias) t += __ldg(bias + n + j);
                if (act) t = fmaxf(t, 0.f);
                if (accumulate) t += dst[j];
                dst[j] = t;
            }
        }
    }
};
struct EAtomic {    // dW[m, n] += v
    float* y;
    int64_t ldy, y_group_stride;
    int vec;
    static constexpr int kCombine3 = 0;
    __device__ __forceinline__ void set_group(int g) { y += (int64_t)g * y_group_stride; }
    __device__ __forceinline__ void emit4(int64_t m, int n, int N, float4 v) const {
        float* dst = y + m * ldy + n;
        if (vec && n + 3 < N) {
            atomicAdd(reinterpret_cast<float4*>(dst), v);
            return;
        }
        const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (n + j < N) atomicAdd(dst + j, e[j]);
    }
};
// post forward: O[i, t*F_out + o] = act(D[i, o] + amp_i D[i, FoP+o] + att_i D[i, 2FoP+o] + b_t[o]);
// the kernel combines the three scaler groups, emit4 adds bias / activation and stores.
struct EPostCombine {
    float* O;
    const float* scal;
    const float* bias;
    int64_t bias_tower_stride;
    int F_out, FoP, towers, tower, act, vec;
    static constexpr int kCombine3 = 1;
    __device__ __forceinline__ void set_group(int g) { tower = g; bias += (int64_t)g * bias_tower_stride; }
    __device__ __forceinline__ void scalers(int64_t i, float& amp, float& att) const {
        amp = __ldg(scal + 2 * i);
        att = __ldg(scal + 2 * i + 1);
    }
    __device__ __forceinline__ void emit4(int64_t i, int o, int N, float4 v) const {
        float* dst = O + i * ((int64_t)towers * F_out) + tower * F_out + o;
        if (vec && o + 3 < F_out) {
            float4 b = ldg4(bias + o);
            v = make_float4(v.x + b.x, v.y + b.y, v.z + b.z, v.w + b.w);
            if (act) v = make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
            *reinterpret_cast<float4*>(dst) = v;
            return;
        }
        const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (o + j < F_out) {
                float t = e[j] + __ldg(bias + o + j);
                dst[j] = act ? fmaxf(t, 0.f) : t;
            }
        }
    }
};
// post wgrad (A part): D[(s,o), r] -> dWpost_t[o, F + s*4F + r]  (rows are s*FoP + o); F % 4 == 0
struct EPostWgrad {
    float* dw;
    int64_t tower_stride;
    int F, F_out, FoP, vec;
    static constexpr int kCombine3 = 0;
    __device__ __forceinline__ void set_group(int g) { dw += (int64_t)g * tower_stride; }
    __device__ __forceinline__ void emit4(int64_t m, int n, int N, float4 v) const {
        int s = (int)m / FoP, o = (int)m - s * FoP;
        if (s >= 3 || o >= F_out) return;
        float* dst = dw + (int64_t)o * 13 * F + F + (int64_t)s * 4 * F + n;
        if (vec && n + 3 < N) {
            atomicAdd(reinterpret_cast<float4*>(dst), v);
            return;
        }
        const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (n + j < N) atomicAdd(dst + j, e[j]);
    }
};

// ------------------------------------------------------------------ geometry of one launch
struct Geom {
    int64_t M, K, k_chunk, num_items;
    int N, BN, splits, groups, m_tiles, n_tiles, stages;
    uint32_t tmem_cols;
};

// ------------------------------------------------------------------ operand producers
// convert four fp32 values and write one 16-byte chunk (TF32X3: hi plane and lo plane)
template <int MODE>
__device__ __forceinline__ void put_chunk_tf32(uint8_t* tile, uint32_t plane_bytes, uint32_t off, float4 v) {
    float4 hi, lo;
    split_tf32(v, hi, lo);
    *reinterpret_cast<float4*>(tile + off) = hi;
    *reinterpret_cast<float4*>(tile + plane_bytes + off) = lo;
}
__device__ __forceinline__ void put_chunk_bf16(uint8_t* tile, uint32_t off, float4 a, float4 b) {
    __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y), p1 = __floats2bfloat162_rn(a.z, a.w);
    __nv_bfloat162 p2 = __floats2bfloat162_rn(b.x, b.y), p3 = __floats2bfloat162_rn(b.z, b.w);
    *reinterpret_cast<uint4*>(tile + off) = make_uint4(*reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1),
                                                       *reinterpret_cast<uint32_t*>(&p2), *reinterpret_cast<uint32_t*>(&p3));
}
__device__ __forceinline__ float4 mask_k(float4 v, int64_t k, int64_t k_end) {   // zero elements at k >= k_end
    if (k + 3 < k_end) return v;
    if (k >= k_end) v.x = 0.f;
    if (k + 1 >= k_end) v.y = 0.f;
    if (k + 2 >= k_end) v.z = 0.f;
    if (k + 3 >= k_end) v.w = 0.f;
    return v;
}

// Fill a [rows x 128 B] SWIZZLE_128B tile with operand rows row0.. and K range [kbase, kbase + KB) clipped to k_end.
// p = thread index inside the 128-thread producer group.
template <int MODE, class L>
__device__ __forceinline__ void produce_tile(const L& ld, uint8_t* tile, uint32_t plane_bytes, int rows, int64_t row0,
                                             int64_t kbase, int64_t k_end, int p) {
    constexpr int EPC = MODE == MODE_TF32X3 ? 4 : 8;     // elements per 16-byte chunk
    if (!L::kTransposed) {
        // 8 lanes sweep the 8 chunks of a row (128 contiguous bytes in memory), 16 rows per pass.
        // All loads of the tile (up to 8 rows per thread) are issued before the first conversion.
        const int c = p & 7, r0 = p >> 3;
        const int64_t k = kbase + c * EPC;
        constexpr int NB = MODE == MODE_TF32X3 ? 8 : 4;        // rows per thread per batch (register budget)
        for (int rb = 0; rb < rows; rb += 16 * NB) {
            float4 v[NB], w[MODE == MODE_BF16 ? NB : 1];
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                const int r = rb + r0 + 16 * i;
                const bool ok = r < rows && k < k_end;
                v[i] = ok ? mask_k(ld.get4(row0 + r, k), k, k_end) : make_float4(0.f, 0.f, 0.f, 0.f);
                if (MODE == MODE_BF16)
                    w[i] = (ok && k + 4 < k_end) ? mask_k(ld.get4(row0 + r, k + 4), k + 4, k_end)
                                                 : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                const int r = rb + r0 + 16 * i;
                if (r < rows) {
                    if (MODE == MODE_TF32X3) put_chunk_tf32<MODE>(tile, plane_bytes, sw128_offset(r, c), v[i]);
                    else put_chunk_bf16(tile, sw128_offset(r, c), v[i], w[MODE == MODE_BF16 ? i : 0]);
                }
            }
        }
    } else {
        // lanes run along rows (the contiguous dimension of a transposed operand); each thread owns whole rows
        constexpr int NB = MODE == MODE_TF32X3 ? CHUNKS : CHUNKS / 2;
        for (int r = p; r < rows; r += PROD_GROUP_THREADS) {
#pragma unroll
            for (int cb = 0; cb < CHUNKS; cb += NB) {
                float4 v[NB], w[MODE == MODE_BF16 ? NB : 1];
#pragma unroll
                for (int i = 0; i < NB; ++i) {
                    const int64_t k = kbase + (cb + i) * EPC;
                    v[i] = k < k_end ? mask_k(ld.get4(row0 + r, k), k, k_end) : make_float4(0.f, 0.f, 0.f, 0.f);
                    if (MODE == MODE_BF16)
                        w[i] = k + 4 < k_end ? mask_k(ld.get4(row0 + r, k + 4), k + 4, k_end)
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int i = 0; i < NB; ++i) {
                    if (MODE == MODE_TF32X3) put_chunk_tf32<MODE>(tile, plane_bytes, sw128_offset(r, cb + i), v[i]);
                    else put_chunk_bf16(tile, sw128_offset(r, cb + i), v[i], w[MODE == MODE_BF16 ? i : 0]);
                }
            }
        }
    }
}

// ------------------------------------------------------------------ the kernel
// Persistent, warp-specialised: work item = (m tile, n tile, group, k split).
//   producers (4 groups x 4 warps): global -> registers -> (tf32 hi/lo | bf16) -> SWIZZLE_128B smem ring
//   MMA warp: waits full[s], one lane issues tcgen05.mma (3 per K step for TF32X3), tcgen05.commit -> empty[s];
//             after the last k-block of an item a commit -> tmem_full[b]
//   epilogue warps: wait tmem_full[b], tcgen05.ld, epilogue math, smem transpose, coalesced 128-bit stores,
//             arrive tmem_empty[b]   (two accumulator buffers in TMEM: epilogue of item i overlaps item i+1)
// grid (k-blocks, n tiles, groups): writes the B operand of every (group, n tile, k-block) as the exact
// shared-memory image (SWIZZLE_128B, hi/lo planes) so that the GEMM can fetch it with one TMA bulk copy
template <int MODE, class BL>
__global__ void __launch_bounds__(PROD_GROUP_THREADS)
pack_b_kernel(const BL bl0, int BN, int64_t K, int n_tiles, int num_kb, uint8_t* __restrict__ out) {
    constexpr int KB_ELEMS = MODE == MODE_TF32X3 ? 32 : 64;
    constexpr int PARTS = MODE == MODE_TF32X3 ? 2 : 1;
    const int kb = blockIdx.x, nt = blockIdx.y, grp = blockIdx.z;
    BL bl = bl0;
    bl.set_group(grp);
    const uint32_t b_plane = (uint32_t)BN * ROW_BYTES;
    uint8_t* tile = out + ((size_t)(grp * n_tiles + nt) * num_kb + kb) * (size_t)(PARTS * b_plane);
    produce_tile<MODE>(bl, tile, b_plane, BN, (int64_t)nt * BN, (int64_t)kb * KB_ELEMS, K, threadIdx.x);
}

template <int MODE, bool PACKED, class AL, class BL, class EP>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tc_kernel(const AL al0, const BL bl0, const EP ep0, const Geom g, const uint8_t* __restrict__ packed_b) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[4], empty_bar[4], tmem_full_bar[2], tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_s;

    constexpr int KB_ELEMS = MODE == MODE_TF32X3 ? 32 : 64;
    constexpr int PARTS = MODE == MODE_TF32X3 ? 2 : 1;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int BN = g.BN, NS = g.stages;

    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t a_plane = BM * ROW_BYTES, b_plane = (uint32_t)BN * ROW_BYTES;
    const uint32_t stage_bytes = PARTS * (a_plane + b_plane);        // multiple of 1024 (BN % 16 == 0)
    float* stg_all = reinterpret_cast<float*>(smem + (size_t)NS * stage_bytes);

    if (warp == MMA_WARP) tmem_alloc(&tmem_base_s, g.tmem_cols);
    if (tid == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(&full_bar[s], PROD_GROUP_THREADS);
            mbar_init(&empty_bar[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tmem_full_bar[b], 1);
            mbar_init(&tmem_empty_bar[b], EPI_WARPS * 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    auto decode = [&](int64_t item, int64_t& m0, int& n0, int& grp, int64_t& k_begin, int64_t& k_end) {
        int64_t mt = item % g.m_tiles, rest = item / g.m_tiles;
        int nt = (int)(rest % g.n_tiles);
        rest /= g.n_tiles;
        grp = (int)(rest % g.groups);
        int split = (int)(rest / g.groups);
        m0 = mt * BM;
        n0 = nt * BN;
        k_begin = (int64_t)split * g.k_chunk;
        k_end = min(g.K, k_begin + g.k_chunk);
    };

    if (warp < EPI_WARPS) {
        // =============================== epilogue ===============================
        float* stg = stg_all + warp * (32 * STG_LD);
        uint32_t t = 0;
        for (int64_t item = blockIdx.x; item < g.num_items; item += gridDim.x, ++t) {
            int64_t m0, k_begin, k_end;
            int n0, grp;
            decode(item, m0, n0, grp, k_begin, k_end);
            EP ep = ep0;
            ep.set_group(grp);
            const uint32_t b = t & 1, tuse = t >> 1;
            mbar_wait(&tmem_full_bar[b], tuse & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + b * (uint32_t)BN + ((uint32_t)(warp * 32) << 16);
            const int64_t m_thread = m0 + warp * 32 + lane;
            const int out_cols = EP::kCombine3 ? BN / 3 : BN;          // columns that are written out
            float amp = 0.f, att = 0.f;
            if constexpr (EP::kCombine3) {
                if (m_thread < g.M) ep.scalers(m_thread, amp, att);
            }
            for (int cb = 0; cb < out_cols; cb += STG_COLS) {
                // thread = one row: 32 columns of the block into the staging tile
#pragma unroll
                for (int h = 0; h < STG_COLS / 16; ++h) {
                    const int col = cb + h * 16;
                    float v[16];
                    if (col < out_cols) {
                        tmem_ld16(taddr + col, v);
                        if constexpr (EP::kCombine3) {
                            float d1[16], d2[16];
                            tmem_ld16(taddr + out_cols + col, d1);
                            tmem_ld16(taddr + 2 * out_cols + col, d2);
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[j] = v[j] + amp * d1[j] + att * d2[j];
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] = 0.f;
                    }
#pragma unroll
                    for (int j = 0; j < 16; j += 4)
                        *reinterpret_cast<float4*>(stg + lane * STG_LD + h * 16 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                }
                __syncwarp();
                // write-out: 8 lanes per row (128 contiguous bytes), 4 rows per pass
                const int cq = (lane & 7) * 4, rsub = lane >> 3;
#pragma unroll 2
                for (int rr = 0; rr < 32; rr += 4) {
                    const int r = rr + rsub;
                    const int64_t m = m0 + warp * 32 + r;
                    const int col = cb + cq;
                    if (m < g.M && col < out_cols) {
                        float4 v = *reinterpret_cast<const float4*>(stg + r * STG_LD + cq);
                        if constexpr (EP::kCombine3) ep.emit4(m, (int)(n0 / 3) + col, g.N, v);
                        else if (n0 + col < g.N) ep.emit4(m, n0 + col, g.N, v);
                    }
                }
                __syncwarp();
            }
            tc_fence_before();
            mbar_arrive(&tmem_empty_bar[b]);
        }
    } else if (warp == MMA_WARP) {
        // =============================== MMA issuer ===============================
        const uint32_t idesc = instr_desc(MODE, BN);
        uint32_t it = 0, t = 0;
        for (int64_t item = blockIdx.x; item < g.num_items; item += gridDim.x, ++t) {
            int64_t m0, k_begin, k_end;
            int n0, grp;
            decode(item, m0, n0, grp, k_begin, k_end);
            const int num_kb = (int)((k_end - k_begin + KB_ELEMS - 1) / KB_ELEMS);
            const uint32_t b = t & 1, tuse = t >> 1;
            if (tuse > 0) mbar_wait(&tmem_empty_bar[b], (tuse - 1) & 1);     // epilogue drained this accumulator
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + b * (uint32_t)BN;
            for (int kb = 0; kb < num_kb; ++kb, ++it) {
                const uint32_t s = it % NS, use = it / NS;
                mbar_wait(&full_bar[s], use & 1);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t a_hi = smem_u32(smem + (size_t)s * stage_bytes), a_lo = a_hi + a_plane;
                    const uint32_t b_hi = a_hi + PARTS * a_plane, b_lo = b_hi + b_plane;
#pragma unroll
                    for (int j = 0; j < CHUNKS / 2; ++j) {       // one MMA consumes 32 bytes of K
                        const uint32_t koff = j * 32;
                        const uint32_t first = (kb == 0 && j == 0) ? 0u : 1u;
                        if (MODE == MODE_TF32X3) {
                            umma<MODE>(tmem_d, smem_desc_sw128(a_lo + koff), smem_desc_sw128(b_hi + koff), idesc, first);
                            umma<MODE>(tmem_d, smem_desc_sw128(a_hi + koff), smem_desc_sw128(b_lo + koff), idesc, 1u);
                            umma<MODE>(tmem_d, smem_desc_sw128(a_hi + koff), smem_desc_sw128(b_hi + koff), idesc, 1u);
                        } else {
                            umma<MODE>(tmem_d, smem_desc_sw128(a_hi + koff), smem_desc_sw128(b_hi + koff), idesc, first);
                        }
                    }
                    umma_commit(&empty_bar[s]);                       // smem stage reusable once these MMAs retire
                    if (kb == num_kb - 1) umma_commit(&tmem_full_bar[b]);
                }
                __syncwarp();
            }
        }
    } else {
        // =============================== producers ===============================
        // NS (<= PROD_GROUPS) groups are active and group pg owns ring stage pg: a group then only ever waits
        // for the MMAs of its own previous k-block, so it can never run two mbarrier phases ahead (a parity
        // wait cannot tell phase n from phase n-2).  Groups pg >= NS idle.
        const int pg = (warp - (EPI_WARPS + 1)) >> 2;            // producer group 0..3
        const int p = tid - (EPI_WARPS + 1) * 32 - pg * PROD_GROUP_THREADS;
        const int total_kb = (int)((g.K + KB_ELEMS - 1) / KB_ELEMS);      // packed images exist only without split-K
        uint32_t it = 0;
        for (int64_t item = blockIdx.x; item < g.num_items; item += gridDim.x) {
            int64_t m0, k_begin, k_end;
            int n0, grp;
            decode(item, m0, n0, grp, k_begin, k_end);
            AL al = al0;
            BL bl = bl0;
            al.set_group(grp);
            bl.set_group(grp);
            const int num_kb = (int)((k_end - k_begin + KB_ELEMS - 1) / KB_ELEMS);
            for (int kb = 0; kb < num_kb; ++kb, ++it) {
                if ((int)(it % NS) != pg) continue;                // group pg owns ring stage pg (see below)
                const uint32_t s = it % NS, use = it / NS;
                if (use > 0) mbar_wait(&empty_bar[s], (use - 1) & 1);
                uint8_t* stage = smem + (size_t)s * stage_bytes;
                const int64_t kbase = k_begin + (int64_t)kb * KB_ELEMS;
                if (PACKED) {
                    // the weight operand was pre-packed as the smem image: one TMA bulk copy per k-block
                    if (p == 0) {
                        const uint32_t bytes = PARTS * b_plane;
                        const size_t tile_idx = ((size_t)(grp * g.n_tiles + n0 / BN) * total_kb + kb);
                        mbar_expect_tx(&full_bar[s], bytes);
                        bulk_g2s(stage + PARTS * a_plane, packed_b + tile_idx * bytes, bytes, &full_bar[s]);
                    }
                    produce_tile<MODE>(al, stage, a_plane, BM, m0, kbase, k_end, p);
                } else {
                    produce_tile<MODE>(al, stage, a_plane, BM, m0, kbase, k_end, p);
                    produce_tile<MODE>(bl, stage + PARTS * a_plane, b_plane, BN, (int64_t)n0, kbase, k_end, p);
                }
                fence_async_smem();                                 // generic-proxy writes -> async (tensor core) proxy
                mbar_arrive(&full_bar[s]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc(tmem_base, g.tmem_cols);
}

// ------------------------------------------------------------------ host side
static const size_t kSmemLimit = 232448;     // 227 KB opt-in maximum per CTA

template <int MODE, class AL, class BL, class EP>
static int launch(AL al, BL bl, EP ep, int64_t M, int N, int BN, int64_t K, int64_t k_chunk, int splits, int groups,
                  cudaStream_t st, void* pack_ws = nullptr, int64_t pack_ws_bytes = 0) {
    if (M <= 0 || N <= 0 || K <= 0) return 0;
    if (BN % 16 != 0 || BN < 16 || BN > 256) return GNN_ERR_UNSUPPORTED;
    constexpr int parts = MODE == MODE_TF32X3 ? 2 : 1;
    constexpr int kb_elems = MODE == MODE_TF32X3 ? 32 : 64;
    const size_t stage_bytes = (size_t)parts * (BM + BN) * ROW_BYTES;
    const size_t fixed = 4 * 32 * STG_LD * sizeof(float) + 1024 + 256;
    int stages = (int)((kSmemLimit - fixed) / stage_bytes);
    if (stages > PROD_GROUPS) stages = PROD_GROUPS;      // one producer group per ring stage
    if (stages < 2) return GNN_ERR_UNSUPPORTED;
    Geom g;
    g.M = M; g.K = K; g.k_chunk = k_chunk; g.N = N; g.BN = BN; g.splits = splits; g.groups = groups;
    g.m_tiles = (int)gnn_ceil_div(M, BM);
    g.n_tiles = EP::kCombine3 ? N / BN : (int)gnn_ceil_div(N, BN);
    g.num_items = (int64_t)g.m_tiles * g.n_tiles * groups * splits;
    g.stages = stages;
    uint32_t cols = 32;
    while ((int)cols < 2 * BN) cols <<= 1;
    g.tmem_cols = cols;
    const size_t smem = (size_t)stages * stage_bytes + fixed;
    int64_t grid = g.num_items < gnn_sm_count() ? g.num_items : gnn_sm_count();

    // weights (B) packed once per call when a workspace is given and the M loop re-uses them
    const int num_kb = (int)gnn_ceil_div(K, kb_elems);
    const int64_t pack_bytes = (int64_t)groups * g.n_tiles * num_kb * parts * BN * ROW_BYTES;
    const bool packed = pack_ws != nullptr && splits == 1 && g.m_tiles >= 2 && pack_bytes <= pack_ws_bytes &&
                        (reinterpret_cast<uintptr_t>(pack_ws) & 127) == 0;
    if (packed) {
        dim3 pgrid((unsigned)num_kb, (unsigned)g.n_tiles, (unsigned)groups);
        pack_b_kernel<MODE, BL><<<pgrid, PROD_GROUP_THREADS, 0, st>>>(bl, BN, K, g.n_tiles, num_kb,
                                                                      reinterpret_cast<uint8_t*>(pack_ws));
        auto kern = gemm_tc_kernel<MODE, true, AL, BL, EP>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        kern<<<(unsigned)grid, THREADS, smem, st>>>(al, bl, ep, g, reinterpret_cast<const uint8_t*>(pack_ws));
    } else {
        auto kern = gemm_tc_kernel<MODE, false, AL, BL, EP>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        kern<<<(unsigned)grid, THREADS, smem, st>>>(al, bl, ep, g, nullptr);
    }
    return gnn_launch_status();
}

static int pick_bn(int n) {
    int bn = (n + 15) / 16 * 16;
    return bn > 256 ? 256 : bn;
}
static int is_vec(const void* p, int64_t ld) { return (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(p) & 15) == 0); }

// split-K geometry for weight gradients (K = rows): about one work item per SM, every split non-empty
static void splitk(int64_t K, int64_t tiles, int kb_elems, int64_t& k_chunk, int& splits) {
    int64_t target = (int64_t)gnn_sm_count();
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
                                 int N, int K, int act, float* Y, int64_t ldy, int precision, void* pack_ws,
                                 int64_t pack_ws_bytes, void* stream) {
    if (M < 0 || N < 1 || K < 1) return GNN_ERR_BAD_ARG;
    VRowMajor al{X, ldx, M, K, 0, is_vec(X, ldx)};
    VRowMajor bl{W, ldw, N, K, 0, is_vec(W, ldw)};
    EStore ep{Y, ldy, 0, bias, 0, act, 0, is_vec(Y, ldy) && (bias == nullptr || is_vec(bias, 0))};
    const int BN = pick_bn(N);
    DISPATCH_MODE(precision, return (launch<MODE>(al, bl, ep, M, N, BN, K, K, 1, 1, gnn_stream(stream), pack_ws,
                                                  pack_ws_bytes)));
}

extern "C" int gnn_tc_linear_dgrad(const float* dY, int64_t lddy, const float* W, int64_t ldw, int64_t M, int N, int K,
                                   int accumulate, float* dX, int64_t lddx, int precision, void* pack_ws,
                                   int64_t pack_ws_bytes, void* stream) {
    if (M < 0 || N < 1 || K < 1) return GNN_ERR_BAD_ARG;
    VRowMajor al{dY, lddy, M, N, 0, is_vec(dY, lddy)};     // (m, n) reduction over n
    VTransposed bl{W, ldw, K, N, 0};                       // row kk, k = n -> W[n, kk]
    EStore ep{dX, lddx, 0, nullptr, 0, 0, accumulate, is_vec(dX, lddx)};
    const int BN = pick_bn(K);
    DISPATCH_MODE(precision, return (launch<MODE>(al, bl, ep, M, K, BN, N, N, 1, 1, gnn_stream(stream), pack_ws,
                                                  pack_ws_bytes)));
}

extern "C" int gnn_tc_linear_wgrad(const float* dY, int64_t lddy, const float* X, int64_t ldx, int64_t M, int N, int K,
                                   float* dW, int64_t lddw, int precision, void* stream) {
    if (M < 0 || N < 1 || K < 1) return GNN_ERR_BAD_ARG;
    VTransposed al{dY, lddy, N, M, 0};                     // row o, k = i -> dY[i, o]
    VTransposed bl{X, ldx, K, M, 0};                       // row kk, k = i -> X[i, kk]
    EAtomic ep{dW, lddw, 0, is_vec(dW, lddw)};
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
                                   void* pack_ws, int64_t pack_ws_bytes, void* stream) {
    if (n < 0 || F < 4 || (F & 3) || F_out < 1 || towers < 1) return GNN_ERR_BAD_ARG;
    int FoP = (F_out + 15) / 16 * 16;      // output-feature tile: three scaler groups of FoP columns per CTA
    if (FoP > 64) FoP = 64;
    const int tiles = (F_out + FoP - 1) / FoP;
    VCat5 al{x, A, n, F, towers, 0};
    VPostWeight bl{wpost, wpost_tower_stride, F, F_out, FoP};
    EPostCombine ep{O, scal, bpost, bpost_tower_stride, F_out, FoP, towers, 0, act,
                    (F_out % 4 == 0) && is_vec(O, 0) && is_vec(bpost, bpost_tower_stride)};
    DISPATCH_MODE(precision, return (launch<MODE>(al, bl, ep, n, tiles * 3 * FoP, 3 * FoP, 5 * (int64_t)F,
                                                  5 * (int64_t)F, 1, towers, gnn_stream(stream), pack_ws, pack_ws_bytes)));
}

extern "C" int gnn_tc_pna_post_dgrad(const float* dO, const float* scal, const float* wpost, int64_t wpost_tower_stride,
                                     int64_t n, int F, int F_out, int towers, float* dx, float* gA, int precision,
                                     void* pack_ws, int64_t pack_ws_bytes, void* stream) {
    if (n < 0 || F < 4 || (F & 3) || F_out < 1 || towers < 1) return GNN_ERR_BAD_ARG;
    cudaStream_t st = gnn_stream(stream);
    // gA[:, t, :] = [dO_t | amp dO_t | att dO_t] . Wslices_t          (M = n, N = 4F, K = 3 F_out)
    {
        VScaledGrad al{dO, scal, n, F_out, towers, 0};
        VPostWeightT bl{wpost, wpost_tower_stride, F, F_out};
        EStore ep{gA, (int64_t)towers * 4 * F, 4 * (int64_t)F, nullptr, 0, 0, 0, is_vec(gA, 0)};
        const int BN = pick_bn(4 * F);
        int rc;
        DISPATCH_MODE(precision, rc = (launch<MODE>(al, bl, ep, n, 4 * F, BN, 3 * (int64_t)F_out, 3 * (int64_t)F_out, 1,
                                                    towers, st, pack_ws, pack_ws_bytes)));
        if (rc) return rc;
    }
    // dx += sum_t dO_t . Wpost_t[:, :F]     (towers accumulate into the same dx -> sequential launches)
    for (int t = 0; t < towers; ++t) {
        const float* d = dO + t * F_out;
        const float* w = wpost + t * wpost_tower_stride;
        VRowMajor al{d, (int64_t)towers * F_out, n, F_out, 0, is_vec(d, (int64_t)towers * F_out)};
        VTransposed bl{w, 13 * (int64_t)F, F, F_out, 0};
        EStore ep{dx, F, 0, nullptr, 0, 0, 1, is_vec(dx, F)};
        const int BN = pick_bn(F);
        int rc;
        DISPATCH_MODE(precision, rc = (launch<MODE>(al, bl, ep, n, F, BN, (int64_t)F_out, (int64_t)F_out, 1, 1, st,
                                                    pack_ws, pack_ws_bytes)));
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
        EPostWgrad ep{dwpost, wpost_tower_stride, F, F_out, FoP, is_vec(dwpost, wpost_tower_stride)};
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
        EAtomic ep{dwpost + t * wpost_tower_stride, 13 * (int64_t)F, 0, is_vec(dwpost + t * wpost_tower_stride, F)};
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
