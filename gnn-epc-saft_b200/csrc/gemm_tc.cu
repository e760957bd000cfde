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

#include <cuda.h>          // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)
#include <cuda_bf16.h>

#include <type_traits>

#include "pna_agg.cuh"

// Per-role cycle counters (tools/tc_probe.py) exist only in a -DGNN_B200_PROFILE build; the product library has
// neither the counters nor the process-global hook that enables them.
#ifdef GNN_B200_PROFILE
#define GNN_PROF(g) ((g).prof)
#else
#define GNN_PROF(g) (static_cast<unsigned long long*>(nullptr))
#endif

namespace tc {

constexpr int BM = 128;          // UMMA M (cta_group::1), one TMEM lane per row
constexpr int ROW_BYTES = 128;   // bytes of K per row per k-block (32 tf32 or 64 bf16) = one SWIZZLE_128B row
constexpr int CHUNKS = ROW_BYTES / 16;
// warp roles of the persistent CTA
// epilogue warps 0..EW-1: TMEM -> registers -> global (warp w: lane quarter w & 3, column phase w >> 2).
// Weight-gradient (MN-major) launches have one light epilogue per CTA and register-hungry producers: 4 warps
// there, the first of which also issues the MMAs (20 warps = 5 per SM sub-partition -> 96 registers per
// thread); 8 epilogue warps + a dedicated MMA warp elsewhere (25 warps -> 72 registers).
// Launches whose producers also compute the PNA aggregates (AggSrc below) need ~96 registers per producer thread as
// well and use the same 20-warp shape (the register file is split over four sub-partitions: a 21st warp would put six
// warps on one of them and cap every thread at 80 registers again).
__host__ __device__ constexpr int epi_warps(bool mn_major, bool agg = false) { return (mn_major || agg) ? 4 : 8; }
                                             // warp EW: one elected lane issues tcgen05.mma
constexpr int PROD_GROUPS = 4;               // warps 5-20: four producer groups take k-blocks round-robin
constexpr int PROD_GROUP_THREADS = 128;
__host__ __device__ constexpr int cta_threads(bool mn_major, bool agg = false) {
    return (epi_warps(mn_major, agg) + ((mn_major || agg) ? 0 : 1)) * 32 + PROD_GROUPS * PROD_GROUP_THREADS;
}
// MODE_TF32X3: three tf32 MMAs per K step (fp32-level accuracy).  MODE_BF16: one bf16 MMA.
enum { MODE_TF32X3 = 0, MODE_BF16 = 1 };
__host__ __device__ constexpr bool tf32_like(int mode) { return mode != MODE_BF16; }

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
// TMA tensor-map load of a 2-D box (SWIZZLE_128B), completion (bytes) signalled on an mbarrier; rows / columns
// outside the tensor arrive as zeros
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int col, int row, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(col), "r"(row)
        : "memory");
}
// 256-bit global store (STG.E.256): one full 32-byte sector per lane per instruction
__device__ __forceinline__ void st_global_v8(float* p, const float* v) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
                 "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
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
__device__ __forceinline__ bool elect_one() {     // one lane of the (converged) warp
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
template <int MODE>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    if (tf32_like(MODE))
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
// D += A . B (accumulate flag as an immediate: no predicate set-up on the single issuing thread)
template <int MODE>
__device__ __forceinline__ void umma_acc(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
    if (tf32_like(MODE))
        asm volatile("tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, 1;" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc)
                     : "memory");
    else
        asm volatile("tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, 1;" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc)
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

// fp32 -> tf32 hi + lo.  hi = x rounded to nearest (ties away) on a 10-bit mantissa with two integer ops
// (cvt.rna.tf32.f32 expands to a long special-case sequence on this target); lo = x - hi is exact in fp32 and the
// tensor core reads its top 19 bits:  x = hi + lo + O(2^-21 |x|).
__device__ __forceinline__ float tf32_rn(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}
__device__ __forceinline__ void split_tf32(float4 v, float4& hi, float4& lo) {
    hi = make_float4(tf32_rn(v.x), tf32_rn(v.y), tf32_rn(v.z), tf32_rn(v.w));
    lo = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
}

// shared-memory matrix descriptor, K-major, SWIZZLE_128B (cute::UMMA::SmemDescriptor bit layout):
// [0,14) start>>4 | [16,30) leading byte offset>>4 (unused with swizzle, 1) | [32,46) stride byte offset>>4
// (8 rows x 128 B = 1024) | [46,48) version = 1 | [61,64) layout type = 2 (SWIZZLE_128B).
// Tiles are [rows][128 B] with the 16-byte chunk index XOR-ed with (row & 7); 1024-byte aligned.
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// MN-major SWIZZLE_128B (the MN dimension is contiguous: 128-byte rows of 32 tf32 / 64 bf16 elements, atoms of
// 8 K-rows x 128 B): leading byte offset = distance between MN groups, stride byte offset = distance between
// groups of 8 K.
// 16-bit types use layout type 2 (atoms of 8 K rows, 16-byte swizzle); tf32 must use SWIZZLE_128B_BASE32B (layout
// type 1: atoms of 4 K rows x 128 B, 32-byte chunks XOR-ed with (k & 3)) - the only MN-major layout for tf32.
__device__ __forceinline__ uint64_t smem_desc_sw128_mn(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                       uint32_t layout_type) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | ((uint64_t)layout_type << 61);
}
__device__ __forceinline__ uint32_t sw128_offset(int row, int chunk) {
    return (uint32_t)row * ROW_BYTES + (uint32_t)((chunk ^ (row & 7)) << 4);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c=F32, a/b format, K-major both, N>>3, M>>4
__host__ __device__ inline uint32_t instr_desc(int mode, int n, int mn_major) {
    uint32_t fmt = mode == MODE_BF16 ? 1u : 2u;     // TF32 = 2, BF16 = 1
    uint32_t major = mn_major ? ((1u << 15) | (1u << 16)) : 0u;   // a_major / b_major: 0 = K, 1 = MN
    return (1u << 4) | (fmt << 7) | (fmt << 10) | major | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
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
// post forward weights.  Output features are tiled FoP at a time (FoP multiple of 16); tile j holds
// rows [j*3*FoP, (j+1)*3*FoP) = three scaler groups of FoP rows: row r -> (tile, s, oo), o = tile*FoP + oo.
//   k <  F : s == 0 ? Wpost_t[o, k] : 0          k >= F : Wpost_t[o, F + s*4F + (k-F)]
// vx != nullptr (folded destination projection, fold.cu): the x columns of ALL three scaler groups come from
//   Vx[t][s][o][k]  instead ( = [s == 0] Wpost_t[o, k] + (Wmean + Wmin + Wmax)_{t,s} . Wdst_t ).
// agg_xpad > 0 selects the K order of the fused aggregate producers (AggSrc): k < agg_xpad are the x columns (zero
// beyond F), then one 32-column k-block per group of 8 features: column cc of k-block j is aggregator (cc >> 3),
// feature 8 j + 4 ((cc >> 2) & 1) + (cc & 3).
struct VPostWeight {
    static constexpr bool kTransposed = false;   // lanes of a producer warp run along K
    const float* w;
    int64_t tower_stride;
    int F, F_out, FoP;
    const float* vx;
    int agg_xpad;
    __device__ __forceinline__ void set_group(int g) {
        w += (int64_t)g * tower_stride;
        if (vx) vx += (int64_t)g * 3 * F_out * F;
    }
    __device__ __forceinline__ float4 get4(int64_t r, int64_t k) const {
        float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        int tile = (int)r / (3 * FoP), rem = (int)r - tile * 3 * FoP;
        int s = rem / FoP, o = tile * FoP + (rem - s * FoP);
        if (agg_xpad > 0) {
            if (o >= F_out || k >= agg_xpad + 4 * (int64_t)F) return z;
            if (k >= agg_xpad) {
                const int kk = (int)k - agg_xpad, j = kk >> 5, cc = kk & 31;
                k = F + (int64_t)(cc >> 3) * F + 8 * j + 4 * ((cc >> 2) & 1);      // cc & 3 == 0: four consecutive features
            } else if (k >= F) {
                return z;
            }
        }
        if (o >= F_out || k >= 5 * (int64_t)F) return z;
        const float* row = w + (int64_t)o * 13 * F;
        if (k < F) {
            if (vx) return ldg4(vx + ((int64_t)s * F_out + o) * F + k);
            if (s != 0) return z;
            return make_float4(__ldg(row + k), __ldg(row + k + 1), __ldg(row + k + 2), __ldg(row + k + 3));
        }
        const float* q = row + F + (int64_t)s * 4 * F + (k - F);
        return make_float4(__ldg(q), __ldg(q + 1), __ldg(q + 2), __ldg(q + 3));
    }
};
// [dO_t | amp dO_t | att dO_t] as K-major A operand: (node i, k = s*F_out + o)
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

// projection weights of PNAConv.pre_nns[t][0] (first two thirds): row n = part*C + t*F + f (part 0 = destination,
// 1 = source), k in [0,F)  ->  Wpre_t[f, part*F + k]
// (first_part, nparts) = (0, 2): both halves; (1, 1): the source half only (folded destination projection)
struct VProjWeight {
    static constexpr bool kTransposed = false;
    const float* w;
    int64_t tower_stride;
    int F, towers, first_part, nparts;
    __device__ __forceinline__ void set_group(int) {}
    __device__ __forceinline__ float4 get4(int64_t r, int64_t k) const {
        const int C = towers * F;
        if (r >= nparts * C || k >= F) return make_float4(0.f, 0.f, 0.f, 0.f);
        const int p0 = (int)r / C, part = first_part + p0, rem = (int)r - p0 * C, t = rem / F, f = rem - t * F;
        return ldg4(w + (int64_t)t * tower_stride + (int64_t)f * 3 * F + part * F + k);
    }
};
// B operand of the fused dx GEMM: row kk in [0,F), k over [dP (2C) | dO (T*F_out)]:
//   k <  2C : (part,t,f) -> Wpre_t[f, part*F + kk]        k >= 2C : (t,o) -> Wpost_t[o, kk]
struct VDxWeight {
    static constexpr bool kTransposed = true;
    const float* wpre;
    const float* wpost;
    int64_t wpre_tower_stride, wpost_tower_stride;
    int F, F_out, towers;
    __device__ __forceinline__ void set_group(int) {}
    __device__ __forceinline__ float elem(int64_t kk, int64_t k) const {
        const int C = towers * F;
        if (k < 2 * C) {
            const int part = (int)k / C, rem = (int)k - part * C, t = rem / F, f = rem - t * F;
            return __ldg(wpre + (int64_t)t * wpre_tower_stride + (int64_t)f * 3 * F + part * F + kk);
        }
        const int j = (int)k - 2 * C;
        if (j >= towers * F_out) return 0.f;
        const int t = j / F_out, o = j - t * F_out;
        return __ldg(wpost + (int64_t)t * wpost_tower_stride + (int64_t)o * 13 * F + kk);
    }
    __device__ __forceinline__ float4 get4(int64_t kk, int64_t k) const {
        if (kk >= F) return make_float4(0.f, 0.f, 0.f, 0.f);
        return make_float4(elem(kk, k), elem(kk, k + 1), elem(kk, k + 2), elem(kk, k + 3));
    }
};

// B operand of the fused dx GEMM with a folded destination projection: row kk in [0,F), k over
// [dPsrc (C) | dO (T*F_out) | amp dO | att dO]:
//   k < C : (t,f) -> Wpre_t[f, F + kk]        k >= C : (s,t,o) -> Vx[t][s][o][kk]
struct VDxWeightFold {
    static constexpr bool kTransposed = true;
    const float* wpre;
    const float* vx;
    int64_t wpre_tower_stride;
    int F, F_out, towers;
    __device__ __forceinline__ void set_group(int) {}
    __device__ __forceinline__ float elem(int64_t kk, int64_t k) const {
        const int C = towers * F;
        if (k < C) {
            const int t = (int)k / F, f = (int)k - t * F;
            return __ldg(wpre + (int64_t)t * wpre_tower_stride + (int64_t)f * 3 * F + F + kk);
        }
        const int j = (int)k - C, TF = towers * F_out;
        if (j >= 3 * TF) return 0.f;
        const int sgrp = j / TF, r = j - sgrp * TF, t = r / F_out, o = r - t * F_out;
        return __ldg(vx + ((int64_t)(t * 3 + sgrp) * F_out + o) * F + kk);
    }
    __device__ __forceinline__ float4 get4(int64_t kk, int64_t k) const {
        if (kk >= F) return make_float4(0.f, 0.f, 0.f, 0.f);
        return make_float4(elem(kk, k), elem(kk, k + 1), elem(kk, k + 2), elem(kk, k + 3));
    }
};

// ------------------------------------------------------------------ activation operands: segmented rows
// Every activation operand of the model is a row-major [node][column] matrix whose columns are a
// concatenation of up to three slices of real tensors, optionally multiplied by a per-node degree scaler:
//   [x | A_t]  (post forward),  [dO_t | amp dO_t | att dO_t]  (post backward),  plain X / dY (Linear).
// A producer thread owns one 16-byte column chunk, resolves its segment once per k-block and then only walks
// rows with pointer arithmetic.  Segment starts / widths are multiples of 4 columns.
struct Seg {
    const float* base;       // element (row 0, first column of the segment), before the group offset
    int64_t ld;              // row stride (elements)
    int64_t group_stride;    // added to base per group (tower)
    int start, width;        // columns [start, start + width) of the operand
    int scale;               // -1: none, 0: amplification, 1: attenuation  (scal[2*row + scale])
};
constexpr int MAX_SEGS = 4;
struct SegOperand {
    Seg seg[MAX_SEGS];
    int nseg;
    int64_t rows;            // valid rows (nodes)
    const float* scal;       // [rows, 2]
};
struct SegRef {
    const float* ptr;        // address of (row 0, col) or nullptr when the column is padding
    int64_t ld;
    int scale;
};
__device__ __forceinline__ SegRef seg_resolve(const SegOperand& op, int64_t col, int grp) {
    SegRef r{nullptr, 0, -1};
#pragma unroll
    for (int s = 0; s < MAX_SEGS; ++s) {
        if (s < op.nseg && col >= op.seg[s].start && col < op.seg[s].start + op.seg[s].width) {
            r.ptr = op.seg[s].base + (int64_t)grp * op.seg[s].group_stride + (col - op.seg[s].start);
            r.ld = op.seg[s].ld;
            r.scale = op.seg[s].scale;
        }
    }
    return r;
}

// ------------------------------------------------------------------ aggregates produced on the fly
// Post forward with the aggregates computed by the GEMM's own producers (folded formulation): instead of reading
// A = [mean | min | max | std] of q = Psrc[src] + tab[code] from HBM (2 KB per node and layer, written by a separate
// kernel), a producer group gathers the source rows of its 128 nodes for 8 features, reduces them in registers
// (pna_agg.cuh - the arithmetic of the stand-alone kernel, bit-identical) and writes the four aggregates straight into
// the SWIZZLE_128B operand stage (tf32 hi / lo planes).  One such pass = one 32-column k-block (4 aggregators x 8
// features); the weight image is packed in the same K order (VPostWeight::agg_xpad).  A_out != NULL (training) also
// stores the aggregates for the backward pass: one 209 MB write instead of a write and a read.
struct NoAgg {};
struct AggSrc {
    const float* P;            // Psrc [n, C], C = towers * F
    const float* tab;          // [60, C]
    const int* rowptr;
    const int* col_src;
    const uint8_t* col_code;
    float* A_out;              // [n, towers, 4, F] or nullptr
    int F, C, xpad;            // xpad: K columns of the x part (F rounded up to the 32-column k-block)
};
// results of one thread - rows r and r + 64 of the tile, four features, four aggregators - live in the eight float4
// of the plain producer's ATile (v[4 * half + aggregator]): the two producer paths never hold values at the same time
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
    static constexpr int kStaged = 1;       // plain stores may go through the shared-memory transposing epilogue
    bool staged_ok(int N) const {           // host: no bias / activation / accumulate, 16-byte aligned rows
        return bias == nullptr && act == 0 && accumulate == 0 && vec && (N & 3) == 0 && ((ldy | y_group_stride) & 3) == 0;
    }
    static constexpr int kWide = 1;
    __device__ __forceinline__ void set_group(int g) {
        y += (int64_t)g * y_group_stride;
        if (bias) bias += (int64_t)g * bias_group_stride;
    }
    // 16 consecutive columns as two 256-bit stores (rows 32-byte aligned, no read-modify-write)
    __device__ __forceinline__ bool wide_ok(int n, int N) const {
        return vec && !accumulate && n + 15 < N && ((ldy | y_group_stride) & 7) == 0 &&
               ((reinterpret_cast<uintptr_t>(y) | (uintptr_t)(n * 4)) & 31) == 0;
    }
    __device__ __forceinline__ void store16(int64_t m, int n, float* v) const {
        if (bias) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
                const float4 b = ldg4(bias + n + j);
                v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
            }
        }
        if (act) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        float* dst = y + m * ldy + n;
        st_global_v8(dst, v);
        st_global_v8(dst + 8, v + 8);
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
// Y[m, n] = act(v + bias[n]) + skip[m, n]   (eval forward: Linear with BatchNorm folded in, ReLU, skip connection)
struct EStoreSkip {
    static constexpr int kWide = 0;
    static constexpr int kCombine3 = 0;
    static constexpr int kStaged = 0;
    float* y;
    int64_t ldy;
    const float* bias;
    const float* skip;
    int64_t ldskip;
    int act;
    __device__ __forceinline__ void set_group(int) {}
    __device__ __forceinline__ void emit4(int64_t m, int n, int N, float4 v) const {   // N % 4 == 0, 16-byte aligned rows
        const float4 b = ldg4(bias + n);
        v = make_float4(v.x + b.x, v.y + b.y, v.z + b.z, v.w + b.w);
        if (act) v = make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
        if (skip != nullptr) {
            const float4 sk = ldg4(skip + m * ldskip + n);
            v = make_float4(v.x + sk.x, v.y + sk.y, v.z + sk.z, v.w + sk.w);
        }
        *reinterpret_cast<float4*>(y + m * ldy + n) = v;
    }
};
struct EAtomic {
    static constexpr int kWide = 0;    // dW[m, n] += v
    float* y;
    int64_t ldy, y_group_stride;
    int vec;
    static constexpr int kCombine3 = 0;
    static constexpr int kStaged = 0;
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
    static constexpr int kWide = 0;
    float* O;
    const float* scal;
    const float* bias;
    int64_t bias_tower_stride;
    int F_out, FoP, towers, tower, act, vec;
    static constexpr int kCombine3 = 1;
    static constexpr int kStaged = 0;
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
    static constexpr int kWide = 0;
    float* dw;
    int64_t tower_stride;
    int F, F_out, FoP, vec;
    static constexpr int kCombine3 = 0;
    static constexpr int kStaged = 0;
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

// fused projection forward: P[i, n] = acc + (n < C ? bpre_{t}[f] : 0), n = part*C + t*F + f
struct EProjStore {
    static constexpr int kWide = 1;
    __device__ __forceinline__ bool wide_ok(int n, int N) const {
        return n + 15 < N && (F & 15) == 0 && (reinterpret_cast<uintptr_t>(P) & 31) == 0;
    }
    __device__ __forceinline__ void store16(int64_t m, int n, float* v) const {
        const int C = towers * F;
        if (n < C) {                                       // F % 16 == 0: the 16 columns share one tower  (source-only P: N == C)
            const int t = n / F, f = n - t * F;
            const float* b = bias + (int64_t)t * bias_tower_stride + f;
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
                const float4 bb = ldg4(b + j);
                v[j] += bb.x; v[j + 1] += bb.y; v[j + 2] += bb.z; v[j + 3] += bb.w;
            }
        }
        float* dst = P + m * ldp + n;
        st_global_v8(dst, v);
        st_global_v8(dst + 8, v + 8);
    }
    float* P;
    const float* bias;
    int64_t bias_tower_stride;
    int F, towers;
    int64_t ldp;
    static constexpr int kCombine3 = 0;
    static constexpr int kStaged = 0;
    __device__ __forceinline__ void set_group(int) {}
    __device__ __forceinline__ void emit4(int64_t m, int n, int N, float4 v) const {
        const int C = towers * F;
        if (n < C) {
            const int t = n / F, f = n - t * F;
            const float* b = bias + (int64_t)t * bias_tower_stride + f;
            v = make_float4(v.x + __ldg(b), v.y + __ldg(b + 1), v.z + __ldg(b + 2), v.w + __ldg(b + 3));
        }
        *reinterpret_cast<float4*>(P + m * ldp + n) = v;     // N = 2C or C, C % 4 == 0
    }
};
// fused projection weight gradient: D[(part,t,f), kk] -> dWpre_t[f, part*F + kk]
struct EProjWgrad {
    static constexpr int kWide = 0;
    float* dw;
    int64_t tower_stride;
    int F, towers, first_part, nparts;
    static constexpr int kCombine3 = 0;
    static constexpr int kStaged = 0;
    __device__ __forceinline__ void set_group(int) {}
    __device__ __forceinline__ void emit4(int64_t m, int n, int N, float4 v) const {
        const int C = towers * F;
        if (m >= nparts * C) return;
        const int p0 = (int)m / C, part = first_part + p0, rem = (int)m - p0 * C, t = rem / F, f = rem - t * F;
        float* dst = dw + (int64_t)t * tower_stride + (int64_t)f * 3 * F + part * F + n;
        if (n + 3 < N) {
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
// bias gradient of a weight-gradient GEMM: db[target(col)] += sum over nodes of the A operand column `col`
// (col < ncols), target = base + group*group_stride + (col / tower_cols)*tower_stride + col % tower_cols
struct ColSum {
    float* base;
    int64_t group_stride, tower_stride;
    int ncols, tower_cols;
};
struct Geom {
    int64_t M, K, k_chunk, num_items;
    int N, BN, splits, groups, m_tiles, n_tiles, stages;
    int kb_rows;                 // K rows (nodes) per ring stage: KB_ELEMS, or half of it for wide MN-major tiles
    uint32_t tmem_cols;
    ColSum colsum;
    unsigned long long* prof;    // optional cycle counters (debug): see gnn_tc_set_profile_buffer
    uint32_t epi_stage_off;      // != 0: byte offset (from the aligned smem base) of 8 x 4 KB epilogue staging tiles
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
    constexpr int EPC = tf32_like(MODE) ? 4 : 8;     // elements per 16-byte chunk
    if (!L::kTransposed) {
        // 8 lanes sweep the 8 chunks of a row (128 contiguous bytes in memory), 16 rows per pass.
        // All loads of the tile (up to 8 rows per thread) are issued before the first conversion.
        const int c = p & 7, r0 = p >> 3;
        const int64_t k = kbase + c * EPC;
        constexpr int NB = tf32_like(MODE) ? 8 : 4;        // rows per thread per batch (register budget)
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
        constexpr int NB = tf32_like(MODE) ? CHUNKS : CHUNKS / 2;
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

// K-major A tile from a segmented operand: 128 rows (nodes m0..) x one k-block of columns.  Thread p owns the
// 16-byte chunk c = p & 7 of rows r0 + 16 i: every load is a 128-bit coalesced access (8 lanes = 128 contiguous
// bytes of one row), all eight are issued before the first conversion, addresses advance by a constant stride and
// smem offsets are compile-time strides.  Full tiles take a branch-free path.
// The tile is produced in two phases so that the global loads of a k-block are already in flight while the producer
// group still waits for its ring stage to be released by the MMAs (the loads never touch shared memory):
//   load_a_seg : resolve segments, issue the eight 16-byte loads (and the degree scalers) - no use of the values
//   store_a_seg: scale, convert and write the swizzled image
template <int MODE, bool SCALED>
struct ATile {
    float4 v[8], w[MODE == MODE_BF16 ? 8 : 1];
    float f0[SCALED ? 8 : 1], f1[(SCALED && MODE == MODE_BF16) ? 8 : 1];
};
template <int MODE, bool SCALED>
__device__ __forceinline__ void load_a_seg(const SegOperand& op, int grp, int64_t m0, int64_t kbase, int64_t k_end, int p,
                                           ATile<MODE, SCALED>& t) {
    constexpr int EPC = tf32_like(MODE) ? 4 : 8;
    const int c = p & 7, r0 = p >> 3;
    const int64_t k = kbase + c * EPC;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    const SegRef s0 = k < k_end ? seg_resolve(op, k, grp) : SegRef{nullptr, 0, -1};
    const SegRef s1 = (MODE == MODE_BF16 && k + 4 < k_end) ? seg_resolve(op, k + 4, grp) : SegRef{nullptr, 0, -1};
    const int64_t row0 = m0 + r0;
    const float* q0 = s0.ptr != nullptr ? s0.ptr + row0 * s0.ld : nullptr;
    const float* q1 = s1.ptr != nullptr ? s1.ptr + row0 * s1.ld : nullptr;
    const int64_t st0 = 16 * s0.ld, st1 = 16 * s1.ld;
    const bool full = m0 + BM <= op.rows;
    if (full) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            t.v[i] = q0 != nullptr ? ldg4(q0 + i * st0) : zero;
            if (MODE == MODE_BF16) t.w[i] = q1 != nullptr ? ldg4(q1 + i * st1) : zero;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const bool ok = row0 + 16 * i < op.rows;
            t.v[i] = (ok && q0 != nullptr) ? ldg4(q0 + i * st0) : zero;
            if (MODE == MODE_BF16) t.w[i] = (ok && q1 != nullptr) ? ldg4(q1 + i * st1) : zero;
        }
    }
    if (SCALED) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int64_t row = row0 + 16 * i;
            const bool ok = row < op.rows;
            t.f0[i] = (ok && s0.scale >= 0) ? __ldg(op.scal + 2 * row + s0.scale) : 1.f;
            if (MODE == MODE_BF16) t.f1[i] = (ok && s1.scale >= 0) ? __ldg(op.scal + 2 * row + s1.scale) : 1.f;
        }
    }
}
template <int MODE, bool SCALED>
__device__ __forceinline__ void store_a_seg(uint8_t* tile, uint32_t plane_bytes, int p, ATile<MODE, SCALED>& t) {
    const int c = p & 7, r0 = p >> 3;
    const uint32_t off0 = (uint32_t)r0 * ROW_BYTES + (uint32_t)((c ^ (r0 & 7)) << 4);   // (r0 + 16 i) & 7 == r0 & 7
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float4 v = t.v[i], w = t.w[MODE == MODE_BF16 ? i : 0];
        if (SCALED) {
            const float f = t.f0[i];
            v = make_float4(v.x * f, v.y * f, v.z * f, v.w * f);
            if (MODE == MODE_BF16) {
                const float g1 = t.f1[MODE == MODE_BF16 ? i : 0];
                w = make_float4(w.x * g1, w.y * g1, w.z * g1, w.w * g1);
            }
        }
        const uint32_t off = off0 + (uint32_t)i * (16 * ROW_BYTES);
        if (MODE == MODE_TF32X3) put_chunk_tf32<MODE>(tile, plane_bytes, off, v);
        else put_chunk_bf16(tile, off, v, w);
    }
}

// MN-major tile from a segmented operand (weight gradients: the reduction index is the node).
// R = MN extent of the tile in {32, 64, 128, 256}, KB K-rows (nodes kbase..).  Layout: [K group][MN group of
// 128 B][KR rows x 128 B], KR = 4 K rows per atom for tf32 (SWIZZLE_128B_BASE32B: 32-byte chunks XOR (k & 3)),
// 8 for bf16 (16-byte chunks XOR (k & 7)).  Thread p owns one 16-byte MN chunk and walks nodes with a constant
// pointer stride: plain coalesced vector copies, no transposition.
template <int MODE, bool SCALED>
__device__ __forceinline__ void produce_mn_seg(const SegOperand& op, int grp, uint8_t* tile, uint32_t plane_bytes, int R,
                                               int64_t mn0, int64_t kbase, int64_t k_end, int p, float4* colacc = nullptr,
                                               int KB = tf32_like(MODE) ? 32 : 64) {
    constexpr int EPC = tf32_like(MODE) ? 4 : 8;
    constexpr int KR = tf32_like(MODE) ? 4 : 8;
    constexpr int KR_SHIFT = tf32_like(MODE) ? 2 : 3;
    constexpr uint32_t ATOM = KR * 128;
    const int cpr = R / EPC;                                     // chunks per K row: power of two, 4..64
    const int ch = p & (cpr - 1), kk0 = p / cpr, step = PROD_GROUP_THREADS / cpr;
    const uint32_t kgs = (uint32_t)(cpr >> 3) * ATOM;            // bytes between K groups
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    const int64_t col = mn0 + (int64_t)ch * EPC;
    const SegRef s0 = seg_resolve(op, col, grp);
    const SegRef s1 = MODE == MODE_BF16 ? seg_resolve(op, col + 4, grp) : SegRef{nullptr, 0, -1};
    const int c8 = ch & 7;
    const uint32_t mg_off = (uint32_t)(ch >> 3) * ATOM;
    const int64_t lim = k_end < op.rows ? k_end : op.rows;
    const bool full = kbase + KB <= lim;
    const float* q0 = s0.ptr != nullptr ? s0.ptr + (kbase + kk0) * s0.ld : nullptr;
    const float* q1 = s1.ptr != nullptr ? s1.ptr + (kbase + kk0) * s1.ld : nullptr;
    const int64_t st0 = (int64_t)step * s0.ld, st1 = (int64_t)step * s1.ld;
    for (int j0 = 0; kk0 + j0 * step < KB; j0 += 8) {
        float4 v[8], w[MODE == MODE_BF16 ? 8 : 1];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int kk = kk0 + (j0 + i) * step;
            const bool ok = kk < KB && (full || kbase + kk < lim);
            v[i] = (ok && q0 != nullptr) ? ldg4(q0 + (j0 + i) * st0) : zero;
            if (MODE == MODE_BF16) w[i] = (ok && q1 != nullptr) ? ldg4(q1 + (j0 + i) * st1) : zero;
        }
        if (SCALED) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int kk = kk0 + (j0 + i) * step;
                const int64_t k = kbase + kk;
                if (kk < KB && k < lim) {
                    if (s0.scale >= 0) {
                        const float f = __ldg(op.scal + 2 * k + s0.scale);
                        v[i] = make_float4(v[i].x * f, v[i].y * f, v[i].z * f, v[i].w * f);
                    }
                    if (MODE == MODE_BF16 && s1.scale >= 0) {
                        const float f = __ldg(op.scal + 2 * k + s1.scale);
                        w[i] = make_float4(w[i].x * f, w[i].y * f, w[i].z * f, w[i].w * f);
                    }
                }
            }
        }
        if (colacc != nullptr) {                                  // running column sums (bias gradient)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                colacc[0] = make_float4(colacc[0].x + v[i].x, colacc[0].y + v[i].y, colacc[0].z + v[i].z, colacc[0].w + v[i].w);
                if (MODE == MODE_BF16)
                    colacc[1] = make_float4(colacc[1].x + w[i].x, colacc[1].y + w[i].y, colacc[1].z + w[i].z,
                                            colacc[1].w + w[i].w);
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int kk = kk0 + (j0 + i) * step;
            if (kk < KB) {
                const int kr = kk & (KR - 1);
                const uint32_t swz = tf32_like(MODE) ? (uint32_t)((((c8 >> 1) ^ kr) << 5) | ((c8 & 1) << 4))
                                                         : (uint32_t)((c8 ^ kr) << 4);
                const uint32_t off = (uint32_t)(kk >> KR_SHIFT) * kgs + mg_off + (uint32_t)kr * 128u + swz;
                if (tf32_like(MODE)) put_chunk_tf32<MODE>(tile, plane_bytes, off, v[i]);
                else put_chunk_bf16(tile, off, v[i], w[MODE == MODE_BF16 ? i : 0]);
            }
        }
    }
}

// TF32X3 weight-gradient producer: both MN-major operands of one k-block in a single load phase.
// A thread owns one 16-byte column chunk of A and one of B (resolved once per work item into an MnPlan) and
// walks node rows kk0, kk0 + step, ...; all NA + NB loads are issued before the first conversion so that
// one k-block costs one memory round trip, and every address / swizzled offset is a precomputed base plus a
// compile-time multiple of a precomputed stride (the producers are instruction-issue bound).
struct MnPlan {
    const float* q;        // (node kk0, this thread's column chunk) or nullptr for padding columns
    const float* sc;       // &scal[2 * kk0 + scale] or nullptr
    int64_t row_ld;        // elements per node row
    int64_t step_ld;       // step * row_ld
    int kk0, step;
    uint32_t off0, off1, dk2;   // smem offset of row i: off[i & 1] + (i >> 1) * dk2
};
__device__ __forceinline__ uint32_t mn_off_tf32(int kk, int c8, uint32_t mg_off, uint32_t kgs) {
    const int kr = kk & 3;
    return (uint32_t)(kk >> 2) * kgs + mg_off + (uint32_t)kr * 128u + (uint32_t)((((c8 >> 1) ^ kr) << 5) | ((c8 & 1) << 4));
}
__device__ __forceinline__ MnPlan mn_plan(const SegOperand& op, int grp, int R, int64_t mn0, int p) {
    const int cpr = R >> 2, ch = p & (cpr - 1);
    const SegRef r = seg_resolve(op, mn0 + (int64_t)ch * 4, grp);
    MnPlan pl;
    pl.kk0 = p / cpr;
    pl.step = PROD_GROUP_THREADS / cpr;                       // 2 (256 columns), 4, 8 or 16 (32 columns)
    pl.row_ld = r.ld;
    pl.step_ld = (int64_t)pl.step * r.ld;
    pl.q = r.ptr != nullptr ? r.ptr + (int64_t)pl.kk0 * r.ld : nullptr;
    pl.sc = (r.ptr != nullptr && r.scale >= 0) ? op.scal + 2 * pl.kk0 + r.scale : nullptr;
    const uint32_t mg_off = (uint32_t)(ch >> 3) * 512u, kgs = (uint32_t)(cpr >> 3) * 512u;
    pl.off0 = mn_off_tf32(pl.kk0, ch & 7, mg_off, kgs);
    pl.off1 = mn_off_tf32(pl.kk0 + pl.step, ch & 7, mg_off, kgs);
    pl.dk2 = mn_off_tf32(pl.kk0 + 2 * pl.step, ch & 7, mg_off, kgs) - pl.off0;
    return pl;
}
// wait_stage() is called once all loads are in flight and before the first shared-memory store (the ring stage may
// still be read by the MMAs of the previous use until then).
template <int MODE, bool SCALED, int NA, int NB_, class WaitFn, bool DO_B = true>
__device__ __forceinline__ void produce_mn_pair_tf32(const MnPlan& pa, const MnPlan& pb, uint8_t* a_tile, uint32_t a_plane,
                                                     uint8_t* b_tile, uint32_t b_plane, int64_t kbase, int64_t lim,
                                                     float4* colacc, WaitFn&& wait_stage) {
    constexpr int NB = DO_B ? NB_ : 0;             // DO_B = false: the B operand arrives by TMA
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 va[NA], vb[NB > 0 ? NB : 1];
    float f[NA];
    const float* qa = pa.q != nullptr ? pa.q + kbase * pa.row_ld : nullptr;
    const float* qb = pb.q != nullptr ? pb.q + kbase * pb.row_ld : nullptr;
    const float* qs = (SCALED && pa.sc != nullptr) ? pa.sc + 2 * kbase : nullptr;
    const int64_t left = lim - kbase;                       // valid node rows from kbase on
    if (left >= (int64_t)NA * pa.step) {                    // whole k-block in range: no per-row predicates
#pragma unroll
        for (int i = 0; i < NA; ++i) {
            va[i] = qa != nullptr ? ldg4(qa + i * pa.step_ld) : zero;
            if (SCALED) f[i] = qs != nullptr ? __ldg(qs + i * 2 * pa.step) : 1.f;
        }
#pragma unroll
        for (int i = 0; i < NB; ++i) vb[i] = qb != nullptr ? ldg4(qb + i * pb.step_ld) : zero;
    } else {
#pragma unroll
        for (int i = 0; i < NA; ++i) {
            const bool ok = pa.kk0 + i * pa.step < left;
            va[i] = (ok && qa != nullptr) ? ldg4(qa + i * pa.step_ld) : zero;
            if (SCALED) f[i] = (ok && qs != nullptr) ? __ldg(qs + i * 2 * pa.step) : 1.f;
        }
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const bool ok = pb.kk0 + i * pb.step < left;
            vb[i] = (ok && qb != nullptr) ? ldg4(qb + i * pb.step_ld) : zero;
        }
    }
    wait_stage();
#pragma unroll
    for (int i = 0; i < NA; ++i) {
        if (SCALED) va[i] = make_float4(va[i].x * f[i], va[i].y * f[i], va[i].z * f[i], va[i].w * f[i]);
        if (colacc != nullptr)
            colacc[0] = make_float4(colacc[0].x + va[i].x, colacc[0].y + va[i].y, colacc[0].z + va[i].z, colacc[0].w + va[i].w);
        const uint32_t off = ((i & 1) ? pa.off1 : pa.off0) + (uint32_t)(i >> 1) * pa.dk2;
        put_chunk_tf32<MODE_TF32X3>(a_tile, a_plane, off, va[i]);
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        const uint32_t off = ((i & 1) ? pb.off1 : pb.off0) + (uint32_t)(i >> 1) * pb.dk2;
        put_chunk_tf32<MODE_TF32X3>(b_tile, b_plane, off, vb[i]);
    }
}

// thread p of the 128-thread producer group -> (row r in [0, 64), feature half h): a warp covers 16 rows, two lanes
// share the 32 contiguous bytes a k-block needs of every gathered row.  The row order inside a quarter warp is
// {0,2,4,6}, {1,3,5,7}, ... so that the eight 16-byte stores of a quarter warp (chunks 2a + h XOR (r & 7)) fall into
// eight different bank groups.
// Measured (tools/fused_probe.py, ncu): the producers are bound by the number of distinct 128-byte lines their gathers
// touch (L1 wavefronts), not by bytes or latency - one lane per row (32 lines per load instead of 16) ran 25 % slower,
// as did re-reading the first message in place of predicated-off slots.
__device__ __forceinline__ void agg_thread_map(int p, int& r, int& h) {
    const int lane = p & 31, q = lane >> 1;
    h = lane & 1;
    r = 16 * (p >> 5) + ((q >> 2) & 1) + 2 * (q & 3) + 8 * (q >> 3);
}
// predicated 128-bit gather: `v` keeps its value when `pred` is false (no branch, no memory request)
__device__ __forceinline__ void ldg4_pred(float4& v, const float* p, bool pred) {
    asm volatile(
        "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\t"
        "@q ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];\n\t}"
        : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w)
        : "l"(p), "r"((int)pred));
}
// Branch-free for the first PF in-edges: slots beyond the segment re-read its first message, which leaves every
// statistic unchanged (q - q0 = 0, min / max idempotent), so the index loads of both rows, then all gathers of both
// rows, are issued back to back - three exposed latencies per k-block instead of three per row and edge.
__device__ __forceinline__ void agg_compute(const AggSrc& ag, int grp, int64_t m0, int64_t rows, int j, int p, float4 (&v)[8]) {
    constexpr int PF = 3;
    int r, h;
    agg_thread_map(p, r, h);
    const int col = grp * ag.F + 8 * j + 4 * h;                     // channel of Psrc / tab
    const float* const pbase = ag.P + col;
    const float* const tbase = ag.tab + col;
    int beg[2], cnt[2];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int64_t node = m0 + r + 64 * half;
        const bool ok = node < rows;
        const int64_t nn = ok ? node : 0;                           // rows > 0: always a valid address
        beg[half] = __ldg(ag.rowptr + nn);
        cnt[half] = ok ? __ldg(ag.rowptr + nn + 1) : 0;
    }
    int so[2][PF], co[2][PF];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        cnt[half] = max(cnt[half] - beg[half], 0);
        if (cnt[half] == 0) beg[half] = 0;                          // empty segment: read slot 0 (valid), result discarded
#pragma unroll
        for (int k = 0; k < PF; ++k) {
            const int slot = beg[half] + (k < cnt[half] ? k : 0);
            so[half][k] = __ldg(ag.col_src + slot);
            co[half][k] = (int)__ldg(ag.col_code + slot);
        }
    }
    float4 tab0 = ldg4(tbase);                                      // bond code 0: every self-loop, plain single bonds
    float4 pr[2][PF], tb[2][PF];
#pragma unroll
    for (int half = 0; half < 2; ++half)
#pragma unroll
        for (int k = 0; k < PF; ++k) {
            const bool on = k < cnt[half] || k == 0;
            pr[half][k] = make_float4(0.f, 0.f, 0.f, 0.f);
            tb[half][k] = tab0;
            ldg4_pred(pr[half][k], pbase + so[half][k] * ag.C, on);
            ldg4_pred(tb[half][k], tbase + co[half][k] * ag.C, on && co[half][k] != 0);
        }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        PnaAcc<false> acc;
        const float4 q0 = pna_add4(pr[half][0], tb[half][0]);
        acc.first(q0);
#pragma unroll
        for (int k = 1; k < PF; ++k) acc.next(k < cnt[half] ? pna_add4(pr[half][k], tb[half][k]) : q0);
        for (int k = PF; k < cnt[half]; ++k)                        // in-degree > PF: serial
            acc.next(pna_add4(ldg4(pbase + __ldg(ag.col_src + beg[half] + k) * ag.C),
                              ldg4(tbase + (int)__ldg(ag.col_code + beg[half] + k) * ag.C)));
        float4 mean, sd;
        acc.finish(max(cnt[half], 1), mean, sd);
        const bool some = cnt[half] > 0;
        const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
        v[4 * half] = some ? mean : zero;
        v[4 * half + 1] = some ? acc.mn : zero;
        v[4 * half + 2] = some ? acc.mx : zero;
        v[4 * half + 3] = some ? sd : zero;
    }
}
// shared-space 128-bit store (the generic ST.E the compiler emits for pointers derived from the aligned dynamic
// shared-memory base costs an address-space check per access)
__device__ __forceinline__ void sts4(uint32_t saddr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(saddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void agg_store(const AggSrc& ag, int grp, int64_t m0, int64_t rows, int j, int p, uint8_t* tile,
                                          uint32_t plane_bytes, const float4 (&v)[8]) {
    int r, h;
    agg_thread_map(p, r, h);
    const uint32_t base = smem_u32(tile);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int row = r + 64 * half;                               // (row & 7) == (r & 7)
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            float4 hi, lo;
            split_tf32(v[4 * half + a], hi, lo);
            const uint32_t off = base + sw128_offset(row, 2 * a + h);
            sts4(off, hi);
            sts4(off + plane_bytes, lo);
        }
        if (ag.A_out != nullptr && m0 + row < rows) {
            float* dst = ag.A_out + (m0 + row) * (4 * (int64_t)ag.C) + (int64_t)grp * 4 * ag.F + 8 * j + 4 * h;
#pragma unroll
            for (int a = 0; a < 4; ++a) st4_stream(dst + a * ag.F, v[4 * half + a]);
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
constexpr int PACK_ROWS = 32;
template <int MODE, class BL>
__global__ void __launch_bounds__(PROD_GROUP_THREADS)
pack_b_kernel(const BL bl0, int BN, int64_t K, int n_tiles, int num_kb, uint8_t* __restrict__ out) {
    constexpr int KB_ELEMS = tf32_like(MODE) ? 32 : 64;
    constexpr int PARTS = tf32_like(MODE) ? 2 : 1;
    // Programmatic dependent launch: the GEMM that consumes this image is launched right behind this kernel with
    // programmatic stream serialization; releasing it here lets its CTAs set up (barriers, TMEM, first activation
    // loads) while the image is being written - only its weight fetches wait (griddepcontrol.wait) for this grid.
    // (This kernel itself is launched the plain way, behind everything earlier, so the GEMM may read its activations
    // at once.  Programmatic launches for EVERY kernel of the step were measured too: 4.82 vs 4.77 ms at 4096 graphs,
    // 1.72 vs 1.62 ms at 512 - the resident waiting CTAs cost more than the overlapped launch latency saves.)
    asm volatile("griddepcontrol.launch_dependents;");
    // blockIdx.x = k-block * row_splits + row split: a CTA writes PACK_ROWS rows (whole 8-row swizzle atoms) of a tile
    const int row_splits = (BN + PACK_ROWS - 1) / PACK_ROWS;
    const int kb = blockIdx.x / row_splits, r_begin = (blockIdx.x % row_splits) * PACK_ROWS;
    const int nt = blockIdx.y, grp = blockIdx.z;
    BL bl = bl0;
    bl.set_group(grp);
    const uint32_t b_plane = (uint32_t)BN * ROW_BYTES;
    uint8_t* tile = out + ((size_t)(grp * n_tiles + nt) * num_kb + kb) * (size_t)(PARTS * b_plane) + (size_t)(r_begin / 8) * 1024;
    const int rows = BN - r_begin < PACK_ROWS ? BN - r_begin : PACK_ROWS;
    produce_tile<MODE>(bl, tile, b_plane, rows, (int64_t)nt * BN + r_begin, (int64_t)kb * KB_ELEMS, K, threadIdx.x);
}

// A_TMA: the activation operand is ONE plain row-major fp32 matrix - its 128 x 32 tiles are fetched by the TMA unit
// through a tensor map (cp.async.bulk.tensor.2d, SWIZZLE_128B, zero fill outside the matrix) straight into the `hi`
// plane: kind::tf32 MMAs read only the top 19 bits of each word, so the raw fp32 tile IS the tf32 hi part (by
// truncation); the producers then only compute lo = x - trunc(x) from shared memory.  No global loads, address
// arithmetic or bounds checks in the producer warps.
template <int MODE, bool PACKED, bool MN_MAJOR, bool SCALED, class BL, class EP, class AG = NoAgg, bool A_TMA = false>
__global__ void __launch_bounds__(cta_threads(MN_MAJOR, !std::is_same<AG, NoAgg>::value), 1)
gemm_tc_kernel(const SegOperand aop, const BL bl0, const EP ep0, const Geom g, const uint8_t* __restrict__ packed_b,
               const AG ag, const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap amap1,
               const int tma_split, const int tma_gstride0, const int tma_gstride1) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[4], empty_bar[4], tmem_full_bar[2], tmem_empty_bar[2], atma_bar[4];
    // MN-major (weight-gradient) launches: A_TMA means the B operand (a plain matrix: the aggregates A_t or x) arrives
    // through the tensor map, as 32-column x KBR-row boxes in SWIZZLE_128B_ATOM_32B - the MN-major tf32 layout; the A
    // operand (degree-scaled gradients) is still produced by the warps.
    static_assert(!A_TMA || (MODE == MODE_TF32X3 && std::is_same<AG, NoAgg>::value &&
                             ((PACKED && !MN_MAJOR && !SCALED) || (MN_MAJOR && !PACKED))),
                  "tensor-map operand tiles: TF32X3 launches, plain K-major activations or plain MN-major B operands");
    __shared__ uint32_t tmem_base_s;

    constexpr int KB_ELEMS = tf32_like(MODE) ? 32 : 64;
    constexpr int PARTS = tf32_like(MODE) ? 2 : 1;
    constexpr bool AGG = !std::is_same<AG, NoAgg>::value;      // aggregate k-blocks produced on the fly (K-major, TF32X3)
    static_assert(!AGG || (MODE == MODE_TF32X3 && !MN_MAJOR), "fused aggregates: K-major TF32X3 launches only");
    constexpr bool MMA_IN_EPI = MN_MAJOR || AGG;               // epilogue warp 0 also issues the MMAs (20-warp CTA)
    constexpr int EPI_WARPS = epi_warps(MN_MAJOR, AGG), MMA_WARP = MMA_IN_EPI ? 0 : EPI_WARPS;
    constexpr int PROD_WARP0 = MMA_IN_EPI ? EPI_WARPS : EPI_WARPS + 1;
    // The warp index goes through a shuffle so that the compiler treats it (and every role branch on it) as
    // warp-uniform: inside a branch it believes divergent it cannot use the uniform datapath, and each tcgen05.mma -
    // whose descriptors live in uniform registers - was wrapped in an ELECT / 4 x R2UR.BROADCAST "waterfall" loop
    // that paced the single issuing thread at ~140 cycles per MMA (tools/mma_noise.cu reproduces the number).
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    const int BN = g.BN, NS = g.stages;
    const long long kernel_t0 = GNN_PROF(g) ? clock64() : 0;

    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int KBR = MN_MAJOR ? g.kb_rows : KB_ELEMS;                 // K rows per ring stage
    const uint32_t kdiv = MN_MAJOR ? (uint32_t)(KB_ELEMS / g.kb_rows) : 1u;
    const uint32_t a_plane = BM * ROW_BYTES / kdiv, b_plane = (uint32_t)BN * ROW_BYTES / kdiv;
    const uint32_t stage_bytes = PARTS * (a_plane + b_plane);        // multiple of 1024

    if (warp == MMA_WARP) tmem_alloc(&tmem_base_s, g.tmem_cols);
    if (tid == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(&full_bar[s], PROD_GROUP_THREADS);
            mbar_init(&empty_bar[s], 1);
            if (A_TMA) mbar_init(&atma_bar[s], 1);
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

    // =============================== epilogue (one work item) ===============================
    // thread = one accumulator row (TMEM lane); warp w reads lane quarter w & 3 and the 16-column chunks
    // c = (w >> 2), (w >> 2) + EPI_WARPS / 4, ...; values go straight from registers to global memory (each lane
    // writes 64 contiguous bytes of its row per chunk)
    auto epi_item = [&](int64_t item, uint32_t t) {
        const int q = warp & 3, half = warp >> 2;
        {
            int64_t m0, k_begin, k_end;
            int n0, grp;
            decode(item, m0, n0, grp, k_begin, k_end);
            EP ep = ep0;
            ep.set_group(grp);
            const uint32_t b = t & 1, tuse = t >> 1;
            const long long et0 = GNN_PROF(g) ? clock64() : 0;
            mbar_wait(&tmem_full_bar[b], tuse & 1);
            tc_fence_after();
            const long long et1 = GNN_PROF(g) ? clock64() : 0;
            const uint32_t taddr = tmem_base + b * (uint32_t)BN + ((uint32_t)(q * 32) << 16);
            const int64_t m = m0 + q * 32 + lane;
            const bool row_ok = m < g.M;
            const int out_cols = EP::kCombine3 ? BN / 3 : BN;          // columns that are written out
            float amp = 0.f, att = 0.f;
            if constexpr (EP::kCombine3) {
                if (row_ok) ep.scalers(m, amp, att);
            }
            long long ld_cyc = 0;
            bool staged = false;
            if constexpr (EP::kStaged && !MN_MAJOR) {
                if (g.epi_stage_off != 0) {
                    // Transposing epilogue for wide plain outputs: with one accumulator row per lane every store
                    // instruction touched 32 different 128-byte lines (32-byte pieces) and the 128 KB of a
                    // 128 x 256 item took ~7000 cycles to leave the SM (the kernel ran 72.8 -> 39.4 us with the
                    // stores removed).  Each warp now stages 32 rows x 32 columns in a swizzled 4 KB tile and writes
                    // them back as full lines: 8 lanes cover the 128 contiguous bytes of a row.
                    staged = true;
                    uint8_t* tile = smem + g.epi_stage_off + (uint32_t)warp * 4096u;
                    for (int col = half * 32; col < out_cols; col += 8 * EPI_WARPS) {
                        float v[32];
                        tmem_ld16(taddr + col, v);
                        tmem_ld16(taddr + col + 16, v + 16);
#pragma unroll
                        for (int c4 = 0; c4 < 8; ++c4)
                            *reinterpret_cast<float4*>(tile + lane * 128 + ((c4 ^ (lane & 7)) << 4)) =
                                make_float4(v[4 * c4], v[4 * c4 + 1], v[4 * c4 + 2], v[4 * c4 + 3]);
                        __syncwarp();
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int r = (lane >> 3) + 4 * i, c4 = lane & 7;
                            const float4 val = *reinterpret_cast<const float4*>(tile + r * 128 + ((c4 ^ (r & 7)) << 4));
                            const int64_t mr = m0 + q * 32 + r;
                            const int n = n0 + col + c4 * 4;
                            if (mr < g.M && n < g.N) *reinterpret_cast<float4*>(ep.y + mr * ep.ldy + n) = val;
                        }
                        __syncwarp();
                    }
                }
            }
            for (int col = half * 16; col < out_cols && !staged; col += 4 * EPI_WARPS) {
                float v[16];
                const long long lt0 = GNN_PROF(g) ? clock64() : 0;
                tmem_ld16(taddr + col, v);
                if (GNN_PROF(g)) ld_cyc += clock64() - lt0;
                if constexpr (EP::kCombine3) {
                    float d1[16], d2[16];
                    tmem_ld16(taddr + out_cols + col, d1);
                    tmem_ld16(taddr + 2 * out_cols + col, d2);
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = v[j] + amp * d1[j] + att * d2[j];
                }
                if (row_ok) {
                    bool done = false;
                    if constexpr (EP::kWide) {
                        if (ep.wide_ok(n0 + col, g.N)) {
                            ep.store16(m, n0 + col, v);
                            done = true;
                        }
                    }
                    if (!done) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            const float4 v4 = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                            if constexpr (EP::kCombine3) ep.emit4(m, (int)(n0 / 3) + col + j, g.N, v4);
                            else if (n0 + col + j < g.N) ep.emit4(m, n0 + col + j, g.N, v4);
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&tmem_empty_bar[b]);
            if (GNN_PROF(g) && tid == 0) {
                atomicAdd(GNN_PROF(g) + 6, (unsigned long long)(et1 - et0));             // epilogue waiting for the accumulator
                atomicAdd(GNN_PROF(g) + 7, (unsigned long long)(clock64() - et1));       // epilogue work
                atomicAdd(GNN_PROF(g) + 8, 1ull);
                atomicAdd(GNN_PROF(g) + 11, (unsigned long long)ld_cyc);                 // ... of which first tcgen05.ld + wait per chunk
                if (t == 0) atomicAdd(GNN_PROF(g) + 14, (unsigned long long)(et1 - kernel_t0));       // first item: accumulator ready
            }
        }
    };
    // =============================== MMA issue (one work item, one warp) ===============================
    const uint32_t idesc = instr_desc(MODE, BN, MN_MAJOR ? 1 : 0);
    // The warp shares its scheduler with six busy producer / epilogue warps, so every instruction of the issue loop
    // counts: running stage counters instead of a division per k-block, descriptors built once per k-block.
    uint32_t ms = 0, muse = 0;                  // ring stage / use count of the next k-block
    auto mma_item = [&](int64_t item, uint32_t t, uint32_t& it) {
        {
            int64_t m0, k_begin, k_end;
            int n0, grp;
            decode(item, m0, n0, grp, k_begin, k_end);
            const int num_kb = (int)((k_end - k_begin + KBR - 1) / KBR);
            const uint32_t b = t & 1, tuse = t >> 1;
            if (tuse > 0) mbar_wait(&tmem_empty_bar[b], (tuse - 1) & 1);     // epilogue drained this accumulator
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + b * (uint32_t)BN;
            for (int kb = 0; kb < num_kb; ++kb, ++it) {
                const uint32_t s = ms, use = muse;
                if (++ms == (uint32_t)NS) {
                    ms = 0;
                    ++muse;
                }
                const long long mt0 = GNN_PROF(g) ? clock64() : 0;
                mbar_wait(&full_bar[s], use & 1);
                if (GNN_PROF(g) && lane == 0) {
                    atomicAdd(GNN_PROF(g) + 4, (unsigned long long)(clock64() - mt0));   // MMA warp waiting for operands
                    atomicAdd(GNN_PROF(g) + 5, 1ull);
                }
                if (elect_one()) {
                    const uint32_t a_hi = smem_u32(smem + (size_t)s * stage_bytes), a_lo = a_hi + a_plane;
                    const uint32_t b_hi = a_hi + PARTS * a_plane, b_lo = b_hi + b_plane;
                    const int ksteps = (CHUNKS / 2) / (int)kdiv;
                    // Descriptors of k-step 0 once per k-block; a k-step only advances the 14-bit start-address field
                    // (16-byte units), so the other k-steps are one 64-bit add each.  This thread's instruction stream
                    // is what paces the tensor pipe: with the descriptors rebuilt per k-step and a predicate set up per
                    // MMA it issued one MMA per ~140 cycles against a hardware rate of 48-64 (tools/mma_rate.cu).
                    uint64_t dah0, dal0, dbh0, dbl0;
                    uint32_t astep, bstep;
                    if (MN_MAJOR) {
                        // one MMA spans two K groups (tf32: 2 x 4 rows, bf16: 2 x 8 rows) linked by the stride offset
                        constexpr uint32_t ATOM = tf32_like(MODE) ? 512 : 1024;
                        constexpr uint32_t LT = tf32_like(MODE) ? 1 : 2;
                        const uint32_t a_kgs = (BM * (tf32_like(MODE) ? 4 : 2) / 128) * ATOM;
                        const uint32_t b_kgs = ((uint32_t)BN * (tf32_like(MODE) ? 4 : 2) / 128) * ATOM;
                        dah0 = smem_desc_sw128_mn(a_hi, ATOM, a_kgs, LT);
                        dal0 = smem_desc_sw128_mn(a_lo, ATOM, a_kgs, LT);
                        if constexpr (A_TMA) {
                            // TMA boxes: [MN group][K row][128 B] - MN groups KBR rows apart, K groups of 4 rows adjacent
                            dbh0 = smem_desc_sw128_mn(b_hi, (uint32_t)KBR * 128u, ATOM, LT);
                            dbl0 = smem_desc_sw128_mn(b_lo, (uint32_t)KBR * 128u, ATOM, LT);
                            bstep = (2 * ATOM) >> 4;
                        } else {
                            dbh0 = smem_desc_sw128_mn(b_hi, ATOM, b_kgs, LT);
                            dbl0 = smem_desc_sw128_mn(b_lo, ATOM, b_kgs, LT);
                            bstep = (2 * b_kgs) >> 4;
                        }
                        astep = (2 * a_kgs) >> 4;
                    } else {
                        dah0 = smem_desc_sw128(a_hi);
                        dal0 = smem_desc_sw128(a_lo);
                        dbh0 = smem_desc_sw128(b_hi);
                        dbl0 = smem_desc_sw128(b_lo);
                        astep = bstep = 32 >> 4;
                    }
#pragma unroll
                    for (int j = 0; j < CHUNKS / 2; ++j) {       // one MMA consumes 32 bytes of K (8 tf32 / 16 bf16)
                        if (j >= ksteps) break;
                        const uint32_t first = (kb == 0 && j == 0) ? 0u : 1u;
                        const uint64_t dah = dah0 + (uint64_t)(j * astep), dal = dal0 + (uint64_t)(j * astep);
                        const uint64_t dbh = dbh0 + (uint64_t)(j * bstep), dbl = dbl0 + (uint64_t)(j * bstep);
                        if (MODE == MODE_TF32X3) {
                            if (j == 0) umma<MODE>(tmem_d, dal, dbh, idesc, first);
                            else umma_acc<MODE>(tmem_d, dal, dbh, idesc);
                            umma_acc<MODE>(tmem_d, dah, dbl, idesc);
                            umma_acc<MODE>(tmem_d, dah, dbh, idesc);
                        } else {
                            if (j == 0) umma<MODE>(tmem_d, dah, dbh, idesc, first);
                            else umma_acc<MODE>(tmem_d, dah, dbh, idesc);
                        }
                    }
                    umma_commit(&empty_bar[s]);                       // smem stage reusable once these MMAs retire
                    if (kb == num_kb - 1) umma_commit(&tmem_full_bar[b]);
                }
                __syncwarp();
            }
            if (GNN_PROF(g) && lane == 0 && t == 0) atomicAdd(GNN_PROF(g) + 13, (unsigned long long)(clock64() - kernel_t0));   // first item: all MMAs issued
        }
    };

    if constexpr (MN_MAJOR) {
        // weight gradients: warp 0 issues the MMAs of an item and then joins warps 1-3 in its epilogue
        if (warp < EPI_WARPS) {
            uint32_t it = 0, t = 0;
            for (int64_t item = blockIdx.x; item < g.num_items; item += gridDim.x, ++t) {
                if (warp == 0) mma_item(item, t, it);
                epi_item(item, t);
            }
        }
    } else if constexpr (AGG) {
        // warp 0 issues the MMAs of item i+1 (the other accumulator) before it joins the epilogue of item i, so the
        // tensor pipe keeps running while the four warps drain an accumulator
        if (warp < EPI_WARPS) {
            uint32_t it = 0, t = 0;
            if (warp == 0 && (int64_t)blockIdx.x < g.num_items) mma_item(blockIdx.x, 0, it);
            for (int64_t item = blockIdx.x; item < g.num_items; item += gridDim.x, ++t) {
                if (warp == 0 && item + gridDim.x < g.num_items) mma_item(item + gridDim.x, t + 1, it);
                epi_item(item, t);
            }
        }
    } else {
        if (warp < EPI_WARPS) {
            uint32_t t = 0;
            for (int64_t item = blockIdx.x; item < g.num_items; item += gridDim.x, ++t) epi_item(item, t);
        } else if (warp == MMA_WARP) {
            uint32_t it = 0, t = 0;
            for (int64_t item = blockIdx.x; item < g.num_items; item += gridDim.x, ++t) mma_item(item, t, it);
        }
    }
    if (warp >= PROD_WARP0) {
        // =============================== producers ===============================
        // NS (<= PROD_GROUPS) groups are active and group pg owns ring stage pg: a group then only ever waits
        // for the MMAs of its own previous k-block, so it can never run two mbarrier phases ahead (a parity
        // wait cannot tell phase n from phase n-2).  Groups pg >= NS idle.
        const int pg = (warp - PROD_WARP0) >> 2;                 // producer group 0..3
        const int p = tid - PROD_WARP0 * 32 - pg * PROD_GROUP_THREADS;
        const int total_kb = (int)((g.K + KB_ELEMS - 1) / KB_ELEMS);      // packed images exist only without split-K
        uint32_t it_mod = 0, my_use = 0;          // running k-block index modulo NS; k-blocks this group has produced
        bool pack_ready = false;                  // PACKED: the pack kernel (programmatic primary) has completed
        for (int64_t item = blockIdx.x; item < g.num_items; item += gridDim.x) {
            int64_t m0, k_begin, k_end;
            int n0, grp;
            decode(item, m0, n0, grp, k_begin, k_end);
            BL bl = bl0;
            if constexpr (!MN_MAJOR) bl.set_group(grp);
            const int num_kb = (int)((k_end - k_begin + KBR - 1) / KBR);
            float4 colacc[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
            const bool do_colsum = MN_MAJOR && g.colsum.base != nullptr && n0 == 0;
            MnPlan pa{}, pb{};
            if constexpr (MN_MAJOR && tf32_like(MODE)) {
                pa = mn_plan(aop, grp, BM, m0, p);
                pb = mn_plan(bl, grp, BN, (int64_t)n0, p);
            }
            // group pg owns ring stage pg: it produces the k-blocks whose running index is pg modulo NS
            int kb_first = pg - (int)it_mod;
            if (kb_first < 0) kb_first += NS;
            if (pg >= NS) kb_first = num_kb;                       // fewer stages than groups: the extra groups idle
            it_mod = (it_mod + (uint32_t)num_kb) % (uint32_t)NS;
            for (int kb = kb_first; kb < num_kb; kb += NS, ++my_use) {
                const uint32_t s = (uint32_t)pg, use = my_use;
                const long long pt0 = GNN_PROF(g) ? clock64() : 0;
                if (GNN_PROF(g) && p == 0 && pg == 0 && my_use == 0) atomicAdd(GNN_PROF(g) + 12, (unsigned long long)(pt0 - kernel_t0));   // set-up until the first k-block
                const int64_t kbase = k_begin + (int64_t)kb * KBR;
                // K-major activation tiles: the global loads go out before the wait for the ring stage
                ATile<MODE, SCALED> at;
                bool agg_kb = false;           // this k-block holds aggregates computed here (warp-uniform)
                if constexpr (AGG) agg_kb = kbase >= ag.xpad;
                if constexpr (AGG) {
                    if (agg_kb) agg_compute(ag, grp, m0, g.M, (int)((kbase - ag.xpad) >> 5), p, at.v);
                    else load_a_seg<MODE, SCALED>(aop, grp, m0, kbase, k_end, p, at);
                } else if constexpr (!MN_MAJOR && !A_TMA) {
                    load_a_seg<MODE, SCALED>(aop, grp, m0, kbase, k_end, p, at);
                }
                if (GNN_PROF(g) && p == 0 && agg_kb) atomicAdd(GNN_PROF(g) + 15, (unsigned long long)(clock64() - pt0));   // gather + reduce
                long long pt1 = 0;
                auto wait_stage = [&]() {
                    if (use > 0) mbar_wait(&empty_bar[s], (use - 1) & 1);
                    pt1 = GNN_PROF(g) ? clock64() : 0;
                };
                // the fp32 MN-major producers call wait_stage themselves, between their load and store phases
                if constexpr (!(MN_MAJOR && tf32_like(MODE) && !PACKED)) wait_stage();
                uint8_t* stage = smem + (size_t)s * stage_bytes;
                if constexpr (PACKED) {
                    // the weight operand was pre-packed as the smem image: one TMA bulk copy per k-block
                    if (p == 0) {
                        if (!pack_ready) {          // no-op unless launched as a programmatic dependent
                            asm volatile("griddepcontrol.wait;" ::: "memory");
                            pack_ready = true;
                        }
                        const uint32_t bytes = PARTS * b_plane;
                        const size_t tile_idx = ((size_t)(grp * g.n_tiles + n0 / BN) * total_kb + kb);
                        mbar_expect_tx(&full_bar[s], bytes);
                        bulk_g2s(stage + PARTS * a_plane, packed_b + tile_idx * bytes, bytes, &full_bar[s]);
                    }
                    if constexpr (A_TMA) {
                        if (p == 0) {
                            // operand = [segment 0 | segment 1] (two plain matrices, e.g. [x | A_t]): columns below
                            // tma_split come from the first tensor map, the rest from the second; a group (tower)
                            // shifts the columns of a segment by its group stride
                            mbar_expect_tx(&atma_bar[s], a_plane);
                            if ((int)kbase < tma_split)
                                tma_load_2d(stage, &amap, (int)kbase + grp * tma_gstride0, (int)m0, &atma_bar[s]);
                            else
                                tma_load_2d(stage, &amap1, (int)kbase - tma_split + grp * tma_gstride1, (int)m0, &atma_bar[s]);
                            mbar_arrive(&atma_bar[s]);
                        }
                        mbar_wait(&atma_bar[s], use & 1);                       // the raw fp32 tile (= hi plane) has landed
                        const int c = p & 7, r0 = p >> 3;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const uint32_t off = sw128_offset(r0 + 16 * i, c);
                            const float4 x = *reinterpret_cast<const float4*>(stage + off);
                            const float4 lo = make_float4(x.x - __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u),
                                                          x.y - __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u),
                                                          x.z - __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u),
                                                          x.w - __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u));
                            *reinterpret_cast<float4*>(stage + a_plane + off) = lo;
                        }
                    } else if constexpr (AGG) {
                        if (agg_kb) agg_store(ag, grp, m0, g.M, (int)((kbase - ag.xpad) >> 5), p, stage, a_plane, at.v);
                        else store_a_seg<MODE, SCALED>(stage, a_plane, p, at);
                    } else {
                        store_a_seg<MODE, SCALED>(stage, a_plane, p, at);
                    }
                } else if constexpr (MN_MAJOR && tf32_like(MODE)) {
                    // loads per thread: A (128 columns) KBR / 4 rows; B (BN columns) KBR * BN / 512 rows
                    const int64_t lim = k_end < aop.rows ? k_end : aop.rows;
                    uint8_t* bt = stage + PARTS * a_plane;
                    float4* ca = do_colsum ? colacc : nullptr;
                    if constexpr (A_TMA) {
                        // B by TMA: one 32-column x KBR-row box per MN group, issued as soon as the stage is free
                        // (while this group converts and stores its A rows); then lo = x - trunc(x) over the tile
                        auto wait_and_issue = [&]() {
                            wait_stage();
                            if (p == 0) {
                                mbar_expect_tx(&atma_bar[s], b_plane);
                                for (int mg = 0; mg < BN / 32; ++mg)
                                    tma_load_2d(bt + (uint32_t)mg * (uint32_t)KBR * 128u, &amap,
                                                n0 + mg * 32 + grp * tma_gstride0, (int)kbase, &atma_bar[s]);
                                mbar_arrive(&atma_bar[s]);
                            }
                        };
                        if (BN >= 128) produce_mn_pair_tf32<MODE, SCALED, 4, 1, decltype(wait_and_issue)&, false>(pa, pb, stage, a_plane, bt, b_plane, kbase, lim, ca, wait_and_issue);
                        else produce_mn_pair_tf32<MODE, SCALED, 8, 1, decltype(wait_and_issue)&, false>(pa, pb, stage, a_plane, bt, b_plane, kbase, lim, ca, wait_and_issue);
                        mbar_wait(&atma_bar[s], use & 1);
                        for (uint32_t off = (uint32_t)p * 16u; off < b_plane; off += PROD_GROUP_THREADS * 16u) {
                            const float4 x = *reinterpret_cast<const float4*>(bt + off);
                            *reinterpret_cast<float4*>(bt + b_plane + off) =
                                make_float4(x.x - __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u),
                                            x.y - __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u),
                                            x.z - __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u),
                                            x.w - __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u));
                        }
                    } else if (BN == 256) produce_mn_pair_tf32<MODE, SCALED, 4, 8>(pa, pb, stage, a_plane, bt, b_plane, kbase, lim, ca, wait_stage);
                    else if (BN == 128) produce_mn_pair_tf32<MODE, SCALED, 4, 4>(pa, pb, stage, a_plane, bt, b_plane, kbase, lim, ca, wait_stage);
                    else if (BN == 64) produce_mn_pair_tf32<MODE, SCALED, 8, 4>(pa, pb, stage, a_plane, bt, b_plane, kbase, lim, ca, wait_stage);
                    else produce_mn_pair_tf32<MODE, SCALED, 8, 2>(pa, pb, stage, a_plane, bt, b_plane, kbase, lim, ca, wait_stage);
                } else if constexpr (MN_MAJOR) {
                    produce_mn_seg<MODE, SCALED>(aop, grp, stage, a_plane, BM, m0, kbase, k_end, p,
                                                 do_colsum ? colacc : nullptr, KBR);
                    produce_mn_seg<MODE, false>(bl, grp, stage + PARTS * a_plane, b_plane, BN, (int64_t)n0, kbase, k_end, p,
                                                nullptr, KBR);
                } else {
                    if constexpr (AGG) {
                        if (agg_kb) agg_store(ag, grp, m0, g.M, (int)((kbase - ag.xpad) >> 5), p, stage, a_plane, at.v);
                        else store_a_seg<MODE, SCALED>(stage, a_plane, p, at);
                    } else {
                        store_a_seg<MODE, SCALED>(stage, a_plane, p, at);
                    }
                    produce_tile<MODE>(bl, stage + PARTS * a_plane, b_plane, BN, (int64_t)n0, kbase, k_end, p);
                }
                const long long pt2 = GNN_PROF(g) ? clock64() : 0;
                fence_async_smem();                                 // generic-proxy writes -> async (tensor core) proxy
                mbar_arrive(&full_bar[s]);
                if (GNN_PROF(g) && p == 0) {
                    atomicAdd(GNN_PROF(g) + 0, (unsigned long long)(pt1 - pt0));     // wait for a free stage
                    atomicAdd(GNN_PROF(g) + 1, (unsigned long long)(pt2 - pt1));     // load + convert + store
                    atomicAdd(GNN_PROF(g) + 2, (unsigned long long)(clock64() - pt2));   // fence + arrive
                    atomicAdd(GNN_PROF(g) + 3, 1ull);
                }
            }
            if constexpr (MN_MAJOR) {
                if (do_colsum) {
                    // bias gradient: reduce the per-thread column sums of this item across the 512 producer
                    // threads in shared memory, then one global atomic per column (the target is a handful of
                    // addresses shared by every CTA: per-thread atomics would serialise in one L2 slice)
                    constexpr int EPC = tf32_like(MODE) ? 4 : 8;
                    constexpr int CPR = BM / EPC;
                    __shared__ float cs_stage[PROD_GROUPS * PROD_GROUP_THREADS * EPC];
                    const int pp = pg * PROD_GROUP_THREADS + p;
                    const float e[8] = {colacc[0].x, colacc[0].y, colacc[0].z, colacc[0].w,
                                        colacc[1].x, colacc[1].y, colacc[1].z, colacc[1].w};
#pragma unroll
                    for (int j = 0; j < EPC; ++j) cs_stage[pp * EPC + j] = e[j];
                    asm volatile("bar.sync 1, %0;" ::"n"(PROD_GROUPS * PROD_GROUP_THREADS) : "memory");
                    if (pp < BM) {
                        const int col = (int)m0 + pp;
                        float t = 0.f;
#pragma unroll
                        for (int gq = 0; gq < PROD_GROUPS; ++gq)
#pragma unroll
                            for (int k0 = 0; k0 < PROD_GROUP_THREADS / CPR; ++k0)
                                t += cs_stage[(gq * PROD_GROUP_THREADS + k0 * CPR + pp / EPC) * EPC + pp % EPC];
                        if (col < g.colsum.ncols && t != 0.f)
                            atomicAdd(g.colsum.base + (int64_t)grp * g.colsum.group_stride +
                                          (int64_t)(col / g.colsum.tower_cols) * g.colsum.tower_stride +
                                          col % g.colsum.tower_cols, t);
                    }
                    asm volatile("bar.sync 1, %0;" ::"n"(PROD_GROUPS * PROD_GROUP_THREADS) : "memory");
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc(tmem_base, g.tmem_cols);
    if (GNN_PROF(g) && tid == 0) {
        atomicAdd(GNN_PROF(g) + 9, (unsigned long long)(clock64() - kernel_t0));
        atomicAdd(GNN_PROF(g) + 10, 1ull);
    }
}

// ------------------------------------------------------------------ host side
static const size_t kSmemLimit = 232448;     // 227 KB opt-in maximum per CTA
#ifdef GNN_B200_PROFILE
static unsigned long long* g_prof = nullptr; // debug build only: per-role cycle counters of the next launches
#endif

// Tensor maps of the activation operand for 32-column x 128-row boxes in SWIZZLE_128B: one plain row-major fp32 matrix,
// or the concatenation of two (e.g. [x | A_t] of the post GEMM; a group = tower shifts the columns of a segment).
// cuTensorMapEncodeTiled comes from the driver through cudaGetDriverEntryPoint (the library links libcudart only).
// Returns false when the operand does not qualify or the encoder is unavailable.
struct ATma {
    CUtensorMap map0, map1;
    int split, gstride0, gstride1;
};
static bool make_a_tensor_maps(const SegOperand& aop, int groups, int64_t K, ATma* out, int box_rows = BM,
                               bool atom32 = false) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            encode = reinterpret_cast<EncodeFn>(fn);
        (void)cudaGetLastError();
    }
    if (encode == nullptr || aop.nseg < 1 || aop.nseg > 2 || aop.rows < 1 || aop.rows >= (1ll << 31) || K >= (1ll << 31))
        return false;
    // every segment: an unscaled plain matrix whose columns start at a k-block boundary; a group shifts its columns
    int64_t covered = 0;
    for (int i = 0; i < aop.nseg; ++i) {
        const Seg& sg = aop.seg[i];
        if (sg.scale >= 0 || sg.start != covered || (sg.start % 32) != 0) return false;
        if (i + 1 < aop.nseg && (sg.width % 32) != 0) return false;
        if ((reinterpret_cast<uintptr_t>(sg.base) & 15) != 0 || (sg.ld % 4) != 0 || sg.ld * 4 >= (1ll << 40)) return false;
        if (groups > 1 && sg.group_stride != 0 && (sg.group_stride % 4) != 0) return false;
        if (sg.group_stride >= (1ll << 30)) return false;
        covered += sg.width;
    }
    if (covered < K) return false;
    for (int i = 0; i < aop.nseg; ++i) {
        const Seg& sg = aop.seg[i];
        // columns of the underlying matrix that any group may address through this segment
        const int64_t cols = sg.width + (int64_t)(groups - 1) * sg.group_stride;
        const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)aop.rows};
        const cuuint64_t strides[1] = {(cuuint64_t)sg.ld * sizeof(float)};
        const cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
        const cuuint32_t estr[2] = {1u, 1u};
        if (encode(i == 0 ? &out->map0 : &out->map1, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(sg.base), dims,
                   strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return false;
    }
    out->split = aop.nseg > 1 ? aop.seg[1].start : (int)(1 << 30);
    out->gstride0 = (int)aop.seg[0].group_stride;
    out->gstride1 = aop.nseg > 1 ? (int)aop.seg[1].group_stride : 0;
    return true;
}

// BL: K-major weight functor (packed through pack_ws when possible) or, for MN_MAJOR, a SegOperand
// TRY_TMA: the A operand may be one plain matrix - fetch its tiles through a tensor map when it qualifies
template <int MODE, bool MN_MAJOR = false, bool SCALED = false, bool TRY_TMA = false, class BL, class EP, class AG = NoAgg>
static int launch(const SegOperand& aop, BL bl, EP ep, int64_t M, int N, int BN, int64_t K, int64_t k_chunk, int splits,
                  int groups, cudaStream_t st, void* pack_ws = nullptr, int64_t pack_ws_bytes = 0,
                  ColSum colsum = ColSum{nullptr, 0, 0, 0, 1}, AG ag = AG{}) {
    constexpr bool AGG = !std::is_same<AG, NoAgg>::value;
    // M == 0 with a pack workspace: only the weight image is packed (a later call passes the same workspace with a
    // negative size); the host uses it to take the pack kernels off the critical path of a captured training step
    const bool pack_only = M == 0 && !MN_MAJOR && pack_ws != nullptr && pack_ws_bytes > 0;
    if ((M <= 0 && !pack_only) || N <= 0 || K <= 0) return 0;
    if (BN % 16 != 0 || BN < 16 || BN > 256) return GNN_ERR_UNSUPPORTED;
    constexpr int parts = tf32_like(MODE) ? 2 : 1;
    constexpr int kb_elems = tf32_like(MODE) ? 32 : 64;
    // wide MN-major TF32X3 tiles use half-height k-blocks: 4 ring stages (one per producer group) still fit
    const int kb_rows = (MN_MAJOR && tf32_like(MODE) && BN >= 128) ? kb_elems / 2 : kb_elems;
    const size_t stage_bytes = (size_t)parts * (BM + BN) * ROW_BYTES * kb_rows / kb_elems;
    const size_t fixed = 1024 + 256 + (MN_MAJOR ? PROD_GROUPS * PROD_GROUP_THREADS * (tf32_like(MODE) ? 16 : 32) : 0);
    int stages = (int)((kSmemLimit - fixed) / stage_bytes);
    if (stages > PROD_GROUPS) stages = PROD_GROUPS;      // one producer group per ring stage
    if (stages < 2) return GNN_ERR_UNSUPPORTED;
    Geom g;
    g.M = M; g.K = K; g.k_chunk = k_chunk; g.N = N; g.BN = BN; g.splits = splits; g.groups = groups;
    g.m_tiles = (int)gnn_ceil_div(M, BM);
    g.n_tiles = EP::kCombine3 ? N / BN : (int)gnn_ceil_div(N, BN);
    g.num_items = (int64_t)g.m_tiles * g.n_tiles * groups * splits;
    g.stages = stages;
    g.kb_rows = kb_rows;
    uint32_t cols = 32;
    while ((int)cols < 2 * BN) cols <<= 1;
    g.tmem_cols = cols;
    g.colsum = colsum;
#ifdef GNN_B200_PROFILE
    g.prof = g_prof;
#else
    g.prof = nullptr;
#endif
    g.epi_stage_off = 0;
    size_t smem = (size_t)stages * stage_bytes + fixed;
    if constexpr (EP::kStaged && !MN_MAJOR) {
        // wide plain outputs (BN = 256): 8 epilogue warps x 4 KB transposing tiles behind the operand ring
        if (BN == 256 && ep.staged_ok(N) && smem + 8 * 4096 <= kSmemLimit) {
            g.epi_stage_off = (uint32_t)((size_t)stages * stage_bytes);
            smem += 8 * 4096;
        }
    }
    int64_t grid = g.num_items < gnn_sm_count() ? g.num_items : gnn_sm_count();

    if constexpr (MN_MAJOR) {
        if ((BN & (BN - 1)) != 0 || BN < (MODE == MODE_TF32X3 ? 32 : 64)) return GNN_ERR_UNSUPPORTED;   // whole 128-byte groups
        if constexpr (AGG) return GNN_ERR_UNSUPPORTED;
        using KernT = void (*)(const SegOperand, const BL, const EP, const Geom, const uint8_t*, const NoAgg, const CUtensorMap,
                               const CUtensorMap, const int, const int, const int);
        KernT kern = gemm_tc_kernel<MODE, false, true, SCALED, BL, EP>;
        ATma tm{};
        if constexpr (TRY_TMA && MODE == MODE_TF32X3 && std::is_same<BL, SegOperand>::value) {
            // B operand (a plain matrix, the MN dimension tiled by BN columns) through a tensor map
            if (stages >= 3 && N % 32 == 0 && BN % 32 == 0 && make_a_tensor_maps(bl, groups, N, &tm, kb_rows, true) &&
                bl.nseg == 1)
                kern = gemm_tc_kernel<MODE, false, true, SCALED, BL, EP, NoAgg, true>;
        }
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        kern<<<(unsigned)grid, cta_threads(true), smem, st>>>(aop, bl, ep, g, nullptr, NoAgg{}, tm.map0, tm.map1, 0, tm.gstride0,
                                                              0);
    } else {
        // weights (B) are packed once per call when a workspace is given and the M loop re-uses them
        const int num_kb = (int)gnn_ceil_div(K, kb_elems);
        const int64_t pack_bytes = (int64_t)groups * g.n_tiles * num_kb * parts * BN * ROW_BYTES;
        // pack_ws_bytes < 0: the image of exactly these weights was packed into pack_ws by an earlier call (inference
        // with constant weights): no pack kernel
        const bool prepacked = pack_ws_bytes < 0;
        if (prepacked) pack_ws_bytes = -pack_ws_bytes;
        const bool fits = pack_ws != nullptr && splits == 1 && pack_bytes <= pack_ws_bytes &&
                          (reinterpret_cast<uintptr_t>(pack_ws) & 127) == 0;
        const bool packed = fits;
        if ((prepacked || pack_only) && !fits) return GNN_ERR_WORKSPACE;
        ATma tm{};
        bool a_tma = false;
        // (with only two ring stages - the wide model's tiles - the exposed TMA latency costs more than the producers
        // save: measured 12.4 vs 11.8 ms per wide training step)
        if constexpr (TRY_TMA && MODE == MODE_TF32X3 && !SCALED && !AGG)
            a_tma = packed && !pack_only && stages >= 3 && make_a_tensor_maps(aop, groups, K, &tm);
        using KernT = void (*)(const SegOperand, const BL, const EP, const Geom, const uint8_t*, const AG, const CUtensorMap,
                               const CUtensorMap, const int, const int, const int);
        KernT kern = gemm_tc_kernel<MODE, true, false, SCALED, BL, EP, AG>;
        if constexpr (TRY_TMA && MODE == MODE_TF32X3 && !SCALED && !AGG)
            if (a_tma) kern = gemm_tc_kernel<MODE, true, false, SCALED, BL, EP, AG, true>;
        if (packed && prepacked) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
            kern<<<(unsigned)grid, cta_threads(false, AGG), smem, st>>>(aop, bl, ep, g,
                                                                        reinterpret_cast<const uint8_t*>(pack_ws), ag, tm.map0,
                                                                        tm.map1, tm.split, tm.gstride0, tm.gstride1);
        } else if (packed) {
            dim3 pgrid((unsigned)(num_kb * ((BN + PACK_ROWS - 1) / PACK_ROWS)), (unsigned)g.n_tiles, (unsigned)groups);
            pack_b_kernel<MODE, BL><<<pgrid, PROD_GROUP_THREADS, 0, st>>>(bl, BN, K, g.n_tiles, num_kb,
                                                                          reinterpret_cast<uint8_t*>(pack_ws));
            if (pack_only) return gnn_launch_status();
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
            // programmatic dependent of the pack kernel above (see pack_b_kernel)
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)grid);
            cfg.blockDim = dim3((unsigned)cta_threads(false, AGG));
            cfg.dynamicSmemBytes = smem;
            cfg.stream = st;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            e = cudaLaunchKernelEx(&cfg, kern, aop, bl, ep, g, reinterpret_cast<const uint8_t*>(pack_ws), ag, tm.map0, tm.map1,
                                   tm.split, tm.gstride0, tm.gstride1);
            if (e != cudaSuccess) return (int)e;
        } else {
            auto kern_u = gemm_tc_kernel<MODE, false, false, SCALED, BL, EP, AG>;
            cudaError_t e = cudaFuncSetAttribute(kern_u, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
            kern_u<<<(unsigned)grid, cta_threads(false, AGG), smem, st>>>(aop, bl, ep, g, nullptr, ag, tm.map0, tm.map1, 0, 0, 0);
        }
    }
    return gnn_launch_status();
}

static int pick_bn(int n) {
    int bn = (n + 15) / 16 * 16;
    return bn > 256 ? 256 : bn;
}
static int pick_bn_mn(int n, int mode) {          // MN-major tiles: (32 tf32 only) / 64 / 128 / 256 columns
    int bn = mode == MODE_TF32X3 ? 32 : 64;       // 16-bit MN-major tiles need whole 128-byte rows (64 columns)
    while (bn < n && bn < 256) bn <<= 1;
    return bn;
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

// one plain row-major matrix as a segmented operand
static SegOperand seg_plain(const float* p, int64_t ld, int64_t rows, int cols) {
    SegOperand op{};
    op.seg[0] = Seg{p, ld, 0, 0, cols, -1};
    op.nseg = 1;
    op.rows = rows;
    op.scal = nullptr;
    return op;
}

}  // namespace tc

using namespace tc;

#define DISPATCH_MODE(precision, CALL)                         \
    do {                                                       \
        if ((precision) == 0) { constexpr int MODE = MODE_TF32X3; CALL; } \
        else if ((precision) == 2) { constexpr int MODE = MODE_BF16; CALL; } \
        else return GNN_ERR_BAD_ARG;                           \
    } while (0)

// The vectorised producers need 16-byte aligned rows: leading dimensions and column offsets % 4 == 0.
#define REQUIRE_VEC(p, ld) do { if (!is_vec((p), (ld))) return GNN_ERR_UNSUPPORTED; } while (0)

extern "C" int gnn_tc_linear_fwd(const float* X, int64_t ldx, const float* W, int64_t ldw, const float* bias, int64_t M,
                                 int N, int K, int act, float* Y, int64_t ldy, int precision, void* pack_ws,
                                 int64_t pack_ws_bytes, void* stream) {
    if (M < 0 || N < 1 || K < 1) return GNN_ERR_BAD_ARG;
    REQUIRE_VEC(X, ldx);
    if (K % 4 != 0) return GNN_ERR_UNSUPPORTED;
    SegOperand aop = seg_plain(X, ldx, M, K);
    VRowMajor bl{W, ldw, N, K, 0, is_vec(W, ldw)};
    EStore ep{Y, ldy, 0, bias, 0, act, 0, is_vec(Y, ldy) && (bias == nullptr || is_vec(bias, 0))};
    const int BN = pick_bn(N);
    DISPATCH_MODE(precision, return (launch<MODE, false, false, true>(aop, bl, ep, M, N, BN, K, K, 1, 1, gnn_stream(stream), pack_ws,
                                                  pack_ws_bytes)));
}

// Y = act(X W^T + bias) + skip   (skip may be NULL; N % 4 == 0).  With W / bias from gnn_bn_fold_linear this is
// Linear -> BatchNorm(eval) -> ReLU -> skip add of reference models.py:128-131 in one GEMM.
extern "C" int gnn_tc_linear_skip_fwd(const float* X, int64_t ldx, const float* W, int64_t ldw, const float* bias,
                                      int64_t M, int N, int K, int act, const float* skip, int64_t ldskip, float* Y,
                                      int64_t ldy, int precision, void* pack_ws, int64_t pack_ws_bytes, void* stream) {
    if (M < 0 || N < 1 || K < 1 || bias == nullptr) return GNN_ERR_BAD_ARG;
    REQUIRE_VEC(X, ldx);
    REQUIRE_VEC(Y, ldy);
    REQUIRE_VEC(bias, 0);
    if (skip != nullptr) REQUIRE_VEC(skip, ldskip);
    if (K % 4 != 0 || N % 4 != 0) return GNN_ERR_UNSUPPORTED;
    SegOperand aop = seg_plain(X, ldx, M, K);
    VRowMajor bl{W, ldw, N, K, 0, is_vec(W, ldw)};
    EStoreSkip ep{Y, ldy, bias, skip, ldskip, act};
    const int BN = pick_bn(N);
    DISPATCH_MODE(precision, return (launch<MODE, false, false, true>(aop, bl, ep, M, N, BN, K, K, 1, 1, gnn_stream(stream), pack_ws,
                                                  pack_ws_bytes)));
}

extern "C" int gnn_tc_linear_dgrad(const float* dY, int64_t lddy, const float* W, int64_t ldw, int64_t M, int N, int K,
                                   int accumulate, float* dX, int64_t lddx, int precision, void* pack_ws,
                                   int64_t pack_ws_bytes, void* stream) {
    if (M < 0 || N < 1 || K < 1) return GNN_ERR_BAD_ARG;
    REQUIRE_VEC(dY, lddy);
    if (N % 4 != 0) return GNN_ERR_UNSUPPORTED;
    SegOperand aop = seg_plain(dY, lddy, M, N);            // (m, n): reduction over n
    VTransposed bl{W, ldw, K, N, 0};                       // row kk, k = n -> W[n, kk]
    EStore ep{dX, lddx, 0, nullptr, 0, 0, accumulate, is_vec(dX, lddx)};
    const int BN = pick_bn(K);
    DISPATCH_MODE(precision, return (launch<MODE, false, false, true>(aop, bl, ep, M, K, BN, N, N, 1, 1, gnn_stream(stream), pack_ws,
                                                  pack_ws_bytes)));
}

extern "C" int gnn_tc_linear_wgrad(const float* dY, int64_t lddy, const float* X, int64_t ldx, int64_t M, int N, int K,
                                   float* dW, int64_t lddw, float* db, int precision, void* stream) {
    if (M < 0 || N < 1 || K < 1) return GNN_ERR_BAD_ARG;
    REQUIRE_VEC(dY, lddy);
    REQUIRE_VEC(X, ldx);
    if (N % 4 != 0 || K % 4 != 0) return GNN_ERR_UNSUPPORTED;
    SegOperand aop = seg_plain(dY, lddy, M, N);            // (node i, o) -> dY[i, o]
    SegOperand bop = seg_plain(X, ldx, M, K);              // (node i, kk) -> X[i, kk]
    EAtomic ep{dW, lddw, 0, is_vec(dW, lddw)};
    int64_t k_chunk;
    int splits;
    DISPATCH_MODE(precision, {
        const int BN = pick_bn_mn(K, MODE);
        const int64_t tiles = gnn_ceil_div(N, BM) * gnn_ceil_div(K, BN);
        splitk(M, tiles, tf32_like(MODE) ? 32 : 64, k_chunk, splits);
        return (launch<MODE, true, false, true>(aop, bop, ep, N, K, BN, M, k_chunk, splits, 1, gnn_stream(stream), nullptr, 0,
                                   ColSum{db, 0, 0, db ? N : 0, N}));
    });
}

// ------------------------------------------------------------------ PNA post stage on tensor cores
static SegOperand seg_cat5(const float* x, const float* A, int64_t n, int F, int towers) {   // [x | A_t], group = tower
    SegOperand op{};
    op.seg[0] = Seg{x, F, 0, 0, F, -1};
    op.seg[1] = Seg{A, (int64_t)towers * 4 * F, 4 * (int64_t)F, F, 4 * F, -1};
    op.nseg = 2;
    op.rows = n;
    op.scal = nullptr;
    return op;
}
// [dO_t | amp dO_t | att dO_t], group = tower; the three groups start `pitch` columns apart (>= F_out)
static SegOperand seg_scaled_grad(const float* dO, const float* scal, int64_t n, int F_out, int towers, int pitch) {
    SegOperand op{};
    for (int s = 0; s < 3; ++s) op.seg[s] = Seg{dO, (int64_t)towers * F_out, F_out, s * pitch, F_out, s - 1};
    op.nseg = 3;
    op.rows = n;
    op.scal = scal;
    return op;
}

static int post_fwd_impl(const float* x, const float* A, const float* scal, const float* vx, const float* wpost,
                         int64_t wpost_tower_stride, const float* bpost, int64_t bpost_tower_stride,
                         int64_t n, int F, int F_out, int towers, int act, float* O, int precision,
                         void* pack_ws, int64_t pack_ws_bytes, void* stream) {
    if (n < 0 || F < 4 || (F & 3) || F_out < 1 || towers < 1) return GNN_ERR_BAD_ARG;
    REQUIRE_VEC(x, F);
    REQUIRE_VEC(A, 4 * (int64_t)F);
    int FoP = (F_out + 15) / 16 * 16;      // output-feature tile: three scaler groups of FoP columns per CTA
    if (FoP > 64) FoP = 64;
    const int tiles = (F_out + FoP - 1) / FoP;
    SegOperand aop = seg_cat5(x, A, n, F, towers);
    VPostWeight bl{wpost, wpost_tower_stride, F, F_out, FoP, vx, 0};
    EPostCombine ep{O, scal, bpost, bpost_tower_stride, F_out, FoP, towers, 0, act,
                    (F_out % 4 == 0) && is_vec(O, 0) && is_vec(bpost, bpost_tower_stride)};
    DISPATCH_MODE(precision, return (launch<MODE, false, false, true>(aop, bl, ep, n, tiles * 3 * FoP, 3 * FoP, 5 * (int64_t)F,
                                                  5 * (int64_t)F, 1, towers, gnn_stream(stream), pack_ws, pack_ws_bytes)));
}

extern "C" int gnn_tc_pna_post_fwd(const float* x, const float* A, const float* scal, const float* wpost,
                                   int64_t wpost_tower_stride, const float* bpost, int64_t bpost_tower_stride,
                                   int64_t n, int F, int F_out, int towers, int act, float* O, int precision,
                                   void* pack_ws, int64_t pack_ws_bytes, void* stream) {
    return post_fwd_impl(x, A, scal, nullptr, wpost, wpost_tower_stride, bpost, bpost_tower_stride, n, F, F_out, towers,
                         act, O, precision, pack_ws, pack_ws_bytes, stream);
}
// folded destination projection: A holds the aggregates of q, Vx [T,3,F_out,F] the folded x columns (gnn_pna_fold_fwd)
extern "C" int gnn_tc_pna_postf_fwd(const float* x, const float* A, const float* scal, const float* Vx,
                                    const float* wpost, int64_t wpost_tower_stride, const float* bpost,
                                    int64_t bpost_tower_stride, int64_t n, int F, int F_out, int towers, int act,
                                    float* O, int precision, void* pack_ws, int64_t pack_ws_bytes, void* stream) {
    if (Vx == nullptr) return GNN_ERR_BAD_ARG;
    REQUIRE_VEC(Vx, F);
    return post_fwd_impl(x, A, scal, Vx, wpost, wpost_tower_stride, bpost, bpost_tower_stride, n, F, F_out, towers, act,
                         O, precision, pack_ws, pack_ws_bytes, stream);
}

// Folded post forward with the aggregates computed by the GEMM's producers (AggSrc): Psrc / tab / CSR in, O out; the
// aggregates reach HBM only when A_out != NULL (training).  TF32x3 only; F % 8 == 0.
extern "C" int gnn_tc_pna_postf_agg_fwd(const float* x, const float* Psrc, const float* tab, const int32_t* rowptr,
                                        const int32_t* col_src, const uint8_t* col_code, const float* scal,
                                        const float* Vx, const float* wpost, int64_t wpost_tower_stride,
                                        const float* bpost, int64_t bpost_tower_stride, int64_t n, int F, int F_out,
                                        int towers, int act, float* O, float* A_out, int precision, void* pack_ws,
                                        int64_t pack_ws_bytes, void* stream) {
    if (n < 0 || F < 8 || (F & 7) || F_out < 1 || towers < 1 || Vx == nullptr) return GNN_ERR_BAD_ARG;
    if (precision != 0) return GNN_ERR_UNSUPPORTED;
    const int C = towers * F;
    if ((int64_t)n * C >= (1ll << 31)) return GNN_ERR_TOO_LARGE;        // 32-bit row offsets in the gather
    REQUIRE_VEC(x, F);
    REQUIRE_VEC(Psrc, (int64_t)C);
    REQUIRE_VEC(tab, (int64_t)C);
    REQUIRE_VEC(Vx, F);
    if (A_out != nullptr) REQUIRE_VEC(A_out, 4 * (int64_t)F);
    int FoP = (F_out + 15) / 16 * 16;
    if (FoP > 64) FoP = 64;
    const int tiles = (F_out + FoP - 1) / FoP;
    if (tiles > 1 && A_out != nullptr) return GNN_ERR_UNSUPPORTED;      // every n tile would store the aggregates again
    const int xpad = (F + 31) / 32 * 32;
    SegOperand aop = seg_plain(x, F, n, F);
    VPostWeight bl{wpost, wpost_tower_stride, F, F_out, FoP, Vx, xpad};
    EPostCombine ep{O, scal, bpost, bpost_tower_stride, F_out, FoP, towers, 0, act,
                    (F_out % 4 == 0) && is_vec(O, 0) && is_vec(bpost, bpost_tower_stride)};
    AggSrc ag{Psrc, tab, rowptr, col_src, col_code, A_out, F, C, xpad};
    const int64_t K = xpad + 4 * (int64_t)F;
    return (launch<MODE_TF32X3, false, false>(aop, bl, ep, n, tiles * 3 * FoP, 3 * FoP, K, K, 1, towers,
                                               gnn_stream(stream), pack_ws, pack_ws_bytes,
                                               ColSum{nullptr, 0, 0, 0, 1}, ag));
}

extern "C" int gnn_tc_pna_post_dgrad(const float* dO, const float* scal, const float* wpost, int64_t wpost_tower_stride,
                                     int64_t n, int F, int F_out, int towers, float* dx, float* gA, int precision,
                                     void* pack_ws, int64_t pack_ws_bytes, void* stream) {
    if (n < 0 || F < 4 || (F & 3) || F_out < 1 || towers < 1) return GNN_ERR_BAD_ARG;
    if (F_out % 4 != 0) return GNN_ERR_UNSUPPORTED;
    REQUIRE_VEC(dO, (int64_t)towers * F_out);
    cudaStream_t st = gnn_stream(stream);
    // gA[:, t, :] = [dO_t | amp dO_t | att dO_t] . Wslices_t          (M = n, N = 4F, K = 3 F_out)
    {
        SegOperand aop = seg_scaled_grad(dO, scal, n, F_out, towers, F_out);
        VPostWeightT bl{wpost, wpost_tower_stride, F, F_out};
        EStore ep{gA, (int64_t)towers * 4 * F, 4 * (int64_t)F, nullptr, 0, 0, 0, is_vec(gA, 0)};
        const int BN = pick_bn(4 * F);
        int rc;
        DISPATCH_MODE(precision, rc = (launch<MODE, false, true>(aop, bl, ep, n, 4 * F, BN, 3 * (int64_t)F_out, 3 * (int64_t)F_out, 1,
                                                    towers, st, pack_ws, pack_ws_bytes)));
        if (rc) return rc;
    }
    if (dx == nullptr) return 0;       // the caller folds the x part into gnn_tc_pna_dx_dgrad
    // dx += sum_t dO_t . Wpost_t[:, :F]     (towers accumulate into the same dx -> sequential launches)
    for (int t = 0; t < towers; ++t) {
        const float* d = dO + t * F_out;
        const float* w = wpost + t * wpost_tower_stride;
        SegOperand aop = seg_plain(d, (int64_t)towers * F_out, n, F_out);
        VTransposed bl{w, 13 * (int64_t)F, F, F_out, 0};
        EStore ep{dx, F, 0, nullptr, 0, 0, 1, is_vec(dx, F)};
        const int BN = pick_bn(F);
        int rc;
        DISPATCH_MODE(precision, rc = (launch<MODE>(aop, bl, ep, n, F, BN, (int64_t)F_out, (int64_t)F_out, 1, 1, st,
                                                    pack_ws, pack_ws_bytes)));
        if (rc) return rc;
    }
    return 0;
}

// ------------------------------------------------------------------ fused projection / dx GEMMs
// P = [x Wdst_t^T + b_t | x Wsrc_t^T] for both towers in one launch  (replaces 4 Linear calls)
extern "C" int gnn_tc_pna_proj_fwd(const float* x, const float* wpre, int64_t wpre_tower_stride, const float* bpre,
                                   int64_t bpre_tower_stride, int64_t n, int F, int towers, float* P, int precision,
                                   void* pack_ws, int64_t pack_ws_bytes, void* stream) {
    if (n < 0 || F < 4 || (F & 3) || towers < 1) return GNN_ERR_BAD_ARG;
    REQUIRE_VEC(x, F);
    REQUIRE_VEC(P, 2 * (int64_t)towers * F);
    REQUIRE_VEC(wpre, wpre_tower_stride);
    REQUIRE_VEC(bpre, bpre_tower_stride);
    const int C = towers * F;
    SegOperand aop = seg_plain(x, F, n, F);
    VProjWeight bl{wpre, wpre_tower_stride, F, towers, 0, 2};
    EProjStore ep{P, bpre, bpre_tower_stride, F, towers, 2 * (int64_t)C};
    int BN = pick_bn(2 * C);
    if (BN > 128) BN = 128;
    DISPATCH_MODE(precision, return (launch<MODE, false, false, true>(aop, bl, ep, n, 2 * C, BN, (int64_t)F, (int64_t)F, 1, 1,
                                                  gnn_stream(stream), pack_ws, pack_ws_bytes)));
}
// source half only, bias included: Psrc[N, C] = x Wsrc_t^T + b_t   (folded destination projection)
extern "C" int gnn_tc_pna_projsrc_fwd(const float* x, const float* wpre, int64_t wpre_tower_stride, const float* bpre,
                                      int64_t bpre_tower_stride, int64_t n, int F, int towers, float* Psrc,
                                      int precision, void* pack_ws, int64_t pack_ws_bytes, void* stream) {
    if (n < 0 || F < 4 || (F & 3) || towers < 1) return GNN_ERR_BAD_ARG;
    REQUIRE_VEC(x, F);
    REQUIRE_VEC(Psrc, (int64_t)towers * F);
    REQUIRE_VEC(wpre, wpre_tower_stride);
    REQUIRE_VEC(bpre, bpre_tower_stride);
    const int C = towers * F;
    SegOperand aop = seg_plain(x, F, n, F);
    VProjWeight bl{wpre, wpre_tower_stride, F, towers, 1, 1};
    EProjStore ep{Psrc, bpre, bpre_tower_stride, F, towers, (int64_t)C};
    int BN = pick_bn(C);
    if (BN > 128) BN = 128;
    DISPATCH_MODE(precision, return (launch<MODE, false, false, true>(aop, bl, ep, n, C, BN, (int64_t)F, (int64_t)F, 1, 1,
                                                  gnn_stream(stream), pack_ws, pack_ws_bytes)));
}

// dx (+)= dP . Wproj + sum_t dO_t . Wpost_t[:, :F]   (replaces 4 projection dgrads + 2 post x-part dgrads)
extern "C" int gnn_tc_pna_dx_dgrad(const float* dP, const float* dO, const float* wpre, int64_t wpre_tower_stride,
                                   const float* wpost, int64_t wpost_tower_stride, int64_t n, int F, int F_out,
                                   int towers, int accumulate, float* dx, int precision, void* pack_ws,
                                   int64_t pack_ws_bytes, void* stream) {
    if (n < 0 || F < 4 || (F & 3) || F_out < 4 || (F_out & 3) || towers < 1) return GNN_ERR_BAD_ARG;
    const int C = towers * F;
    REQUIRE_VEC(dP, 2 * (int64_t)C);
    REQUIRE_VEC(dO, (int64_t)towers * F_out);
    REQUIRE_VEC(dx, F);
    SegOperand aop{};
    aop.seg[0] = Seg{dP, 2 * (int64_t)C, 0, 0, 2 * C, -1};
    aop.seg[1] = Seg{dO, (int64_t)towers * F_out, 0, 2 * C, towers * F_out, -1};
    aop.nseg = 2;
    aop.rows = n;
    VDxWeight bl{wpre, wpost, wpre_tower_stride, wpost_tower_stride, F, F_out, towers};
    EStore ep{dx, F, 0, nullptr, 0, 0, accumulate, is_vec(dx, F)};
    const int64_t K = 2 * (int64_t)C + (int64_t)towers * F_out;
    const int BN = pick_bn(F);
    DISPATCH_MODE(precision, return (launch<MODE, false, false, true>(aop, bl, ep, n, F, BN, K, K, 1, 1, gnn_stream(stream), pack_ws,
                                                  pack_ws_bytes)));
}

// folded variant: dx (+)= dPsrc . Wsrc + sum_{s,t} scal_s dO_t . Vx_{t,s}
extern "C" int gnn_tc_pna_dxf_dgrad(const float* dPsrc, const float* dO, const float* scal, const float* wpre,
                                    int64_t wpre_tower_stride, const float* Vx, int64_t n, int F, int F_out, int towers,
                                    int accumulate, float* dx, int precision, void* pack_ws, int64_t pack_ws_bytes,
                                    void* stream) {
    if (n < 0 || F < 4 || (F & 3) || F_out < 4 || (F_out & 3) || towers < 1) return GNN_ERR_BAD_ARG;
    const int C = towers * F, TF = towers * F_out;
    REQUIRE_VEC(dPsrc, (int64_t)C);
    REQUIRE_VEC(dO, (int64_t)TF);
    REQUIRE_VEC(dx, F);
    SegOperand aop{};
    aop.seg[0] = Seg{dPsrc, (int64_t)C, 0, 0, C, -1};
    for (int s = 0; s < 3; ++s) aop.seg[1 + s] = Seg{dO, (int64_t)TF, 0, C + s * TF, TF, s - 1};
    aop.nseg = 4;
    aop.rows = n;
    aop.scal = scal;
    VDxWeightFold bl{wpre, Vx, wpre_tower_stride, F, F_out, towers};
    EStore ep{dx, F, 0, nullptr, 0, 0, accumulate, is_vec(dx, F)};
    const int64_t K = (int64_t)C + 3 * (int64_t)TF;
    const int BN = pick_bn(F);
    DISPATCH_MODE(precision, return (launch<MODE, false, true>(aop, bl, ep, n, F, BN, K, K, 1, 1, gnn_stream(stream),
                                                               pack_ws, pack_ws_bytes)));
}

// dWpre_t[:, :2F] += dP^T x, dbpre_t += colsum(dPdst_t)   (replaces 4 Linear wgrads + colsum)
extern "C" int gnn_tc_pna_proj_wgrad(const float* dP, const float* x, int64_t n, int F, int towers, float* dwpre,
                                     int64_t wpre_tower_stride, float* dbpre, int64_t bpre_tower_stride, int precision,
                                     void* stream) {
    if (n < 0 || F < 4 || (F & 3) || towers < 1) return GNN_ERR_BAD_ARG;
    if (n == 0) return 0;
    const int C = towers * F;
    REQUIRE_VEC(dP, 2 * (int64_t)C);
    REQUIRE_VEC(x, F);
    REQUIRE_VEC(dwpre, wpre_tower_stride);
    SegOperand aop = seg_plain(dP, 2 * (int64_t)C, n, 2 * C);
    SegOperand bop = seg_plain(x, F, n, F);
    EProjWgrad ep{dwpre, wpre_tower_stride, F, towers, 0, 2};
    int64_t k_chunk;
    int splits;
    DISPATCH_MODE(precision, {
        const int BN = pick_bn_mn(F, MODE);
        const int64_t tiles = gnn_ceil_div(2 * C, BM) * gnn_ceil_div(F, BN);
        splitk(n, tiles, tf32_like(MODE) ? 32 : 64, k_chunk, splits);
        return (launch<MODE, true, false, true>(aop, bop, ep, 2 * C, F, BN, n, k_chunk, splits, 1, gnn_stream(stream), nullptr, 0,
                                   ColSum{dbpre, 0, bpre_tower_stride, dbpre ? C : 0, F}));
    });
}
// source half only: dWpre_t[:, F:2F] += dPsrc_t^T x, dbpre_t += colsum(dPsrc_t)
extern "C" int gnn_tc_pna_projsrc_wgrad(const float* dPsrc, const float* x, int64_t n, int F, int towers, float* dwpre,
                                        int64_t wpre_tower_stride, float* dbpre, int64_t bpre_tower_stride,
                                        int precision, void* stream) {
    if (n < 0 || F < 4 || (F & 3) || towers < 1) return GNN_ERR_BAD_ARG;
    if (n == 0) return 0;
    const int C = towers * F;
    REQUIRE_VEC(dPsrc, (int64_t)C);
    REQUIRE_VEC(x, F);
    REQUIRE_VEC(dwpre, wpre_tower_stride);
    SegOperand aop = seg_plain(dPsrc, (int64_t)C, n, C);
    SegOperand bop = seg_plain(x, F, n, F);
    EProjWgrad ep{dwpre, wpre_tower_stride, F, towers, 1, 1};
    int64_t k_chunk;
    int splits;
    DISPATCH_MODE(precision, {
        const int BN = pick_bn_mn(F, MODE);
        const int64_t tiles = gnn_ceil_div(C, BM) * gnn_ceil_div(F, BN);
        splitk(n, tiles, tf32_like(MODE) ? 32 : 64, k_chunk, splits);
        return (launch<MODE, true, false, true>(aop, bop, ep, C, F, BN, n, k_chunk, splits, 1, gnn_stream(stream), nullptr, 0,
                                   ColSum{dbpre, 0, bpre_tower_stride, dbpre ? C : 0, F}));
    });
}

static int post_wgrad_impl(const float* dO, const float* x, const float* A, const float* scal, int64_t n,
                           int F, int F_out, int towers, float* dwpost, int64_t wpost_tower_stride,
                           float* dbpost, int64_t bpost_tower_stride, float* dVx, int precision, void* stream) {
    if (n < 0 || F < 4 || (F & 3) || F_out < 1 || towers < 1) return GNN_ERR_BAD_ARG;
    if (n == 0) return 0;
    if (F_out % 4 != 0) return GNN_ERR_UNSUPPORTED;
    REQUIRE_VEC(dO, (int64_t)towers * F_out);
    REQUIRE_VEC(x, F);
    REQUIRE_VEC(A, 4 * (int64_t)F);
    cudaStream_t st = gnn_stream(stream);
    const int FoP = (F_out + 15) / 16 * 16;
    // (1) aggregate part: D[(s,o), r] = sum_i scale_s(i) dO[i,o] A_t[i,r]   (M = 3 FoP, N = 4F, K = n), MN-major
    {
        SegOperand aop = seg_scaled_grad(dO, scal, n, F_out, towers, FoP);
        SegOperand bop{};
        bop.seg[0] = Seg{A, (int64_t)towers * 4 * F, 4 * (int64_t)F, 0, 4 * F, -1};
        bop.nseg = 1;
        bop.rows = n;
        EPostWgrad ep{dwpost, wpost_tower_stride, F, F_out, FoP, is_vec(dwpost, wpost_tower_stride)};
        int64_t k_chunk;
        int splits, rc;
        DISPATCH_MODE(precision, {
            const int BN = pick_bn_mn(4 * F, MODE);
            const int64_t tiles = gnn_ceil_div(4 * F, BN) * gnn_ceil_div(3 * FoP, BM) * towers;
            splitk(n, tiles, tf32_like(MODE) ? 32 : 64, k_chunk, splits);
            rc = (launch<MODE, true, true, true>(aop, bop, ep, 3 * FoP, 4 * F, BN, n, k_chunk, splits, towers, st, nullptr, 0,
                                     ColSum{dbpost, bpost_tower_stride, 0, dbpost ? F_out : 0, F_out}));
        });
        if (rc) return rc;
    }
    // (2') folded destination projection: dVx[t][s][o][:] += sum_i scale_s(i) dO[i,o] x[i,:]  (M = 3 F_out)
    if (dVx != nullptr) {
        SegOperand aop = seg_scaled_grad(dO, scal, n, F_out, towers, F_out);
        SegOperand bop = seg_plain(x, F, n, F);
        EAtomic ep{dVx, (int64_t)F, 3 * (int64_t)F_out * F, is_vec(dVx, F)};
        int64_t k_chunk;
        int splits, rc;
        DISPATCH_MODE(precision, {
            const int BN = pick_bn_mn(F, MODE);
            const int64_t tiles = gnn_ceil_div(F, BN) * gnn_ceil_div(3 * F_out, BM) * towers;
            splitk(n, tiles, tf32_like(MODE) ? 32 : 64, k_chunk, splits);
            rc = (launch<MODE, true, true, true>(aop, bop, ep, 3 * F_out, F, BN, n, k_chunk, splits, towers, st));
        });
        return rc;
    }
    // (2) x part: dWpost_t[o, :F] += dO_t^T x   (both towers in one launch: group = tower)
    {
        SegOperand aop{};
        aop.seg[0] = Seg{dO, (int64_t)towers * F_out, F_out, 0, F_out, -1};
        aop.nseg = 1;
        aop.rows = n;
        SegOperand bop = seg_plain(x, F, n, F);
        EAtomic ep{dwpost, 13 * (int64_t)F, wpost_tower_stride, is_vec(dwpost, F) && (wpost_tower_stride % 4 == 0)};
        int64_t k_chunk;
        int splits, rc;
        DISPATCH_MODE(precision, {
            const int BN = pick_bn_mn(F, MODE);
            const int64_t tiles = gnn_ceil_div(F, BN) * towers;
            splitk(n, tiles, tf32_like(MODE) ? 32 : 64, k_chunk, splits);
            rc = (launch<MODE, true>(aop, bop, ep, F_out, F, BN, n, k_chunk, splits, towers, st));
        });
        if (rc) return rc;
    }
    return 0;
}

#ifdef GNN_B200_PROFILE
// debug build only: 16 x uint64 device buffer that receives per-role cycle counters of subsequent tensor-core GEMM
// launches (NULL disables).  [0] producer wait-for-stage, [1] produce, [2] fence+arrive, [3] #k-blocks (per group
// leader), [4] MMA wait-for-operands, [5] #k-blocks, [6] epilogue wait, [7] epilogue work, [8] #items, [9] CTA cycles,
// [10] #CTAs
extern "C" int gnn_tc_set_profile_buffer(void* buf) {
    tc::g_prof = reinterpret_cast<unsigned long long*>(buf);
    return 0;
}
#endif

extern "C" int gnn_tc_pna_post_wgrad(const float* dO, const float* x, const float* A, const float* scal, int64_t n,
                                     int F, int F_out, int towers, float* dwpost, int64_t wpost_tower_stride,
                                     float* dbpost, int64_t bpost_tower_stride, int precision, void* stream) {
    return post_wgrad_impl(dO, x, A, scal, n, F, F_out, towers, dwpost, wpost_tower_stride, dbpost, bpost_tower_stride,
                           nullptr, precision, stream);
}
// folded destination projection: the aggregate columns go to dwpost as above, the x columns of the three scaler
// groups go to dVx [T,3,F_out,F] (+=, caller zeroes) and are mapped back to parameters by gnn_pna_fold_bwd
extern "C" int gnn_tc_pna_postf_wgrad(const float* dO, const float* x, const float* A, const float* scal, int64_t n,
                                      int F, int F_out, int towers, float* dwpost, int64_t wpost_tower_stride,
                                      float* dbpost, int64_t bpost_tower_stride, float* dVx, int precision,
                                      void* stream) {
    if (dVx == nullptr) return GNN_ERR_BAD_ARG;
    return post_wgrad_impl(dO, x, A, scal, n, F, F_out, towers, dwpost, wpost_tower_stride, dbpost, bpost_tower_stride,
                           dVx, precision, stream);
}

// =====================================================================================================================
// Inference "tile" kernel: the whole PNAConv stack (all layers) of a tile of <= 128 nodes made of WHOLE graphs in one
// persistent CTA.  Replaces, for store-resident inference of the default architecture (H = 64, 2 towers, folded
// destination projection, post_layers == 1, skip connections), the per-layer chain
//   gnn_tc_pna_projsrc_fwd -> gnn_tc_pna_postf_agg_fwd -> gnn_bn_fold_linear + gnn_tc_linear_skip_fwd
// i.e. PNAConv + BatchNorm(running statistics) + ReLU + skip of reference models.py:126-131, for all layers.
// No edge leaves a graph, so a tile never needs another tile's rows: node features x (fp32 [128][64]) and the source
// projections Psrc ([128][128]) live in shared memory for the whole stack and the CSR slice of the tile is staged
// once; per layer only the packed weight images (589 KB, L2-resident) are streamed in by TMA bulk copies.
// Per layer, 26 k-blocks of 32 K-columns run through one two-stage operand ring (SWIZZLE_128B, tf32 hi / lo planes,
// three tcgen05.mma per K step = fp32-level accuracy, accumulators in TMEM):
//   phase A  Psrc_t = x . Wsrc_t^T + b_t                 2 towers x 2 k-blocks, N = 64   -> TMEM -> Psrc (smem)
//   phase B  [x | mean min max std of (Psrc[src] + tab[code])] . [Vx | Wagg]_t^T     2 x 10 k-blocks, N = 96
//            the aggregates are computed by the producers from the Psrc tile (pna_agg.cuh, the arithmetic of the
//            stand-alone kernel); epilogue: identity + amplification + attenuation groups combined with the node's
//            degree scalers, + bias -> written straight back into the ring as the A operand of phase C
//   phase C  x <- relu(O . Wlin'^T + b') + x             2 k-blocks, N = 64 (BatchNorm folded into Wlin', b')
// Warps 0-15 produce operands (two groups of 256 threads, one per ring stage) and run the epilogues; warp 16 issues
// the MMAs.  Graphs with more than 128 nodes cannot be tiled; the host only takes this path for stores that have none.
namespace tc {
constexpr int MK_F = 64, MK_T = 2, MK_FO = 32, MK_C = MK_T * MK_F;
constexpr int MK_ROWS = 128;
constexpr int MK_XLD = 68, MK_PLD = 132;               // padded row strides (floats): conflict-free row-per-lane access
constexpr int MK_SLOTS = 1024;                         // CSR slots (in-edges incl. self-loops) per tile
constexpr int MK_PROD = 512, MK_THREADS = MK_PROD + 32;
constexpr uint32_t MK_APLANE = MK_ROWS * ROW_BYTES;    // 16 KB
constexpr uint32_t MK_A_BYTES = 2 * MK_APLANE;         // hi + lo
constexpr uint32_t MK_B_BYTES = 2 * 96 * ROW_BYTES;    // widest weight k-block (N = 96), hi + lo
constexpr uint32_t MK_STAGE = MK_A_BYTES + MK_B_BYTES; // 56 KB
constexpr uint32_t MK_OFF_X = 2 * MK_STAGE;
constexpr uint32_t MK_OFF_PS = MK_OFF_X + MK_ROWS * MK_XLD * 4;
constexpr uint32_t MK_OFF_CSR = MK_OFF_PS + MK_ROWS * MK_PLD * 4;
constexpr uint32_t MK_CSR_BYTES = 528 + 2 * MK_SLOTS;
constexpr uint32_t MK_SMEM = MK_OFF_CSR + MK_CSR_BYTES + 1024;      // + alignment slack
// packed weight images of one layer: [proj (t, kb) 4 x 16 KB | post (t, kb) 20 x 24 KB | lin (kb) 2 x 16 KB]
constexpr uint32_t MK_IMG_PROJ = 2 * 64 * ROW_BYTES, MK_IMG_POST = 2 * 96 * ROW_BYTES, MK_IMG_LIN = 2 * 64 * ROW_BYTES;
constexpr uint32_t MK_IMG_POST_OFF = 4 * MK_IMG_PROJ, MK_IMG_LIN_OFF = MK_IMG_POST_OFF + 20 * MK_IMG_POST;
constexpr uint32_t MK_IMG_LAYER = MK_IMG_LIN_OFF + 2 * MK_IMG_LIN;
constexpr int MK_KB_LAYER = 26;
static_assert(MK_SMEM <= 232448, "tile kernel shared memory");
static_assert(MK_STAGE % 1024 == 0 && MK_A_BYTES % 1024 == 0, "operand tiles are 1024-byte aligned");

struct MkParams {
    const float* x0;
    float* x_out;
    const int* rowptr;
    const int* col_src;
    const uint8_t* col_code;
    const int2* tiles;
    const int* num_tiles;
    const uint8_t* images;
    const float* tab;               // [L, 60, C]
    const float* bpre;              // bias of pre_nns[t][0] of layer 0 tower 0
    int64_t bpre_ts, bpre_ls;
    const float* bpost;
    int64_t bpost_ts, bpost_ls;
    const float* blin;              // [L, H]  (BatchNorm folded)
    const float* avg_deg_log;       // [L]
    int L, skip;
};

__device__ __forceinline__ void mk_bar_producers() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

__global__ void __launch_bounds__(MK_THREADS, 1) pna_infer_tiles_kernel(const MkParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[2], empty_bar[2], acc_bar;
    __shared__ uint32_t tmem_base_s;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    float* const X = reinterpret_cast<float*>(smem + MK_OFF_X);
    float* const PS = reinterpret_cast<float*>(smem + MK_OFF_PS);
    int* const rp = reinterpret_cast<int*>(smem + MK_OFF_CSR);
    uint8_t* const csrc = smem + MK_OFF_CSR + 528;
    uint8_t* const ccode = csrc + MK_SLOTS;

    if (warp == 16) tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(&full_bar[s], 256);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(&acc_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const int ntiles = __ldg(p.num_tiles);
    uint32_t kc = 0;                 // k-blocks so far (ring stage = kc & 1, use = kc >> 1); identical in every thread
    uint32_t accph = 0;              // accumulator hand-overs so far

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int2 tr = p.tiles[tile];
        const int n0 = tr.x, rows = tr.y - tr.x;
        if (warp == 16) {
            // ================================ MMA issue ================================
            for (int l = 0; l < p.L; ++l) {
                for (int i = 0; i < MK_KB_LAYER; ++i, ++kc) {
                    int N, kb;
                    uint32_t dcol;
                    bool last;
                    if (i < 4) { N = 64; kb = i & 1; dcol = (uint32_t)(i >> 1) * 64u; last = i == 3; }
                    else if (i < 24) { const int j = i - 4; N = 96; kb = j % 10; dcol = 128u + (uint32_t)(j / 10) * 96u; last = i == 23; }
                    else { N = 64; kb = i - 24; dcol = 320u; last = i == 25; }
                    const uint32_t s = kc & 1u, use = kc >> 1;
                    mbar_wait(&full_bar[s], use & 1u);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t idesc = instr_desc(MODE_TF32X3, N, 0);
                        const uint32_t a_hi = smem_u32(smem + s * MK_STAGE), a_lo = a_hi + MK_APLANE;
                        const uint32_t b_hi = a_hi + MK_A_BYTES, b_lo = b_hi + (uint32_t)N * ROW_BYTES;
                        const uint64_t dah0 = smem_desc_sw128(a_hi), dal0 = smem_desc_sw128(a_lo);
                        const uint64_t dbh0 = smem_desc_sw128(b_hi), dbl0 = smem_desc_sw128(b_lo);
                        const uint32_t d = tmem + dcol;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {                  // 4 K steps of 8 tf32 (32 bytes)
                            const uint64_t o = (uint64_t)(j * 2);
                            if (j == 0) umma<MODE_TF32X3>(d, dal0, dbh0, idesc, kb == 0 ? 0u : 1u);
                            else umma_acc<MODE_TF32X3>(d, dal0 + o, dbh0 + o, idesc);
                            umma_acc<MODE_TF32X3>(d, dah0 + o, dbl0 + o, idesc);
                            umma_acc<MODE_TF32X3>(d, dah0 + o, dbh0 + o, idesc);
                        }
                        umma_commit(&empty_bar[s]);
                        if (last) umma_commit(&acc_bar);
                    }
                    __syncwarp();
                }
            }
            continue;
        }
        // ================================ producers / epilogues (warps 0-15) ================================
        const int grp = tid >> 8, pt = tid & 255;
        const int q = warp & 3, cg = warp >> 2, row = q * 32 + lane;          // epilogue: TMEM lane quarter, column group
        const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
        // ---- stage the tile: node features and the CSR slice (tile-local indices)
        const int slot0 = __ldg(p.rowptr + n0);
        const int total = __ldg(p.rowptr + n0 + rows) - slot0;
        for (int it = tid; it < MK_ROWS * 16; it += MK_PROD) {
            const int r = it >> 4, c4 = it & 15;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < rows) v = ldg4(p.x0 + (size_t)(n0 + r) * MK_F + 4 * c4);
            *reinterpret_cast<float4*>(X + r * MK_XLD + 4 * c4) = v;
        }
        for (int r = tid; r <= MK_ROWS; r += MK_PROD) rp[r] = r <= rows ? __ldg(p.rowptr + n0 + r) - slot0 : total;
        for (int e = tid; e < total; e += MK_PROD) {
            csrc[e] = (uint8_t)(__ldg(p.col_src + slot0 + e) - n0);
            ccode[e] = __ldg(p.col_code + slot0 + e);
        }
        mk_bar_producers();

        for (int l = 0; l < p.L; ++l) {
            const uint8_t* const img = p.images + (size_t)l * MK_IMG_LAYER;
            const float* const tab = p.tab + (size_t)l * GNN_NUM_BOND_CODES * MK_C;
            // ---------------- phases A and B: 24 k-blocks, group g produces the k-blocks of ring stage g
            for (int i = grp; i < 24; i += 2) {
                if (i == 4 + grp) {
                    // phase A accumulators -> Psrc tile (+ bias); every producer takes part, then phase B may gather
                    mbar_wait(&acc_bar, accph & 1u);
                    tc_fence_after();
                    {
                        const int t = cg >> 1, f0 = (cg & 1) * 32;
                        const float* b = p.bpre + (size_t)l * p.bpre_ls + (size_t)t * p.bpre_ts + f0;
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            float v[16];
                            tmem_ld16(lane_addr + (uint32_t)(cg * 32 + h * 16), v);
                            float* dst = PS + row * MK_PLD + cg * 32 + h * 16;
#pragma unroll
                            for (int j = 0; j < 16; j += 4) {
                                const float4 bb = ldg4(b + h * 16 + j);
                                *reinterpret_cast<float4*>(dst + j) =
                                    make_float4(v[j] + bb.x, v[j + 1] + bb.y, v[j + 2] + bb.z, v[j + 3] + bb.w);
                            }
                        }
                    }
                    tc_fence_before();
                    mk_bar_producers();
                }
                const uint32_t kcur = kc + (uint32_t)i, s = kcur & 1u, use = kcur >> 1;
                if (use > 0) mbar_wait(&empty_bar[s], (use - 1) & 1u);
                uint8_t* const stage = smem + s * MK_STAGE;
                if (i < 4) {
                    const int t = i >> 1, kb = i & 1;
                    if (pt == 0) {
                        mbar_expect_tx(&full_bar[s], MK_IMG_PROJ);
                        bulk_g2s(stage + MK_A_BYTES, img + (size_t)(t * 2 + kb) * MK_IMG_PROJ, MK_IMG_PROJ, &full_bar[s]);
                    }
                    for (int it = pt; it < MK_ROWS * 8; it += 256) {
                        const int r = it >> 3, c = it & 7;
                        put_chunk_tf32<MODE_TF32X3>(stage, MK_APLANE, sw128_offset(r, c),
                                                    *reinterpret_cast<const float4*>(X + r * MK_XLD + kb * 32 + 4 * c));
                    }
                } else {
                    const int j = i - 4, t = j / 10, kb = j % 10;
                    if (pt == 0) {
                        mbar_expect_tx(&full_bar[s], MK_IMG_POST);
                        bulk_g2s(stage + MK_A_BYTES, img + MK_IMG_POST_OFF + (size_t)(t * 10 + kb) * MK_IMG_POST, MK_IMG_POST,
                                 &full_bar[s]);
                    }
                    if (kb < 2) {
                        for (int it = pt; it < MK_ROWS * 8; it += 256) {
                            const int r = it >> 3, c = it & 7;
                            put_chunk_tf32<MODE_TF32X3>(stage, MK_APLANE, sw128_offset(r, c),
                                                        *reinterpret_cast<const float4*>(X + r * MK_XLD + kb * 32 + 4 * c));
                        }
                    } else {
                        // aggregates of 8 features (4 per thread) of tower t for one row: mean / min / max / std
                        const int r = pt >> 1, h = pt & 1, col = t * MK_F + 8 * (kb - 2) + 4 * h;
                        const int beg = rp[r], cnt = rp[r + 1] - beg;
                        float4 mean = make_float4(0.f, 0.f, 0.f, 0.f), sd = mean, mn = mean, mx = mean;
                        if (cnt > 0) {
                            // the loads of the first PF in-edges go out together (slots beyond the segment re-read
                            // its first message, which leaves every statistic unchanged); molecules rarely have more
                            constexpr int PF = 5;
                            float4 pr[PF], tb[PF];
#pragma unroll
                            for (int e = 0; e < PF; ++e) {
                                const int slot = beg + (e < cnt ? e : 0);
                                pr[e] = *reinterpret_cast<const float4*>(PS + (int)csrc[slot] * MK_PLD + col);
                                tb[e] = ldg4(tab + (int)ccode[slot] * MK_C + col);
                            }
                            PnaAcc<false> acc;
                            acc.first(pna_add4(pr[0], tb[0]));
#pragma unroll
                            for (int e = 1; e < PF; ++e) acc.next(pna_add4(pr[e], tb[e]));
                            for (int e = PF; e < cnt; ++e)
                                acc.next(pna_add4(*reinterpret_cast<const float4*>(PS + (int)csrc[beg + e] * MK_PLD + col),
                                                  ldg4(tab + (int)ccode[beg + e] * MK_C + col)));
                            acc.finish(cnt, mean, sd);
                            mn = acc.mn;
                            mx = acc.mx;
                        }
                        put_chunk_tf32<MODE_TF32X3>(stage, MK_APLANE, sw128_offset(r, 0 + h), mean);
                        put_chunk_tf32<MODE_TF32X3>(stage, MK_APLANE, sw128_offset(r, 2 + h), mn);
                        put_chunk_tf32<MODE_TF32X3>(stage, MK_APLANE, sw128_offset(r, 4 + h), mx);
                        put_chunk_tf32<MODE_TF32X3>(stage, MK_APLANE, sw128_offset(r, 6 + h), sd);
                    }
                }
                fence_async_smem();
                mbar_arrive(&full_bar[s]);
            }
            accph += 1;              // phase A hand-over consumed above
            // ---------------- phase B accumulators -> O (combined, + bias) -> A operand of phase C in the ring
            mbar_wait(&acc_bar, accph & 1u);
            tc_fence_after();
            accph += 1;
            {
                const int t = cg >> 1, half = cg & 1;
                const uint32_t kcur = kc + 24u + (uint32_t)t, s = kcur & 1u, use = kcur >> 1;
                if (use > 0) mbar_wait(&empty_bar[s], (use - 1) & 1u);
                uint8_t* const stage = smem + s * MK_STAGE;
                if (q == 0 && half == 0 && lane == 0) {
                    mbar_expect_tx(&full_bar[s], MK_IMG_LIN);
                    bulk_g2s(stage + MK_A_BYTES, img + MK_IMG_LIN_OFF + (size_t)t * MK_IMG_LIN, MK_IMG_LIN, &full_bar[s]);
                }
                const int deg = rp[row + 1] - rp[row];
                const float avg = __ldg(p.avg_deg_log + l);
                float amp, att;
                pna_scalers(deg, avg, amp, att);
                const uint32_t base = lane_addr + 128u + (uint32_t)t * 96u + (uint32_t)half * 16u;
                float v[16], d1[16], d2[16];
                tmem_ld16(base, v);
                tmem_ld16(base + 32u, d1);
                tmem_ld16(base + 64u, d2);
                const float* b = p.bpost + (size_t)l * p.bpost_ls + (size_t)t * p.bpost_ts + half * 16;
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    const float4 bb = ldg4(b + j);
                    const float4 o = make_float4(v[j] + amp * d1[j] + att * d2[j] + bb.x,
                                                 v[j + 1] + amp * d1[j + 1] + att * d2[j + 1] + bb.y,
                                                 v[j + 2] + amp * d1[j + 2] + att * d2[j + 2] + bb.z,
                                                 v[j + 3] + amp * d1[j + 3] + att * d2[j + 3] + bb.w);
                    put_chunk_tf32<MODE_TF32X3>(stage, MK_APLANE, sw128_offset(row, half * 4 + (j >> 2)), o);
                }
                tc_fence_before();
                fence_async_smem();
                mbar_arrive(&full_bar[s]);
            }
            // ---------------- phase C accumulator -> x = relu(y + b') (+ x)
            mbar_wait(&acc_bar, accph & 1u);
            tc_fence_after();
            accph += 1;
            {
                float v[16];
                tmem_ld16(lane_addr + 320u + (uint32_t)cg * 16u, v);
                const float* b = p.blin + (size_t)l * MK_F + cg * 16;
                float* xr = X + row * MK_XLD + cg * 16;
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    const float4 bb = ldg4(b + j);
                    float4 y = make_float4(fmaxf(v[j] + bb.x, 0.f), fmaxf(v[j + 1] + bb.y, 0.f), fmaxf(v[j + 2] + bb.z, 0.f),
                                           fmaxf(v[j + 3] + bb.w, 0.f));
                    if (p.skip) {
                        const float4 xo = *reinterpret_cast<const float4*>(xr + j);
                        y = make_float4(y.x + xo.x, y.y + xo.y, y.z + xo.z, y.w + xo.w);
                    }
                    *reinterpret_cast<float4*>(xr + j) = y;
                }
            }
            tc_fence_before();
            mk_bar_producers();
            kc += MK_KB_LAYER;
        }
        // ---- write the tile's node features back
        for (int it = tid; it < rows * 16; it += MK_PROD) {
            const int r = it >> 4, c4 = it & 15;
            *reinterpret_cast<float4*>(p.x_out + (size_t)(n0 + r) * MK_F + 4 * c4) =
                *reinterpret_cast<const float4*>(X + r * MK_XLD + 4 * c4);
        }
        mk_bar_producers();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 16) tmem_dealloc(tmem, 512);
}

// Greedy packing of consecutive graphs into tiles of <= 128 nodes and <= MK_SLOTS CSR slots.  One CTA per range of
// 2048 graphs: node / slot prefixes of the range in shared memory, next[g] (where a tile starting at graph g ends) by
// binary search in parallel, then one thread follows the chain; tiles are appended to an unordered list.
constexpr int TB_GRAPHS = 2048;
__global__ void __launch_bounds__(1024)
build_tiles_kernel(const int* __restrict__ graph_ptr, const int* __restrict__ rowptr, int num_graphs,
                   const int* __restrict__ valid_graphs, int2* __restrict__ tiles, int max_tiles, int* num_tiles, int* err) {
    __shared__ int gp[TB_GRAPHS + 1], sp[TB_GRAPHS + 1], nxt[TB_GRAPHS];
    int G = num_graphs;
    if (valid_graphs != nullptr) G = min(G, __ldg(valid_graphs));
    const int g0 = blockIdx.x * TB_GRAPHS;
    const int cnt = min(TB_GRAPHS, G - g0);
    if (cnt <= 0) return;
    for (int i = threadIdx.x; i <= cnt; i += blockDim.x) {
        const int n = __ldg(graph_ptr + g0 + i);
        gp[i] = n;
        sp[i] = __ldg(rowptr + n);
    }
    __syncthreads();
    for (int g = threadIdx.x; g < cnt; g += blockDim.x) {
        // largest e in (g, cnt] with gp[e] - gp[g] <= 128 and sp[e] - sp[g] <= MK_SLOTS (both prefixes are monotone)
        int lo = g, hi = cnt;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (gp[mid] - gp[g] <= MK_ROWS && sp[mid] - sp[g] <= MK_SLOTS) lo = mid;
            else hi = mid - 1;
        }
        nxt[g] = lo;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int g = 0;
        while (g < cnt) {
            int e = nxt[g];
            if (e == g) {                    // a single graph exceeds a tile: reported, skipped
                if (err) atomicOr(err, GNN_INPUT_ERR_EDGE_INDEX);
                e = g + 1;
            } else if (gp[e] > gp[g]) {
                const int slot = atomicAdd(num_tiles, 1);
                if (slot < max_tiles) tiles[slot] = make_int2(gp[g], gp[e]);
            }
            g = e;
        }
    }
}
}  // namespace tc

extern "C" int64_t gnn_tc_infer_image_bytes(int num_layers) { return (int64_t)num_layers * tc::MK_IMG_LAYER; }

extern "C" int gnn_tc_infer_pack(const float* wpre, int64_t wpre_tower_stride, const float* Vx, const float* wpost,
                                 int64_t wpost_tower_stride, int64_t layer_stride, const float* wlin_folded,
                                 int num_layers, int F, int F_out, int towers, void* images, int64_t image_bytes,
                                 void* stream) {
    if (F != MK_F || F_out != MK_FO || towers != MK_T) return GNN_ERR_UNSUPPORTED;
    if (num_layers < 1 || image_bytes < gnn_tc_infer_image_bytes(num_layers)) return GNN_ERR_WORKSPACE;
    if ((reinterpret_cast<uintptr_t>(images) & 127) != 0) return GNN_ERR_BAD_ARG;
    cudaStream_t st = gnn_stream(stream);
    for (int l = 0; l < num_layers; ++l) {
        uint8_t* img = reinterpret_cast<uint8_t*>(images) + (size_t)l * MK_IMG_LAYER;
        VProjWeight pw{wpre + l * layer_stride, wpre_tower_stride, F, towers, 1, 1};
        pack_b_kernel<MODE_TF32X3, VProjWeight><<<dim3(2 * 2, 2, 1), PROD_GROUP_THREADS, 0, st>>>(pw, 64, (int64_t)F, 2, 2, img);
        VPostWeight ow{wpost + l * layer_stride, wpost_tower_stride, F, F_out, MK_FO,
                       Vx + (size_t)l * towers * 3 * F_out * F, 64};
        pack_b_kernel<MODE_TF32X3, VPostWeight><<<dim3(10 * 3, 1, 2), PROD_GROUP_THREADS, 0, st>>>(
            ow, 96, (int64_t)(64 + 4 * F), 1, 10, img + MK_IMG_POST_OFF);
        VRowMajor lw{wlin_folded + (size_t)l * F * F, F, F, F, 0, 1};
        pack_b_kernel<MODE_TF32X3, VRowMajor><<<dim3(2 * 2, 1, 1), PROD_GROUP_THREADS, 0, st>>>(lw, 64, (int64_t)F, 1, 2,
                                                                                                img + MK_IMG_LIN_OFF);
    }
    return gnn_launch_status();
}

extern "C" int gnn_tc_infer_tiles(const int32_t* graph_ptr, const int32_t* rowptr, int64_t num_graphs,
                                  const int32_t* valid_graphs, int32_t* tiles, int64_t max_tiles, int32_t* num_tiles,
                                  int32_t* err_flag, void* stream) {
    if (num_graphs < 1 || max_tiles < 1) return GNN_ERR_BAD_ARG;
    cudaStream_t st = gnn_stream(stream);
    cudaError_t e = cudaMemsetAsync(num_tiles, 0, sizeof(int), st);
    if (e != cudaSuccess) return (int)e;
    const int blocks = (int)gnn_ceil_div(num_graphs, TB_GRAPHS);
    build_tiles_kernel<<<blocks, 1024, 0, st>>>(graph_ptr, rowptr, (int)num_graphs, valid_graphs,
                                                reinterpret_cast<int2*>(tiles), (int)max_tiles, num_tiles, err_flag);
    return gnn_launch_status();
}

extern "C" int gnn_tc_pna_stack_infer(const float* x0, const int32_t* rowptr, const int32_t* col_src,
                                      const uint8_t* col_code, const int32_t* tiles, const int32_t* num_tiles,
                                      const void* images, const float* tab, const float* bpre, int64_t bpre_tower_stride,
                                      int64_t bpre_layer_stride, const float* bpost, int64_t bpost_tower_stride,
                                      int64_t bpost_layer_stride, const float* blin, const float* avg_deg_log,
                                      int num_layers, int F, int F_out, int towers, int skip, float* x_out,
                                      void* stream) {
    if (F != MK_F || F_out != MK_FO || towers != MK_T) return GNN_ERR_UNSUPPORTED;
    if (num_layers < 1) return GNN_ERR_BAD_ARG;
    REQUIRE_VEC(x0, F);
    REQUIRE_VEC(x_out, F);
    REQUIRE_VEC(tab, (int64_t)MK_C);
    REQUIRE_VEC(bpre, bpre_tower_stride);
    REQUIRE_VEC(bpre, bpre_layer_stride);
    REQUIRE_VEC(bpost, bpost_tower_stride);
    REQUIRE_VEC(bpost, bpost_layer_stride);
    REQUIRE_VEC(blin, F);
    MkParams p{x0, x_out, rowptr, col_src, col_code, reinterpret_cast<const int2*>(tiles), num_tiles,
               reinterpret_cast<const uint8_t*>(images), tab, bpre, bpre_tower_stride, bpre_layer_stride, bpost,
               bpost_tower_stride, bpost_layer_stride, blin, avg_deg_log, num_layers, skip};
    cudaError_t e = cudaFuncSetAttribute(pna_infer_tiles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MK_SMEM);
    if (e != cudaSuccess) return (int)e;
    pna_infer_tiles_kernel<<<gnn_sm_count(), MK_THREADS, MK_SMEM, gnn_stream(stream)>>>(p);
    return gnn_launch_status();
}
