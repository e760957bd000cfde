// Microbenchmark (debug aid, run on the GPU box): issue rate of tcgen05.mma for the shapes the GEMM kernels use.
// main() runs rate3_kernel (hoisted descriptors, optional concurrent tcgen05.ld); rate_kernel is the first version with
// descriptors rebuilt in the loop (kept for reference, see tools/mma_noise.cu for the scheduler-contention experiment).
// One CTA per SM, one thread issues REPS x (ksteps MMAs) back to back from shared-memory tiles, then commits and
// waits; prints cycles per MMA for kind::tf32 / kind::f16, N in {64, 96, 128, 192, 256}, K-major SWIZZLE_128B.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_rate tools/mma_rate.cu && tools/mma_rate
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ inline uint32_t instr_desc(int bf16, int n) {
    uint32_t fmt = bf16 ? 1u : 2u;
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

template <int BF16>
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int reps, int distinct, int use_other_acc, long long* out) {
    extern __shared__ uint8_t raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x;
    // A planes: 2 x [128 x 128 B]; B planes: 2 x [256 x 128 B]  -> 96 KB, zero-filled
    for (int i = tid; i < (2 * 16384 + 2 * 32768) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) {
        const uint32_t a_hi = smem_u32(smem), a_lo = a_hi + 16384, b_hi = a_hi + 32768, b_lo = b_hi + 32768;
        const uint32_t idesc = instr_desc(BF16, N);
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {                   // 4 k-steps of 32 bytes = one 128-byte k-block
                const uint32_t koff = j * 32;
                for (int v = 0; v < distinct; ++v) {        // distinct = 3: (lo,hi), (hi,lo), (hi,hi) as TF32x3 does
                    const uint64_t da = desc_sw128((v == 0 ? a_lo : a_hi) + koff);
                    const uint64_t db = desc_sw128((v == 1 ? b_lo : b_hi) + koff);
                    const uint32_t d = tmem + (use_other_acc ? (uint32_t)(v & 1) * 256u : 0u);
                    if (BF16)
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
                    else
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
                }
            }
        }
        const long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        const long long t2 = clock64();
        if (blockIdx.x == 0) {
            out[0] = t1 - t0;
            out[1] = t2 - t0;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// TF32x3 issue pattern of gemm_tc.cu with everything hoisted: 16 descriptors precomputed, 12 MMAs per k-block fully
// unrolled, optionally one tcgen05.commit per k-block (as the ring release does)
template <int BF16>
__global__ void __launch_bounds__(160, 1) rate3_kernel(int N, int reps, int commit_each, long long* out) {
    extern __shared__ uint8_t raw[];
    __shared__ __align__(8) uint64_t bar, bar2;
    __shared__ uint32_t tmem_base_s;
    __shared__ volatile int stop_flag;
    const int issuer = (commit_each & 16) ? 128 : 0;
    if (threadIdx.x == 0) stop_flag = 0;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x;
    for (int i = tid; i < (2 * 16384 + 2 * 32768) / 4; i += blockDim.x) {
        // commit_each bit 1: random finite fp32 operands instead of zeros (data-dependent power / pacing?)
        uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        reinterpret_cast<uint32_t*>(smem)[i] = (commit_each & 2) ? (0x3f000000u | (h & 0x80ffffffu)) : 0u;
    }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar2)), "r"(1 << 20));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    if ((commit_each & 16) && tid < 128) {
        // four warps read a second accumulator region (columns 256..) with tcgen05.ld while the MMAs run
        uint32_t acc = 0;
        const uint32_t taddr = tmem + 256u + ((uint32_t)((tid >> 5) * 32) << 16);
        while (!stop_flag) {
            uint32_t r[16];
#pragma unroll 1
            for (int c = 0; c < 96; c += 16) {
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                               "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                             : "r"(taddr + c));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                acc += r[0] ^ r[15];
            }
            if (commit_each & 32) __nanosleep(2000);      // duty cycle like a real epilogue
        }
        if (acc == 0x12345678u) out[1] = 0;
    }
    if (tid == issuer) {
        const uint32_t a_hi = smem_u32(smem), a_lo = a_hi + 16384, b_hi = a_hi + 32768, b_lo = b_hi + 32768;
        const uint32_t idesc = instr_desc(BF16, N);
        uint64_t dah[4], dal[4], dbh[4], dbl[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            dah[j] = desc_sw128(a_hi + j * 32); dal[j] = desc_sw128(a_lo + j * 32);
            dbh[j] = desc_sw128(b_hi + j * 32); dbl[j] = desc_sw128(b_lo + j * 32);
        }
#define MMA(DA, DB)                                                                                                      \
    if (BF16) asm volatile("tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, 1;" ::"r"(tmem), "l"(DA), "l"(DB), "r"(idesc) : "memory"); \
    else asm volatile("tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, 1;" ::"r"(tmem), "l"(DA), "l"(DB), "r"(idesc) : "memory");
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                MMA(dal[j], dbh[j]) MMA(dah[j], dbl[j]) MMA(dah[j], dbh[j])
            }
            if (commit_each & 4) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (commit_each & 8) {
                uint32_t ok = 0;
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(1u) : "memory");
            }
            if (commit_each & 1)
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)) : "memory");
        }
        const long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        const long long t2 = clock64();
        if (blockIdx.x == 0) {
            out[0] = t1 - t0;
            out[1] = t2 - t0;
        }
        stop_flag = 1;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

int main() {
    {
        long long* out;
        cudaMallocManaged(&out, 2 * sizeof(long long));
        const int smem = 98304 + 1024;
        cudaFuncSetAttribute(rate3_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(rate3_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        for (int bf16 = 0; bf16 < 2; ++bf16)
            for (int commit : {0, 16, 48})
                for (int n : {96}) {
                    for (int it = 0; it < 2; ++it) {
                        if (bf16) rate3_kernel<1><<<148, 160, smem>>>(n, 4096, commit, out);
                        else rate3_kernel<0><<<148, 160, smem>>>(n, 4096, commit, out);
                        cudaError_t e = cudaDeviceSynchronize();
                        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                    }
                    printf("hoisted %s N=%3d flags(16 concurrent tcgen05.ld, 32 duty-cycled)=%2d : issue %.1f cyc/MMA, complete %.1f cyc/MMA (floor %.0f)\n",
                           bf16 ? "bf16" : "tf32", n, commit, out[0] / (4096.0 * 12), out[1] / (4096.0 * 12), 128.0 * n / 256.0);
                }
    }
    return 0;
}
