// tools/ubench/umma_issue.cu -- how long do the 16 tcgen05.mma (kind::f16, M=128, K=16) of one work item of tau_group_tc_kernel
// take, as a function of the shared-memory operand layout (descriptor LBO / SBO / swizzle) and of N?  Timing only: the operand
// contents are whatever shared memory holds.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ubench/umma_issue tools/ubench/umma_issue.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout)
{
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) | ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) |
           (1ull << 46) | ((uint64_t)layout << 61);
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

struct Cfg { uint32_t lbo_a, sbo_a, step_a, lbo_b, sbo_b, step_b, layout, N, nk, atom_a, atom_b, nchain; };

__global__ void __launch_bounds__(128, 1) k(Cfg c, int reps, long long *out)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t tmem_s;
    for (int i = threadIdx.x; i < (64 + 32) * 1024 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;   // fp16 1.0
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_s;
    if (threadIdx.x == 0) {
        const uint32_t idesc = (1u << 4) | ((c.N >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t a_s = smem_u32(smem), b_s = smem_u32(smem + 64 * 1024);
        long long t_issue = 0, t_all = 0;
        for (int r = 0; r < reps; r++) {
            const long long t0 = clock64();
            if (c.layout == 0 && c.nchain == 0 && c.nk == 16) {       // the kernel's form: descriptors advanced by an add, unrolled
                uint64_t a = desc(a_s, c.lbo_a, c.sbo_a, 0), b = desc(b_s, c.lbo_b, c.sbo_b, 0);
                const uint64_t da = c.step_a >> 4, db = c.step_b >> 4;
#pragma unroll
                for (uint32_t kk = 0; kk < 16; kk++, a += da, b += db) mma(tm, a, b, idesc, kk ? 1u : 0u);
            } else
            for (uint32_t kk = 0; kk < c.nk; kk++)
            {
                const uint32_t oa = c.layout ? (kk >> 2) * c.atom_a + (kk & 3) * c.step_a : kk * c.step_a;
                const uint32_t ob = c.layout ? (kk >> 2) * c.atom_b + (kk & 3) * c.step_b : kk * c.step_b;
                const uint32_t nch = c.nchain ? c.nchain : 1u;
                mma(tm + (kk % nch) * 64, desc(a_s + oa, c.lbo_a, c.sbo_a, c.layout), desc(b_s + ob, c.lbo_b, c.sbo_b, c.layout), idesc, kk >= nch ? 1u : 0u);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            const long long t1 = clock64();
            mbar_wait(smem_u32(&bar), r & 1);
            const long long t2 = clock64();
            if (r) { t_issue += t1 - t0; t_all += t2 - t0; }
        }
        out[2 * blockIdx.x] = t_issue / (reps - 1); out[2 * blockIdx.x + 1] = t_all / (reps - 1);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}

int main()
{
    long long *out;
    cudaMalloc(&out, 2 * 148 * sizeof(long long));
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    struct { const char *name; Cfg c; } v[] = {
        // item of C3: 128 rows x K = 256 fp16 (KC = 32 chunks of 16 B), 16 K steps
        {"no swizzle, [rowgroup][chunk] (current): LBO 128 SBO 4096, N 48", {128, 4096, 256, 128, 4096, 256, 0, 48, 16}},
        {"no swizzle, [rowgroup][chunk] (current), N 64                  ", {128, 4096, 256, 128, 4096, 256, 0, 64, 16}},
        {"no swizzle, [rowgroup][chunk], SBO padded 4096+128, N 48       ", {128, 4224, 256, 128, 4224, 256, 0, 48, 16}},
        {"no swizzle, [chunk][rowgroup]: A LBO 2048 SBO 128, B LBO 768   ", {2048, 128, 4096, 768, 128, 1536, 0, 48, 16}},
        {"128B swizzle (SBO 1024, K step 32 B inside the atom), N 48     ", {16, 1024, 32, 16, 1024, 32, 2, 48, 4, 16384, 6144}},
        {"128B swizzle, 16 K steps over 4 atoms (A atom 16 KB, B 6 KB)   ", {16, 1024, 32, 16, 1024, 32, 2, 48, 16, 16384, 6144}},
        {"no swizzle (current), 1 K step                                  ", {128, 4096, 256, 128, 4096, 256, 0, 48, 1}},
        {"no swizzle (current), 4 K steps                                 ", {128, 4096, 256, 128, 4096, 256, 0, 48, 4}},
        {"no swizzle (current), 16 K steps on 2 accumulators              ", {128, 4096, 256, 128, 4096, 256, 0, 48, 16, 0, 0, 2}},
        {"no swizzle (current), 16 K steps on 4 accumulators              ", {128, 4096, 256, 128, 4096, 256, 0, 48, 16, 0, 0, 4}},
        {"no swizzle (current), 16 K steps on 8 accumulators              ", {128, 4096, 256, 128, 4096, 256, 0, 48, 16, 0, 0, 8}},
        {"no swizzle, B overlapped (SBO 128), N 16                        ", {128, 4096, 256, 128, 128, 256, 0, 16, 16}},
        {"no swizzle, B overlapped (SBO 128), N 48                        ", {128, 4096, 256, 128, 128, 256, 0, 48, 16}},
        {"no swizzle, B overlapped (SBO 128), N 128                       ", {128, 4096, 256, 128, 128, 256, 0, 128, 16}},
        {"no swizzle, B overlapped (SBO 128), N 256                       ", {128, 4096, 256, 128, 128, 256, 0, 256, 16}},
    };
    for (auto &x : v) {
        for (int ctas : {1, 148}) {
            k<<<ctas, 128, 96 * 1024>>>(x.c, 50, out);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[2];
            cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
            printf("%s  ctas %3d  issue %6lld clk  issue+complete %6lld clk  (%s)\n", x.name, ctas, h[0], h[1], cudaGetErrorString(e));
        }
    }
    return 0;
}
