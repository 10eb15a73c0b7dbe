// tools/ubench/tma_bw.cu -- aggregate bandwidth of cp.async.bulk (global -> shared, mbarrier completion) on B200 as a function of
// the copy size and the number of copies in flight per SM: what the stage ring of tau_group_tc_kernel can expect.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ubench/tma_bw tools/ubench/tma_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}

// every CTA streams its own contiguous slice of `src` through a ring of `stages` buffers of `bytes` each; nothing is computed
__global__ void __launch_bounds__(128, 1) k(const unsigned char *src, size_t per_cta, int bytes, int stages, int hint, int split, unsigned long long *sink)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bars[8];
    const uint32_t bar0 = smem_u32(bars);
    if (threadIdx.x == 0) {
        for (int i = 0; i < stages; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8 * i));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const unsigned char *p = src + (size_t)blockIdx.x * per_cta;
    const int n = (int)(per_cta / bytes);
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    auto issue = [&](int i) {
        const uint32_t s = i % stages, bar = bar0 + 8 * s, dst = smem_u32(smem) + s * bytes;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        const int piece = bytes / split;
        for (int q = 0; q < split; q++) {
            if (hint) asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                                   ::"r"(dst + q * piece), "l"(p + (size_t)i * bytes + (size_t)q * piece), "r"(piece), "r"(bar), "l"(pol) : "memory");
            else asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                              ::"r"(dst + q * piece), "l"(p + (size_t)i * bytes + (size_t)q * piece), "r"(piece), "r"(bar) : "memory");
        }
    };
    for (int i = 0; i < stages && i < n; i++) issue(i);
    for (int i = 0; i < n; i++) {
        mbar_wait(bar0 + 8 * (i % stages), (i / stages) & 1);
        if (i + stages < n) issue(i + stages);
    }
    sink[blockIdx.x] = smem[0];
}

int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const size_t per_cta = (size_t)4 << 20, total = per_cta * sms;      // 592 MB: larger than L2
    unsigned char *src; unsigned long long *sink;
    cudaMalloc(&src, total); cudaMalloc(&sink, sms * 8);
    cudaMemset(src, 1, total);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    printf("%d SMs, %zu MB streamed per launch (HBM, cold L2)\n", sms, total >> 20);
    for (int hint = 0; hint < 2; hint++)
        for (int bytes : {8192, 16384, 32768, 65536})
            for (int stages : {1, 2, 3, 4, 6})
                for (int split : {1, 4}) {
                    if ((size_t)bytes * stages > 192 * 1024) continue;
                    k<<<sms, 128, bytes * stages>>>(src, per_cta, bytes, stages, hint, split, sink);   // warm-up
                    cudaEventRecord(a);
                    k<<<sms, 128, bytes * stages>>>(src, per_cta, bytes, stages, hint, split, sink);
                    cudaEventRecord(b); cudaEventSynchronize(b);
                    float ms = 0; cudaEventElapsedTime(&ms, a, b);
                    printf("copy %6d B x %d in flight, %d piece(s), evict_first %d: %7.1f GB/s  (%.2f us per copy per SM)\n", bytes, stages, split, hint,
                           total / (ms * 1e-3) / 1e9, ms * 1e3 / (per_cta / bytes));
                }
    return 0;
}
