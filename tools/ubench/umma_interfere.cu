// tools/ubench/umma_interfere.cu -- does a stream of small tcgen05.mma (M=128, N=48, K=16: the shape of tau_group_tc_kernel) slow
// down what other warps of the SM execute (FP64 FMA, FP64->FP32 conversion, MUFU, SHFL, shared-memory loads/stores), and vice versa?
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ubench/umma_interfere tools/ubench/umma_interfere.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) | ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// warp 0: MMA batches (if mma_on); warps 1..nw: `op` loop
__global__ void __launch_bounds__(544, 1) k(int op, int mma_on, int iters, int batches, long long *out)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t tmem_s;
    __shared__ volatile int stop;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) {
        stop = 0;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        if (lane == 0 && mma_on) {
            const uint32_t idesc = (1u << 4) | ((48u >> 3) << 17) | ((128u >> 4) << 24);
            const uint32_t a_s = smem_u32(smem), b_s = smem_u32(smem + 64 * 1024);
            long long tot = 0, worst = 0;
            int n = 0;
            for (int r = 0; (batches > 0) ? (r < batches) : !stop; r++) {
                const long long t0 = clock64();
                uint64_t a = desc(a_s, 128, 4096), b = desc(b_s, 128, 4096);
#pragma unroll
                for (uint32_t kk = 0; kk < 16; kk++, a += 16, b += 16) mma(tm, a, b, idesc, kk ? 1u : 0u);
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
                mbar_wait(smem_u32(&bar), r & 1);
                const long long dt = clock64() - t0;
                if (r) { tot += dt; n++; if (dt > worst) worst = dt; }
            }
            out[0] = n ? tot / n : 0; out[1] = worst;
        }
    } else {
        // the op loop: 8 independent chains per lane
        double d[8]; float f[8]; uint32_t u[8];
        for (int i = 0; i < 8; i++) { d[i] = 1.0 + 1e-9 * (lane + i); f[i] = 1.0f + 1e-3f * (lane + i); u[i] = lane + i; }
        float *sf = reinterpret_cast<float *>(smem + 80 * 1024) + threadIdx.x;
        __syncwarp();
        const long long t0 = clock64();
        long long worst = 0, tprev = t0;
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (op == 0) d[i] = fma(d[i], 1.0000001, 1e-12);
                else if (op == 1) { f[i] = (float)(d[i]); d[i] += (double)f[i] * 1e-30; }          // F2F.F32.F64 (+ F2F back)
                else if (op == 2) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
                else if (op == 3) u[i] = __shfl_sync(0xffffffffu, u[i], (lane + 1) & 31);
                else if (op == 4) { sf[i * 1024] = f[i]; f[i] += sf[((i + 1) & 7) * 1024]; }         // STS + LDS
                else if (op == 5) f[i] = fmaf(f[i], 1.0000001f, 1e-12f);
            }
            if ((it & 15) == 15) { const long long t = clock64(); if (t - tprev > worst) worst = t - tprev; tprev = t; }
        }
        const long long t1 = clock64();
        double acc = 0; for (int i = 0; i < 8; i++) acc += d[i] + f[i] + u[i];
        if (acc == 1.2345) out[7] = 1;
        if (lane == 0 && warp == 1) { out[2] = (t1 - t0); out[3] = worst; }
        if (lane == 0) atomicAdd((int *)&stop, 1);            // (the MMA loop of a `batches == 0` run ends when an op warp is done)
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(64u) : "memory");
}

int main()
{
    long long *out, h[8];
    cudaMalloc(&out, 8 * sizeof(long long));
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    const char *names[] = {"DFMA", "F2F.F32.F64 + back", "MUFU.LG2", "SHFL", "STS+LDS", "FFMA"};
    // the MMA stream alone
    cudaMemset(out, 0, sizeof(h));
    k<<<1, 32, 128 * 1024>>>(0, 1, 0, 200, out);
    cudaDeviceSynchronize(); cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("MMA batch (16 x M128 N48 K16 + commit + wait) alone: mean %lld clk, worst %lld  (%s)\n", h[0], h[1], cudaGetErrorString(cudaGetLastError()));
    for (int op = 0; op < 6; op++)
        for (int nw : {4, 16}) {
            long long base[2];
            for (int mma_on = 0; mma_on < 2; mma_on++) {
                cudaMemset(out, 0, sizeof(h));
                k<<<1, 32 * (nw + 1), 128 * 1024>>>(op, mma_on, 2000, 0, out);
                cudaError_t e = cudaDeviceSynchronize(); cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
                if (!mma_on) { base[0] = h[2]; base[1] = h[3]; }
                else printf("%-20s %2d warps: op loop %8lld clk alone (worst 16-iteration gap %6lld) -> %8lld with the MMA stream (worst gap %6lld) | MMA batch mean %5lld worst %6lld  (%s)\n",
                            names[op], nw, base[0], base[1], h[2], h[3], h[0], h[1], cudaGetErrorString(e));
            }
        }
    return 0;
}
