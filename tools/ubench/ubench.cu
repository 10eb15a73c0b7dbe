// tools/ubench/ubench.cu -- pipe-rate probes used to pick the lane mapping of the grouped tau kernel (B200, sm_100a).
// Each probe runs `warps` warps per SM on every SM and reports warp-instructions per clock per SM.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

#define ITER 2048

__global__ void k_ffma(float *out, float x, float y, long long *cyc)
{
    float a[24];
#pragma unroll
    for (int i = 0; i < 24; i++) a[i] = threadIdx.x * 0.001f + i;
    float b0 = x, b1 = y, b2 = x + y, b3 = x - y;
    long long t0 = clock64();
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < 24; i += 4) {
            a[i] = fmaf(a[i], b0, b1); a[i + 1] = fmaf(a[i + 1], b1, b2); a[i + 2] = fmaf(a[i + 2], b2, b3); a[i + 3] = fmaf(a[i + 3], b3, b0);
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 24; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// accumulate form used by the real kernel: acc += n * w, n and w registers
__global__ void k_ffma_acc(float *out, float x, float y, long long *cyc)
{
    float a[24];
#pragma unroll
    for (int i = 0; i < 24; i++) a[i] = 0.f;
    float n[4] = {x, y, x + y, x - y}, w[6] = {y, x, y - x, x * y, x + 1, y + 1};
    long long t0 = clock64();
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < 24; i++) a[i] = fmaf(n[i & 3], w[i % 6], a[i]);
        n[0] += 1.0f;
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 24; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__device__ __forceinline__ void fma2(unsigned long long &d, unsigned long long a, unsigned long long b)
{
    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}

__global__ void k_ffma2(float *out, float x, float y, long long *cyc)
{
    unsigned long long a[24];
#pragma unroll
    for (int i = 0; i < 24; i++) a[i] = 0ull;
    unsigned long long n[4], w[6];
    for (int i = 0; i < 4; i++) n[i] = ((unsigned long long)__float_as_uint(x + i) << 32) | __float_as_uint(y + i);
    for (int i = 0; i < 6; i++) w[i] = ((unsigned long long)__float_as_uint(y * i) << 32) | __float_as_uint(x - i);
    long long t0 = clock64();
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < 24; i++) fma2(a[i], n[i & 3], w[i % 6]);
    }
    long long t1 = clock64();
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < 24; i++) s ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float((uint32_t)s) + __uint_as_float((uint32_t)(s >> 32));
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// the real inner-loop shape: per step `NL` LDS.128 of a W table + 8*NL FFMA (two sites per lane); mode selects the address pattern
//   0: all lanes the same address (thread <-> site, broadcast)      1: 8 consecutive 16-B words replicated over the 4 quarter-warps
//   2: 32 distinct consecutive 16-B words                           3: no LDS (register operands)
template <int MODE, int R2>
__global__ void k_mix(float *out, float x, long long *cyc)
{
    extern __shared__ float4 W[];   // [12][256]
    for (int i = threadIdx.x; i < 12 * 256; i += blockDim.x) W[i] = make_float4(x + i, x - i, x * i, x);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int off = MODE == 0 ? 0 : MODE == 1 ? (lane & 7) : lane;
    float acc[R2][12];
#pragma unroll
    for (int r = 0; r < R2; r++)
#pragma unroll
        for (int j = 0; j < 12; j++) acc[r][j] = 0.f;
    float4 n[R2];
#pragma unroll
    for (int r = 0; r < R2; r++) n[r] = make_float4(x + r, x - r, x * r, x + lane);
    long long t0 = clock64();
    for (int it = 0; it < ITER; it++) {
        const int s = (it * 8) & 255;
#pragma unroll
        for (int j = 0; j < 12; j++) {
            float4 w;
            if (MODE == 3) w = make_float4(n[0].x + j, n[0].y, n[0].z, n[0].w);
            else w = W[j * 256 + ((s + off) & 255)];
#pragma unroll
            for (int r = 0; r < R2; r++)
                acc[r][j] = fmaf(n[r].x, w.x, fmaf(n[r].y, w.y, fmaf(n[r].z, w.z, fmaf(n[r].w, w.w, acc[r][j]))));
        }
#pragma unroll
        for (int r = 0; r < R2; r++) n[r].x += 1.0f;
    }
    long long t1 = clock64();
    float sres = 0;
#pragma unroll
    for (int r = 0; r < R2; r++)
#pragma unroll
        for (int j = 0; j < 12; j++) sres += acc[r][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = sres;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}


// same shape with packed FFMA2 (pairs over the base index: (b0,b1) and (b2,b3)); partial sums are added at the end
template <int MODE, int R2>
__global__ void k_mix2(float *out, float x, long long *cyc)
{
    extern __shared__ float4 W[];
    for (int i = threadIdx.x; i < 12 * 256; i += blockDim.x) W[i] = make_float4(x + i, x - i, x * i, x);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int off = MODE == 0 ? 0 : MODE == 1 ? (lane & 7) : lane;
    unsigned long long acc[R2][12];
#pragma unroll
    for (int r = 0; r < R2; r++)
#pragma unroll
        for (int j = 0; j < 12; j++) acc[r][j] = 0ull;
    ulonglong2 n[R2];
#pragma unroll
    for (int r = 0; r < R2; r++) { n[r].x = ((unsigned long long)__float_as_uint(x + r) << 32) | __float_as_uint(x - lane); n[r].y = n[r].x + 12345; }
    long long t0 = clock64();
    for (int it = 0; it < ITER; it++) {
        const int s = (it * 8) & 255;
#pragma unroll
        for (int j = 0; j < 12; j++) {
            const ulonglong2 w = *reinterpret_cast<const ulonglong2 *>(&W[j * 256 + ((s + off) & 255)]);
#pragma unroll
            for (int r = 0; r < R2; r++) { fma2(acc[r][j], n[r].x, w.x); fma2(acc[r][j], n[r].y, w.y); }
        }
    }
    long long t1 = clock64();
    unsigned long long sres = 0;
#pragma unroll
    for (int r = 0; r < R2; r++)
#pragma unroll
        for (int j = 0; j < 12; j++) sres ^= acc[r][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float((uint32_t)sres) + __uint_as_float((uint32_t)(sres >> 32));
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// LDS.128 alone (sum into 4 accumulators so loads cannot be dropped)
template <int MODE>
__global__ void k_lds(float *out, float x, long long *cyc)
{
    extern __shared__ float4 W[];
    for (int i = threadIdx.x; i < 12 * 256; i += blockDim.x) W[i] = make_float4(x + i, x - i, x * i, x);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int off = MODE == 0 ? 0 : MODE == 1 ? (lane & 7) : lane;
    float4 a = make_float4(0, 0, 0, 0);
    long long t0 = clock64();
    for (int it = 0; it < ITER; it++) {
        const int s = (it * 8) & 255;
#pragma unroll
        for (int j = 0; j < 12; j++) {
            const float4 w = W[j * 256 + ((s + off) & 255)];
            a.x += w.x; a.y += w.y; a.z += w.z; a.w += w.w;
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a.x + a.y + a.z + a.w;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <typename F>
static void run(const char *name, F launch, int sms, int warps, double instr_per_iter_per_warp, double fma_equiv)
{
    long long *cyc; float *out;
    cudaMalloc(&cyc, sms * 4 * sizeof(long long)); cudaMalloc(&out, (size_t)sms * 4 * 1024 * sizeof(float));
    launch(out, cyc); launch(out, cyc);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a); launch(out, cyc); cudaEventRecord(b);
    cudaError_t e = cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    long long h[1024]; cudaMemcpy(h, cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
    double mc = 0; for (int i = 0; i < sms; i++) mc = mc > h[i] ? mc : (double)h[i];
    const double wi = instr_per_iter_per_warp * ITER * warps;
    printf("%-28s warps/SM %2d  cycles %9.0f  %.3f ms (%.0f MHz)  warp-instr/clk/SM %.3f  FMA-lanes/clk/SM %.1f  %s\n", name, warps, mc, ms,
           mc / ms / 1e3, wi / mc, fma_equiv * ITER * warps * 32 / mc, e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(cyc); cudaFree(out);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    printf("%s, %d SMs\n", p.name, sms);
    const int smem = 12 * 256 * 16;
    for (int warps : {4, 8, 16}) {
        const int th = warps * 32;
        run("ffma chain x24", [&](float *o, long long *c) { k_ffma<<<sms, th>>>(o, 1.0001f, 0.5f, c); }, sms, warps, 24, 24);
        run("ffma acc+=n*w x24", [&](float *o, long long *c) { k_ffma_acc<<<sms, th>>>(o, 1.0001f, 0.5f, c); }, sms, warps, 24, 24);
        run("ffma2 acc+=n*w x24", [&](float *o, long long *c) { k_ffma2<<<sms, th>>>(o, 1.0001f, 0.5f, c); }, sms, warps, 24, 48);
        run("lds128 uniform", [&](float *o, long long *c) { k_lds<0><<<sms, th, smem>>>(o, 1.0f, c); }, sms, warps, 12, 0);
        run("lds128 8x16B replicated x4", [&](float *o, long long *c) { k_lds<1><<<sms, th, smem>>>(o, 1.0f, c); }, sms, warps, 12, 0);
        run("lds128 32 distinct", [&](float *o, long long *c) { k_lds<2><<<sms, th, smem>>>(o, 1.0f, c); }, sms, warps, 12, 0);
        run("mix R2=2 regs only", [&](float *o, long long *c) { k_mix<3, 2><<<sms, th, smem>>>(o, 1.0f, c); }, sms, warps, 96, 96);
        run("mix R2=2 uniform", [&](float *o, long long *c) { k_mix<0, 2><<<sms, th, smem>>>(o, 1.0f, c); }, sms, warps, 108, 96);
        run("mix R2=2 replicated", [&](float *o, long long *c) { k_mix<1, 2><<<sms, th, smem>>>(o, 1.0f, c); }, sms, warps, 108, 96);
        run("mix R2=2 distinct", [&](float *o, long long *c) { k_mix<2, 2><<<sms, th, smem>>>(o, 1.0f, c); }, sms, warps, 108, 96);
        run("mix R2=4 uniform", [&](float *o, long long *c) { k_mix<0, 4><<<sms, th, smem>>>(o, 1.0f, c); }, sms, warps, 204, 192);
        run("mix R2=4 replicated", [&](float *o, long long *c) { k_mix<1, 4><<<sms, th, smem>>>(o, 1.0f, c); }, sms, warps, 204, 192);
        run("mix R2=4 distinct", [&](float *o, long long *c) { k_mix<2, 4><<<sms, th, smem>>>(o, 1.0f, c); }, sms, warps, 204, 192);
        run("mix2 (FFMA2) R2=2 uniform", [&](float *o, long long *c) { k_mix2<0, 2><<<sms, th, smem>>>(o, 1.0f, c); }, sms, warps, 60, 96);
        run("mix2 (FFMA2) R2=2 replicated", [&](float *o, long long *c) { k_mix2<1, 2><<<sms, th, smem>>>(o, 1.0f, c); }, sms, warps, 60, 96);
        run("mix2 (FFMA2) R2=4 replicated", [&](float *o, long long *c) { k_mix2<1, 4><<<sms, th, smem>>>(o, 1.0f, c); }, sms, warps, 108, 192);
        run("mix2 (FFMA2) R2=4 distinct", [&](float *o, long long *c) { k_mix2<2, 4><<<sms, th, smem>>>(o, 1.0f, c); }, sms, warps, 108, 192);
        run("mix R2=1 uniform", [&](float *o, long long *c) { k_mix<0, 1><<<sms, th, smem>>>(o, 1.0f, c); }, sms, warps, 60, 48);
        run("mix R2=1 replicated", [&](float *o, long long *c) { k_mix<1, 1><<<sms, th, smem>>>(o, 1.0f, c); }, sms, warps, 60, 48);
    }
    return 0;
}
