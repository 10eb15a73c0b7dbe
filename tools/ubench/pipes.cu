// tools/ubench/pipes.cu -- measured issue rates of the pipes the table build of the screening pass leans on (B200).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ubench/pipes tools/ubench/pipes.cu ; run under gpurun.
// For each op: W warps per SM, each running ITER iterations of 8 independent chains; reported as lanes per clock per SM, and
// the latency of ONE dependent chain (1 warp per SM).
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#define ITER 4096
enum { OP_FFMA, OP_DFMA, OP_F2F_DS, OP_LG2, OP_F2FP, OP_H2F, OP_SHFL, OP_DCHAIN, OP_LG2CHAIN, OP_F2FCHAIN, OP_COUNT };
const char *names[] = {"FFMA", "DFMA", "F2F.F32.F64", "MUFU.LG2", "F2FP.F16.F32.PACK", "HADD2.F32 (f16->f32)", "SHFL.IDX", "DFMA dependent chain", "MUFU.LG2 dependent chain", "F2F.F32.F64 + back chain"};

template <int OP>
__global__ void k(float *out, long long *cyc, float seed)
{
    float f[8];
    double d[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { f[i] = seed + threadIdx.x * 1e-3f + i; d[i] = (double)f[i]; }
    const long long t0 = clock64();
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (OP == OP_FFMA) f[i] = fmaf(f[i], 1.0000001f, 1e-7f);
            if (OP == OP_DFMA) d[i] = fma(d[i], 1.0000001, 1e-7);
            if (OP == OP_F2F_DS) { f[i] += (float)d[i]; d[i] = __longlong_as_double(__double_as_longlong(d[i]) + 1ll); }   // (+ FADD + 64-bit IADD)
            if (OP == OP_LG2) { float y; asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(f[i])); f[i] = y + 3.0f; }
            if (OP == OP_F2FP) { __half2 h = __floats2half2_rn(f[i], f[(i + 1) & 7]); f[i] += __low2float(h) * 1e-9f; }
            if (OP == OP_H2F) { __half2 h = *reinterpret_cast<__half2 *>(&f[i]); float2 g = __half22float2(h); f[i] = g.x + g.y; }
            if (OP == OP_SHFL) f[i] = __shfl_sync(0xffffffffu, f[i], (threadIdx.x + 1) & 31);
        }
        if (OP == OP_DCHAIN) d[0] = fma(d[0], 1.0000001, 1e-7);
        if (OP == OP_LG2CHAIN) { float y; asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(f[0])); f[0] = y; }
        if (OP == OP_F2FCHAIN) { f[0] = (float)d[0]; d[0] = (double)f[0]; }
    }
    const long long t1 = clock64();
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) acc += f[i] + (float)d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(int warps, float *out, long long *cyc, int sms)
{
    k<OP><<<sms, warps * 32>>>(out, cyc, 1.5f);
    cudaDeviceSynchronize();
    long long h[1024];
    cudaMemcpy(h, cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
    double c = 0;
    for (int i = 0; i < sms; i++) c += (double)h[i];
    c /= sms;
    const bool chain = OP >= OP_DCHAIN;
    const double ops = chain ? (double)ITER : (double)ITER * 8 * warps * 32;
    if (chain) printf("%-28s warps/SM %2d  cycles/op %.1f\n", names[OP], warps, c / ops);
    else printf("%-28s warps/SM %2d  lanes/clk/SM %.2f  (warp-instr/clk/SM %.3f)\n", names[OP], warps, ops / c, ops / c / 32);
}

int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float *out; long long *cyc;
    cudaMalloc(&out, sms * 1024 * sizeof(float));
    cudaMalloc(&cyc, 1024 * sizeof(long long));
    for (int w : {4, 16, 32}) {
        run<OP_FFMA>(w, out, cyc, sms); run<OP_DFMA>(w, out, cyc, sms); run<OP_F2F_DS>(w, out, cyc, sms); run<OP_LG2>(w, out, cyc, sms);
        run<OP_F2FP>(w, out, cyc, sms); run<OP_H2F>(w, out, cyc, sms); run<OP_SHFL>(w, out, cyc, sms);
    }
    run<OP_DCHAIN>(1, out, cyc, sms); run<OP_LG2CHAIN>(1, out, cyc, sms); run<OP_F2FCHAIN>(1, out, cyc, sms);
    printf("note: F2F.F32.F64 row issues one FADD and a 64-bit integer add per conversion as well; LG2 row one FADD; F2FP row one FFMA; HADD2 row one FADD\n");
    return 0;
}
