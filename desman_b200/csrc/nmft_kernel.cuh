// desman_b200/csrc/nmft_kernel.cuh -- K6-K8: the NMFT initialiser (replaces Init_NMFT.py:98-205).
//
// KL-divergence multiplicative updates of X[4V,S] ~ tau[4V,G] * gamma[G,S] with per-(v,g) simplex
// renormalisation.  The reference does four np.dot products and a Python V*G loop per iteration
// (Init_NMFT.py:158-181).  Here one iteration is three launches, all bandwidth-bound on X:
//   nmft_gamma_kernel  (small)  reduce the per-block partials of the previous pass in fixed order,
//                               evaluate the stop rule |div_prev - div| > min_change on the device,
//                               apply the gamma update (:161-166) and the eps clamp (:88-91)
//   nmft_tau_kernel    (site)   tau update (:170-181): one warp per site, lanes over samples
//   nmft_stats_kernel  (site)   objective (:152-156) + numerators tau^T (X / (tau gamma)) and column
//                               sums for the NEXT gamma update; lane <-> sample fixed per warp
// Every reduction has a fixed order: results are bitwise reproducible run to run.
// Device layouts: X[v][a][s], tau[v][a][g], gamma[g][s]; the C-ABI keeps the reference layouts
// (rows v + a*V, Init_NMFT.py:58-60).
#pragma once
#include "common.cuh"
#include <float.h>
#include <stdio.h>
#include <vector>

#define NMFT_EPS DBL_EPSILON   // np.finfo(float64).eps (Desman_Utils.py:16-17)
#define NMFT_WARPS 8

__device__ __forceinline__ double nzd(double x) { return x == 0.0 ? NMFT_EPS : x; }   // du.elop zero rule

struct NmftState {   // device-resident loop state (Init_NMFT.py:103-115)
    double div, divl;
    int iter, done;
};

// X[v][a][s] = (n_vsa + 1) / sum_b (n_vsb + 1)    (Init_NMFT.py:49-60)
__global__ void nmft_freq_kernel(const long long *__restrict__ snps, double *__restrict__ X, int V, int S)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < (size_t)V * S; i += (size_t)gridDim.x * blockDim.x) {
        const size_t v = i / S, s = i - v * S;
        const long long *n = snps + i * 4;
        const double x0 = (double)n[0] + 1.0, x1 = (double)n[1] + 1.0, x2 = (double)n[2] + 1.0, x3 = (double)n[3] + 1.0;
        const double tot = ((x0 + x1) + x2) + x3;
        double *o = X + v * 4 * S + s;
        o[0] = x0 / tot; o[(size_t)S] = x1 / tot; o[2 * (size_t)S] = x2 / tot; o[3 * (size_t)S] = x3 / tot;
    }
}

struct NmftParams {
    const double *X;        // [V][4][S]
    double *tau;            // [V][4][G]
    double *gamma;          // [G][S]   gamma' (normalised, not clamped) -- used by the tau update
    double *gamma_adj;      // [G][S]   max(gamma', eps) -- used by objective / next gamma update
    double *t1;             // [G]      row sums of gamma'
    double *partial;        // [nblocks][G*S + G + 1]  num | H1 | div
    int nblocks;
    NmftState *st_in, *st_out;
    double *trace;
    int V, S, G;
    int max_iter, fix_gamma;   // fix_gamma: 0 factorize() (both factors, eps clamps), 1 factorize_tau() (tau only), 2 factorize_gamma() (gamma only)
    double min_change;
};

// grid = ceil(S/4) blocks of 256 threads; block b owns sample columns [4b, 4b+4).  Warp <-> strain (8 at a time), lane =
// (sample of the block) x (8 partial-sum lanes striding over the per-block partials of the stats pass): every (g,s) sum over
// the ~300 partials is 37 pipelined loads and a fixed-order 3-level shuffle tree instead of one thread's 300 dependent loads.
#define NMFT_GS 4
__global__ void __launch_bounds__(256) nmft_gamma_kernel(NmftParams p)
{
    __shared__ double red[256];
    __shared__ double col[32][NMFT_GS + 1];
    __shared__ int go;
    const int S = p.S, G = p.G, stride = G * S + G + 1;
    // total divergence of the previous pass (every block computes the same bits)
    double acc = 0.0;
    for (int b = threadIdx.x; b < p.nblocks; b += 256) acc += p.partial[(size_t)b * stride + G * S + G];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int m = 128; m > 0; m >>= 1) {
        if (threadIdx.x < m) red[threadIdx.x] += red[threadIdx.x + m];
        __syncthreads();
    }
    const double div = red[0];
    const NmftState st = *p.st_in;
    if (threadIdx.x == 0) {
        NmftState o = st;
        if (!st.done) {
            o.div = div;
            if (st.iter > 0 && p.trace) { if (blockIdx.x == 0) p.trace[st.iter - 1] = div; }
            // while iter < max_iter and |divl - div| > min_change   (Init_NMFT.py:106)
            const bool cont = st.iter < p.max_iter && fabs(st.divl - div) > p.min_change;
            if (cont) { o.divl = div; o.iter = st.iter + 1; } else o.done = 1;
        }
        go = !o.done;
        if (blockIdx.x == 0) *p.st_out = o;
    }
    __syncthreads();
    if (!go) return;
    const int lane = threadIdx.x & 31, gsub = threadIdx.x >> 5;       // 8 strain rows at a time
    const int s_local = lane >> 3, j = lane & 7;
    const int s = blockIdx.x * NMFT_GS + s_local;
    if (p.fix_gamma != 1) {
        for (int g0 = 0; g0 < G; g0 += 8) {
            const int g = g0 + gsub;
            double newg = 0.0;
            if (g < G) {                                                  // (warp-uniform)
                double num = 0.0, h1 = 0.0;
                if (G > 1) {
                    for (int b = j; b < p.nblocks; b += 8) {
                        if (s < S) num += p.partial[(size_t)b * stride + g * S + s];
                        h1 += p.partial[(size_t)b * stride + G * S + g];
                    }
#pragma unroll
                    for (int m = 4; m > 0; m >>= 1) {
                        num += __shfl_xor_sync(DESMAN_FULL_MASK, num, m);
                        h1 += __shfl_xor_sync(DESMAN_FULL_MASK, h1, m);
                    }
                    if (s < S) newg = p.gamma_adj[g * S + s] * (nzd(num) / nzd(h1));      // :163
                } else newg = 1.0;                                                        // :167-168
                if (j == 0) col[g][s_local] = newg;
            }
        }
        __syncthreads();
        if (threadIdx.x < NMFT_GS && blockIdx.x * NMFT_GS + (int)threadIdx.x < S) {
            const int sl = threadIdx.x, ss = blockIdx.x * NMFT_GS + sl;
            double cs = 0.0;
            for (int g = 0; g < G; g++) cs += col[g][sl];                                 // :165
            for (int g = 0; g < G; g++) {
                const double x = (G > 1) ? col[g][sl] / cs : 1.0;                         // :166
                p.gamma[g * S + ss] = x;
                p.gamma_adj[g * S + ss] = p.fix_gamma ? x : fmax(x, NMFT_EPS);            // _adjustment :88-91 (factorize() only: :126 is commented out)
            }
        }
    }
}

// tau update (Init_NMFT.py:170-181) + clamp (:88-91 when gamma is being fitted).  One warp per site, three phases:
//   1  lanes over samples:  r[a][s] = (X / (tau gamma))[a][s], zeros -> eps in both operands (du.elop)
//   2  lanes over the 4G outputs (a,g): numT[a][g] = sum_s r[a][s] gamma[g][s] (:172) as a sequential dot product per lane -- no
//      cross-lane reduction at all -- and the multiplicative update tau * numT / t1 with its division in parallel
//   3  lanes over strains: renormalisation over the 4 bases (:176-181)
// t1[g] = sum_s gamma'[g][s] (:170) is formed by every CTA in its prologue in a fixed order (it was a launch of its own).
// Shared-memory rows are padded by 2 doubles so that lanes reading different rows at the same sample hit different banks.
__global__ void __launch_bounds__(NMFT_WARPS * 32) nmft_tau_kernel(NmftParams p)
{
    if (p.st_out->done || p.fix_gamma == 2) return;         // (factorize_gamma, Init_NMFT.py:117-132: tau stays)
    extern __shared__ double sm[];
    const int S = p.S, G = p.G, Sr = S + 2;
    double *gm = sm;                                   // [G][Sr] gamma'
    double *t1 = gm + (size_t)G * Sr;                  // [G]
    double *tw = t1 + G;                               // [NMFT_WARPS][8][G]: old tau rows [4][G], new rows [4][G]
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < G * S; i += blockDim.x) { const int g = i / S, ss = i - g * S; gm[g * Sr + ss] = p.gamma[i]; }
    __syncthreads();
    for (int g = wib; g < G; g += NMFT_WARPS) {
        double acc = 0.0;
        for (int ss = lane; ss < S; ss += 32) acc += gm[g * Sr + ss];
        acc = warp_sum(acc);
        if (lane == 0) t1[g] = acc;
    }
    __syncthreads();
    double *told = tw + (size_t)wib * 8 * G, *tnew = told + 4 * G;
    double *rb = tw + (size_t)NMFT_WARPS * 8 * G + (size_t)wib * 4 * Sr;   // [4][Sr] per warp
    const int gw = blockIdx.x * NMFT_WARPS + wib, nw = gridDim.x * NMFT_WARPS;
    for (int v = gw; v < p.V; v += nw) {
        double *tv = p.tau + (size_t)v * 4 * G;
        for (int i = lane; i < 4 * G; i += 32) told[i] = tv[i];
        __syncwarp();
        const double *Xv = p.X + (size_t)v * 4 * S;
        for (int ss = lane; ss < S; ss += 32) {
#pragma unroll
            for (int a = 0; a < 4; a++) {
                double pa = 0.0;
                for (int h = 0; h < G; h++) pa = fma(told[a * G + h], gm[h * Sr + ss], pa);
                rb[a * Sr + ss] = nzd(Xv[a * S + ss]) / nzd(pa);
            }
        }
        __syncwarp();
        for (int o = lane; o < 4 * G; o += 32) {
            const int a = o / G, g = o - a * G;
            const double *ra = rb + a * Sr, *gg = gm + g * Sr;
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;                                // four chains: DFMA latency, not rate
            int ss = 0;
            for (; ss + 4 <= S; ss += 4) {
                a0 = fma(ra[ss], gg[ss], a0); a1 = fma(ra[ss + 1], gg[ss + 1], a1);
                a2 = fma(ra[ss + 2], gg[ss + 2], a2); a3 = fma(ra[ss + 3], gg[ss + 3], a3);
            }
            for (; ss < S; ss++) a0 = fma(ra[ss], gg[ss], a0);
            const double acc = (a0 + a1) + (a2 + a3);
            tnew[o] = told[o] * (nzd(acc) / nzd(t1[g]));
        }
        __syncwarp();
        for (int g = lane; g < G; g += 32) {
            double sumvg = 0.0;
            for (int a = 0; a < 4; a++) sumvg += tnew[a * G + g];                          // :176-178
            for (int a = 0; a < 4; a++) {
                double x = tnew[a * G + g] / sumvg;                                       // :180-181
                if (!p.fix_gamma) x = fmax(x, NMFT_EPS);                                  // :88-91 (factorize only)
                tv[a * G + g] = x;
            }
        }
        __syncwarp();
    }
}

// objective (:152-156) and, for the next gamma update, num[g][s] = sum_n tau[n][g] X/(tau gamma) (:163),
// H1[g] = sum_n tau[n][g] (:160).  Warp w owns sample chunk (w mod nch); lane <-> s.
template <int GP>
__global__ void __launch_bounds__(NMFT_WARPS * 32) nmft_stats_kernel(NmftParams p)
{
    if (p.st_out->done) return;
    extern __shared__ double sm[];
    const int S = p.S, G = p.G;
    double *gm = sm;                                    // [G][S] gamma_adj
    double *numS = gm + (size_t)G * S;                  // [G][S] block partial
    double *h1S = numS + (size_t)G * S;                 // [G]
    double *tw = h1S + G;                               // [NMFT_WARPS][4][G]
    __shared__ double divS[NMFT_WARPS];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < G * S; i += blockDim.x) { gm[i] = p.gamma_adj[i]; numS[i] = 0.0; }
    for (int i = threadIdx.x; i < G; i += blockDim.x) h1S[i] = 0.0;
    __syncthreads();
    const int nch = (S + 31) >> 5;
    const int gw = blockIdx.x * NMFT_WARPS + wib, nw = gridDim.x * NMFT_WARPS;
    const int chunk = gw % nch, s = chunk * 32 + lane;
    const bool valid = s < S;
    double *tl = tw + (size_t)wib * 4 * G;
    double gcol[GP], num[GP];
#pragma unroll
    for (int g = 0; g < GP; g++) { gcol[g] = (valid && g < G) ? gm[g * S + s] : 0.0; num[g] = 0.0; }
    double h1 = 0.0, dv = 0.0;
    for (int v = gw / nch; v < p.V; v += nw / nch) {
        const double *tv = p.tau + (size_t)v * 4 * G;
        for (int i = lane; i < 4 * G; i += 32) tl[i] = tv[i];
        __syncwarp();
        if (chunk == 0 && lane < G) h1 += ((tl[lane] + tl[G + lane]) + tl[2 * G + lane]) + tl[3 * G + lane];
        if (valid) {
#pragma unroll
            for (int a = 0; a < 4; a++) {
                const double x = p.X[((size_t)v * 4 + a) * S + s];
                double pa = 0.0;
#pragma unroll
                for (int g = 0; g < GP; g++) if (g < G) pa = fma(tl[a * G + g], gcol[g], pa);
                const double pc = fmax(pa, NMFT_EPS);                                     // _adjustment_input :93-97
                const double q = nzd(x) / nzd(pc);
                dv += x * log(q) - x + pc;                                                // :156
                const double r = (pc == pa) ? q : nzd(x) / nzd(pa);                       // same quotient unless pa was clamped
#pragma unroll
                for (int g = 0; g < GP; g++) if (g < G) num[g] = fma(tl[a * G + g], r, num[g]);
            }
        }
        __syncwarp();
    }
    dv = warp_sum(dv);
    if (lane == 0) divS[wib] = dv;
    // combine the warps of this block in warp order (fixed -> deterministic)
    for (int w = 0; w < NMFT_WARPS; w++) {
        if (wib == w) {
            if (valid) {
#pragma unroll
                for (int g = 0; g < GP; g++) if (g < G) numS[g * S + s] += num[g];
            }
            if (chunk == 0 && lane < G) h1S[lane] += h1;
        }
        __syncthreads();
    }
    const int stride = G * S + G + 1;
    double *out = p.partial + (size_t)blockIdx.x * stride;
    for (int i = threadIdx.x; i < G * S; i += blockDim.x) out[i] = numS[i];
    for (int i = threadIdx.x; i < G; i += blockDim.x) out[G * S + i] = h1S[i];
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < NMFT_WARPS; w++) t += divS[w];
        out[G * S + G] = t;
    }
}

#define NMFT_CU(call)                                                                              \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            snprintf(err, errn, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e_)); \
            rc = -2;                                                                               \
            goto done;                                                                             \
        }                                                                                          \
    } while (0)

template <int GP>
static void nmft_launch_stats(const NmftParams &p, int grid, size_t smem, cudaStream_t st)
{
    cudaFuncSetAttribute(nmft_stats_kernel<GP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    nmft_stats_kernel<GP><<<grid, NMFT_WARPS * 32, smem, st>>>(p);
}

// device time of the iteration loop of the last factorisation (CUDA events on its stream) and the iterations it ran
static double g_nmft_last_ms = 0.0;
static int g_nmft_last_iters = 0;

static int nmft_factorize_impl(cudaStream_t stream, int sm_count, const int64_t *snps, int64_t V, int S, int G, double *tau,
                               double *gamma, int max_iter, double min_change, int fix_gamma, int *n_iter_done,
                               double *div_final, double *div_trace, char *err, size_t errn)
{
    int rc = 0;
    const size_t nX = (size_t)V * 4 * S, nT = (size_t)V * 4 * G, nG = (size_t)G * S;
    const int nch = (S + 31) / 32;
    int grid_stats = sm_count * 2;
    grid_stats = ((grid_stats * NMFT_WARPS + nch - 1) / nch * nch + NMFT_WARPS - 1) / NMFT_WARPS;
    while ((grid_stats * NMFT_WARPS) % nch) grid_stats++;
    int grid_tau = sm_count * 2;        // one resident wave (set from the occupancy below)
    const size_t stride = nG + G + 1;
    const size_t smem_tau = sizeof(double) * ((size_t)G * (S + 2) + G + (size_t)NMFT_WARPS * 8 * G + (size_t)NMFT_WARPS * 4 * (S + 2));
    const size_t smem_stats = sizeof(double) * (2 * nG + G + (size_t)NMFT_WARPS * 4 * G);
    double *dX = nullptr, *dT = nullptr, *dG = nullptr, *dGa = nullptr, *dt1 = nullptr, *dP = nullptr, *dTr = nullptr;
    long long *dS = nullptr;
    NmftState *dSt = nullptr;
    std::vector<double> hT(nT), hG(nG);
    NmftState hst;
    NmftParams p;
    int it_guard = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (smem_tau > 220 * 1024 || smem_stats > 220 * 1024) { snprintf(err, errn, "NMFT: G*S too large for shared memory"); return -1; }

    NMFT_CU(cudaMalloc(&dX, nX * 8));
    NMFT_CU(cudaMalloc(&dT, nT * 8));
    NMFT_CU(cudaMalloc(&dG, nG * 8));
    NMFT_CU(cudaMalloc(&dGa, nG * 8));
    NMFT_CU(cudaMalloc(&dt1, G * 8));
    NMFT_CU(cudaMalloc(&dP, (size_t)grid_stats * stride * 8));
    NMFT_CU(cudaMalloc(&dSt, 2 * sizeof(NmftState)));
    NMFT_CU(cudaMalloc(&dTr, (size_t)(max_iter > 0 ? max_iter : 1) * 8));
    {   // counts -> X in bounded chunks
        const size_t chunk_v = ((size_t)4 << 20) / (size_t)S + 1;
        NMFT_CU(cudaMalloc(&dS, chunk_v * S * 4 * 8));
        for (size_t v0 = 0; v0 < (size_t)V; v0 += chunk_v) {
            const size_t nv = ((size_t)V - v0 < chunk_v) ? (size_t)V - v0 : chunk_v;
            NMFT_CU(cudaMemcpyAsync(dS, snps + v0 * S * 4, nv * S * 32, cudaMemcpyHostToDevice, stream));
            nmft_freq_kernel<<<sm_count * 4, 256, 0, stream>>>(dS, dX + v0 * 4 * S, (int)nv, S);
            NMFT_CU(cudaGetLastError());
            NMFT_CU(cudaStreamSynchronize(stream));
        }
    }
    // reference layouts -> device layouts; initial _adjustment (:101) only in factorize()
    for (int64_t v = 0; v < V; v++)
        for (int a = 0; a < 4; a++)
            for (int g = 0; g < G; g++) {
                double x = tau[((size_t)v + (size_t)a * V) * G + g];
                if (!fix_gamma && x < NMFT_EPS) x = NMFT_EPS;
                hT[((size_t)v * 4 + a) * G + g] = x;
            }
    for (size_t i = 0; i < nG; i++) hG[i] = (!fix_gamma && gamma[i] < NMFT_EPS) ? NMFT_EPS : gamma[i];
    NMFT_CU(cudaMemcpyAsync(dT, hT.data(), nT * 8, cudaMemcpyHostToDevice, stream));
    NMFT_CU(cudaMemcpyAsync(dG, hG.data(), nG * 8, cudaMemcpyHostToDevice, stream));
    NMFT_CU(cudaMemcpyAsync(dGa, hG.data(), nG * 8, cudaMemcpyHostToDevice, stream));
    hst.div = 0.0; hst.divl = 0.0; hst.iter = 0; hst.done = 0;
    NMFT_CU(cudaMemcpyAsync(dSt, &hst, sizeof(hst), cudaMemcpyHostToDevice, stream));
    NMFT_CU(cudaMemcpyAsync(dSt + 1, &hst, sizeof(hst), cudaMemcpyHostToDevice, stream));

    p.X = dX; p.tau = dT; p.gamma = dG; p.gamma_adj = dGa; p.t1 = dt1; p.partial = dP; p.nblocks = grid_stats;
    p.trace = div_trace ? dTr : nullptr; p.V = (int)V; p.S = S; p.G = G; p.max_iter = max_iter; p.fix_gamma = fix_gamma;
    p.min_change = min_change;
    NMFT_CU(cudaFuncSetAttribute(nmft_tau_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tau));
    {
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, nmft_tau_kernel, NMFT_WARPS * 32, smem_tau) != cudaSuccess || occ < 1) occ = 1;
        grid_tau = sm_count * occ;
        if ((int64_t)grid_tau * NMFT_WARPS > V) grid_tau = (int)((V + NMFT_WARPS - 1) / NMFT_WARPS);
    }

#define NMFT_STATS()                                                                     \
    do {                                                                                 \
        if (G <= 4) nmft_launch_stats<4>(p, grid_stats, smem_stats, stream);             \
        else if (G <= 8) nmft_launch_stats<8>(p, grid_stats, smem_stats, stream);        \
        else if (G <= 16) nmft_launch_stats<16>(p, grid_stats, smem_stats, stream);      \
        else nmft_launch_stats<32>(p, grid_stats, smem_stats, stream);                   \
    } while (0)

    // pass 0: objective and numerators of the initial factors (div = div_objective(), :102-103)
    p.st_in = dSt; p.st_out = dSt;
    NMFT_STATS();
    NMFT_CU(cudaGetLastError());
    cudaEventCreate(&ev0); cudaEventCreate(&ev1);
    cudaEventRecord(ev0, stream);
    {
        int parity = 0;
        bool finished = false;
        while (!finished) {
            for (int b = 0; b < 64; b++) {   // enqueue a batch of iterations; the stop rule lives on the device
                p.st_in = dSt + parity; p.st_out = dSt + (parity ^ 1);
                nmft_gamma_kernel<<<(S + NMFT_GS - 1) / NMFT_GS, 256, 0, stream>>>(p);
                nmft_tau_kernel<<<grid_tau, NMFT_WARPS * 32, smem_tau, stream>>>(p);
                NMFT_STATS();
                parity ^= 1;
            }
            NMFT_CU(cudaGetLastError());
            NMFT_CU(cudaMemcpyAsync(&hst, dSt + parity, sizeof(hst), cudaMemcpyDeviceToHost, stream));
            NMFT_CU(cudaStreamSynchronize(stream));
            finished = hst.done != 0;
            if (++it_guard > (max_iter / 64) + 4) finished = true;
        }
    }
    cudaEventRecord(ev1, stream);
    NMFT_CU(cudaMemcpyAsync(hT.data(), dT, nT * 8, cudaMemcpyDeviceToHost, stream));
    NMFT_CU(cudaMemcpyAsync(hG.data(), fix_gamma == 1 ? dG : dGa, nG * 8, cudaMemcpyDeviceToHost, stream));
    if (div_trace && hst.iter > 0) NMFT_CU(cudaMemcpyAsync(div_trace, dTr, (size_t)hst.iter * 8, cudaMemcpyDeviceToHost, stream));
    NMFT_CU(cudaStreamSynchronize(stream));
    for (int64_t v = 0; v < V; v++)
        for (int a = 0; a < 4; a++)
            for (int g = 0; g < G; g++) tau[((size_t)v + (size_t)a * V) * G + g] = hT[((size_t)v * 4 + a) * G + g];
    for (size_t i = 0; i < nG; i++) gamma[i] = hG[i];
    if (n_iter_done) *n_iter_done = hst.iter;
    if (div_final) *div_final = hst.div;
    {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ev0, ev1) == cudaSuccess) { g_nmft_last_ms = ms; g_nmft_last_iters = hst.iter; }
    }
done:
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    for (void *q : {(void *)dX, (void *)dT, (void *)dG, (void *)dGa, (void *)dt1, (void *)dP, (void *)dTr, (void *)dS, (void *)dSt})
        if (q) cudaFree(q);
    return rc;
}
