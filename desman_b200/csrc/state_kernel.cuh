// desman_b200/csrc/state_kernel.cuh -- joint-state enumeration: for N sites and all T = 4^G joint
// haplotype states, stateLogProb[n][t] = sum_{s,b} n_sb * log(sum_g gamma[s,g] * eta[tau_t[g], b])
// (HaploSNP_Sampler.py:233-261 assignTau, :498-524 logTauProb), and the likelihood under a real-valued (non-one-hot) tau
// (:431-442 as called by DIC, :486-496).  FP64 throughout: the callers exponentiate differences of sums of ~1e4 nats.
//
// The table is a dense product counts[N, 4S] x logSite[4S, T]: K10 builds logSite^T (k-major, so that consecutive states are
// consecutive addresses), K11 is a shared-memory tiled FP64 product (64 sites x 64 states per CTA, 4x4 per thread) whose
// epilogue leaves, per (site, block of 64 states), the partial (max, sum exp, argmax) of an online log-sum-exp, the value at a
// requested state and -- only if asked for -- the full row; K12 folds the partials per site.  State numbering is the
// reference's: strain g is digit G-1-g of t in base 4 (tauMap, :112-116; Desman_Utils.cartesian order, :95-103).
#pragma once
#include "common.cuh"

#define ST_BM 64
#define ST_BN 64
#define ST_BK 16

// logSiteT[k][t], k = s*4 + b
__global__ void state_logsite_kernel(const double *__restrict__ gamma, const double *__restrict__ eta, int S, int G, long long T,
                                     double *__restrict__ logSiteT)
{
    __shared__ double eta_s[16];
    if (threadIdx.x < 16) eta_s[threadIdx.x] = eta[threadIdx.x];
    __syncthreads();
    const long long total = T * S;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long t = i % T;
        const int s = (int)(i / T);
        double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;
        for (int g = 0; g < G; g++) {
            const int a = (int)((t >> (2 * (G - 1 - g))) & 3);
            const double gm = gamma[(size_t)s * G + g];
            p0 = fma(gm, eta_s[4 * a + 0], p0); p1 = fma(gm, eta_s[4 * a + 1], p1);
            p2 = fma(gm, eta_s[4 * a + 2], p2); p3 = fma(gm, eta_s[4 * a + 3], p3);
        }
        logSiteT[(size_t)(4 * s + 0) * T + t] = log(p0);
        logSiteT[(size_t)(4 * s + 1) * T + t] = log(p1);
        logSiteT[(size_t)(4 * s + 2) * T + t] = log(p2);
        logSiteT[(size_t)(4 * s + 3) * T + t] = log(p3);
    }
}

// counts of the chunk as doubles, k-major: Cd[k][n] (n < Nc), from int64 [N][S][4] (host upload) or the engine's int4 cells
__global__ void state_counts_kernel(const long long *__restrict__ v64, const int4 *__restrict__ v4, long long n0, int Nc, int K,
                                    double *__restrict__ Cd)
{
    const long long total = (long long)Nc * K;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(i / K), k = (int)(i % K);
        double x;
        if (v64) x = (double)v64[(size_t)(n0 + n) * K + k];
        else {
            const int4 c = v4[((size_t)(n0 + n) * K + k) >> 2];
            x = (double)((k & 3) == 0 ? c.x : (k & 3) == 1 ? c.y : (k & 3) == 2 ? c.z : c.w);
        }
        Cd[(size_t)k * Nc + n] = x;
    }
}

struct StateParams {
    const double *Cd;         // [K][Nc]
    const double *logSiteT;   // [K][T]
    int Nc, K;
    long long T;
    int nblk;                 // ceil(T / ST_BN)
    double *part_max, *part_sum;   // [Nc][nblk]
    long long *part_arg;      // [Nc][nblk]
    const long long *index;   // [Nc] state whose log-probability is wanted, or nullptr
    double *lp_at_index;      // [Nc]
    double *logprob;          // [Nc][T] or nullptr
};

__global__ void __launch_bounds__(256) state_logprob_kernel(StateParams p)
{
    __shared__ double As[ST_BK][ST_BM + 1];     // counts tile   [k][site]
    __shared__ double Bs[ST_BK][ST_BN + 1];     // logSite tile  [k][state]
    __shared__ double Cs[ST_BM / 2][ST_BN + 1]; // result tile for the epilogue, half of the sites at a time (48 KB static limit)
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int n0 = blockIdx.y * ST_BM;
    const long long t0 = (long long)blockIdx.x * ST_BN;
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.0;
    for (int k0 = 0; k0 < p.K; k0 += ST_BK) {
        for (int i = threadIdx.x; i < ST_BK * ST_BM; i += 256) {
            const int kk = i / ST_BM, m = i % ST_BM;
            const int k = k0 + kk, n = n0 + m;
            As[kk][m] = (k < p.K && n < p.Nc) ? p.Cd[(size_t)k * p.Nc + n] : 0.0;
            const long long t = t0 + m;
            Bs[kk][m] = (k < p.K && t < p.T) ? p.logSiteT[(size_t)k * p.T + t] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < ST_BK; kk++) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; i++) { a[i] = As[kk][ty * 4 + i]; b[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    for (int half = 0; half < 2; half++) {
        const int m0 = half * (ST_BM / 2);
        if ((ty >> 3) == half) {
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) Cs[(ty & 7) * 4 + i][tx * 4 + j] = acc[i][j];
        }
        __syncthreads();
        // full rows, coalesced over the states
        if (p.logprob)
            for (int i = threadIdx.x; i < (ST_BM / 2) * ST_BN; i += 256) {
                const int m = i / ST_BN, j = i % ST_BN;
                if (n0 + m0 + m < p.Nc && t0 + j < p.T) p.logprob[(size_t)(n0 + m0 + m) * p.T + t0 + j] = Cs[m][j];
            }
        // partial log-sum-exp of the 64 states of this block, one thread per site (first maximum wins, like numpy.argmax)
        if (threadIdx.x < ST_BM / 2 && n0 + m0 + (int)threadIdx.x < p.Nc) {
            const int m = threadIdx.x, n = n0 + m0 + m;
            const int nj = (int)((p.T - t0 < ST_BN) ? p.T - t0 : ST_BN);
            double mx = Cs[m][0];
            int arg = 0;
            for (int j = 1; j < nj; j++) if (Cs[m][j] > mx) { mx = Cs[m][j]; arg = j; }
            double sum = 0.0;
            for (int j = 0; j < nj; j++) sum += exp(Cs[m][j] - mx);
            const size_t o = (size_t)n * p.nblk + blockIdx.x;
            p.part_max[o] = mx; p.part_sum[o] = sum; p.part_arg[o] = t0 + arg;
            if (p.index) {
                const long long want = p.index[n];
                if (want >= t0 && want < t0 + nj) p.lp_at_index[n] = Cs[m][(int)(want - t0)];
            }
        }
        __syncthreads();
    }
}

// per site: maximum, its state, log sum exp over all states
__global__ void state_reduce_kernel(const double *__restrict__ part_max, const double *__restrict__ part_sum,
                                    const long long *__restrict__ part_arg, int Nc, int nblk, double *__restrict__ maxlp,
                                    double *__restrict__ lse, long long *__restrict__ argmax)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= Nc) return;
    const size_t o = (size_t)n * nblk;
    double mx = part_max[o];
    long long arg = part_arg[o];
    for (int b = 1; b < nblk; b++) if (part_max[o + b] > mx) { mx = part_max[o + b]; arg = part_arg[o + b]; }
    double sum = 0.0;
    for (int b = 0; b < nblk; b++) sum += part_sum[o + b] * exp(part_max[o + b] - mx);
    maxlp[n] = mx; lse[n] = mx + log(sum); argmax[n] = arg;
}

// log-likelihood terms sum n*log p under a real-valued tau [V][G][4] (HaploSNP_Sampler.py:435,:441 with the tauMean of :479-484):
// p_vsb = sum_g gamma[s,g] * sum_a tau[v,g,a] * eta[a,b].  One warp per site, lanes over samples; per-block partial sums,
// folded on the host in block order (fixed order => reproducible).
__global__ void __launch_bounds__(256) loglik_general_kernel(const int4 *__restrict__ counts, const double *__restrict__ tau,
                                                             const double *__restrict__ gamma, const double *__restrict__ eta,
                                                             int V, int S, int G, double *__restrict__ partial)
{
    __shared__ double eta_s[16];
    __shared__ double red[8];
    if (threadIdx.x < 16) eta_s[threadIdx.x] = eta[threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    double acc = 0.0;
    for (int v = blockIdx.x * 8 + wib; v < V; v += gridDim.x * 8) {
        for (int s = lane; s < S; s += 32) {
            const int4 n = counts[(size_t)v * S + s];
            double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;
            for (int g = 0; g < G; g++) {
                const double gm = gamma[(size_t)s * G + g];
                const double *t = tau + ((size_t)v * G + g) * 4;
                double w0 = 0.0, w1 = 0.0, w2 = 0.0, w3 = 0.0;
#pragma unroll
                for (int a = 0; a < 4; a++) {
                    w0 = fma(t[a], eta_s[4 * a + 0], w0); w1 = fma(t[a], eta_s[4 * a + 1], w1);
                    w2 = fma(t[a], eta_s[4 * a + 2], w2); w3 = fma(t[a], eta_s[4 * a + 3], w3);
                }
                p0 = fma(gm, w0, p0); p1 = fma(gm, w1, p1); p2 = fma(gm, w2, p2); p3 = fma(gm, w3, p3);
            }
            if (n.x) acc = fma((double)n.x, log(p0), acc);
            if (n.y) acc = fma((double)n.y, log(p1), acc);
            if (n.z) acc = fma((double)n.z, log(p2), acc);
            if (n.w) acc = fma((double)n.w, log(p3), acc);
        }
    }
    acc = warp_sum(acc);
    if (lane == 0) red[wib] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < 8; i++) t += red[i];
        partial[blockIdx.x] = t;
    }
}
