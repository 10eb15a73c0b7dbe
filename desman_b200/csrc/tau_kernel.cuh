// desman_b200/csrc/tau_kernel.cuh -- K1: the tau Gibbs update (replaces sampletau/c_sample_tau.c:95-204)
// fused with K4, the sum_n n*log(p) part of logLikelihood (HaploSNP_Sampler.py:431-442).
//
// Mapping: one warp per variant position v (sites are independent, c_sample_tau.c:130), lanes over
// samples s, strains g strictly in order (each draw conditions on the tau just written, :133,:180).
// The site's S count cells (one 128-bit int32x4 word per (v,s)) are staged once into the warp's
// shared-memory tile and re-read for every strain.
#pragma once
#include "common.cuh"

struct TauParams {
    const int4 *counts;      // [V][S] int32x4
    uint8_t *tau;            // [V][G] base index, updated in place
    const double *gamma;     // [S][G]
    const double *eta;       // [16] eta used for the draw (row = true base)
    const double *eta_ll;    // [16] eta used for the log-likelihood term, or nullptr (no ll)
    const uint32_t *words;   // MT19937 words [V*G] (u = w/2^32), or nullptr -> Philox
    uint64_t seed;
    uint32_t sweep;
    int64_t v0;              // global index of local site 0 (Philox counter / sharding)
    int V, S, G;
    unsigned long long *nchange;  // += flips
    double *ll_partial;      // [gridDim.x] per-block sum of n*log p, or nullptr
    uint32_t *tau_cnt;       // [V][G][4] lazy per-base occupancy counters, or nullptr
    uint32_t *tau_last;      // [V][G] iteration at which the current base was adopted
    uint32_t iter;           // iteration index inside the current update() call
    int do_draw;             // 0: skip the Gibbs draw, only accumulate the log-likelihood term
};

#define TAU_WARPS 8

// Reference arithmetic (c_sample_tau.c:136-176) in FP64: base over h ascending skipping g,
// candidate term added last, count through float, softmax with max subtraction, strict '<' CDF.
// Terms with n == 0 are skipped: 0*log(p) adds exactly -0.0 there (p > 0 because eta, gamma > 0).
__global__ void __launch_bounds__(TAU_WARPS * 32) tau_sample_kernel(TauParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int S = p.S, G = p.G;
    const int Sp = (S + 31) & ~31;
    double *gT = reinterpret_cast<double *>(smem_raw);           // [G][Sp] gamma transposed
    double *eta_s = gT + (size_t)G * Sp;                         // [16]
    double *etall_s = eta_s + 16;                                // [16]
    int4 *tiles = reinterpret_cast<int4 *>(etall_s + 16);        // [TAU_WARPS][S]
    __shared__ double ll_warp[TAU_WARPS];

    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < G * Sp; i += blockDim.x) {
        int g = i / Sp, s = i - g * Sp;
        gT[i] = (s < S) ? p.gamma[(size_t)s * G + g] : 0.0;
    }
    if (threadIdx.x < 16) {
        eta_s[threadIdx.x] = p.eta[threadIdx.x];
        etall_s[threadIdx.x] = p.eta_ll ? p.eta_ll[threadIdx.x] : 0.0;
    }
    __syncthreads();

    int4 *tile = tiles + (size_t)wib * S;
    const int gw = blockIdx.x * TAU_WARPS + wib, nw = gridDim.x * TAU_WARPS;
    const uint32_t k0 = (uint32_t)p.seed, k1 = (uint32_t)(p.seed >> 32);
    unsigned int flips = 0;
    double ll_acc = 0.0;

    for (int v = gw; v < p.V; v += nw) {
        const int4 *src = p.counts + (size_t)v * S;
        for (int s = lane; s < S; s += 32) tile[s] = ld_counts(src + s);
        uint64_t code = load_tau_code(p.tau + (size_t)v * G, G, lane);
        const uint64_t code_in = code;
        __syncwarp();

        for (int g = 0; g < (p.do_draw ? G : 0); g++) {
            const int cur = code_get(code, g);
            double L0 = 0.0, L1 = 0.0, L2 = 0.0, L3 = 0.0;
            for (int s = lane; s < S; s += 32) {
                const int4 n = tile[s];
                if ((n.x | n.y | n.z | n.w) == 0) continue;
                double b0 = 0.0, b1 = 0.0, b2 = 0.0, b3 = 0.0;
                for (int h = 0; h < G; h++) {
                    if (h == g) continue;
                    const double *e = eta_s + 4 * code_get(code, h);
                    const double gm = gT[h * Sp + s];
                    b0 = fma(e[0], gm, b0); b1 = fma(e[1], gm, b1);
                    b2 = fma(e[2], gm, b2); b3 = fma(e[3], gm, b3);
                }
                const double gg = gT[g * Sp + s];
                const double f0 = (double)(float)n.x, f1 = (double)(float)n.y,
                             f2 = (double)(float)n.z, f3 = (double)(float)n.w;
#define TAU_CAND(a, L)                                                        \
    {                                                                         \
        const double *e = eta_s + 4 * (a);                                    \
        if (n.x) L = fma(f0, log(fma(e[0], gg, b0)), L);                      \
        if (n.y) L = fma(f1, log(fma(e[1], gg, b1)), L);                      \
        if (n.z) L = fma(f2, log(fma(e[2], gg, b2)), L);                      \
        if (n.w) L = fma(f3, log(fma(e[3], gg, b3)), L);                      \
    }
                TAU_CAND(0, L0) TAU_CAND(1, L1) TAU_CAND(2, L2) TAU_CAND(3, L3)
#undef TAU_CAND
            }
            L0 = warp_sum(L0); L1 = warp_sum(L1); L2 = warp_sum(L2); L3 = warp_sum(L3);
            // normaliseLog4 (c_sample_tau.c:48-70)
            double mx = L0;
            if (L1 > mx) mx = L1;
            if (L2 > mx) mx = L2;
            if (L3 > mx) mx = L3;
            const double e0 = exp(L0 - mx), e1 = exp(L1 - mx), e2 = exp(L2 - mx), e3 = exp(L3 - mx);
            const double sum = ((0.0 + e0) + e1) + e2 + e3;
            const double p0 = e0 / sum, p1 = e1 / sum, p2 = e2 / sum;
            uint32_t w;
            if (p.words) w = p.words[(size_t)v * G + g];
            else w = philox4x32_10((uint32_t)(p.v0 + v), (uint32_t)g, p.sweep, (uint32_t)STAGE_TAU << 28, k0, k1).x;
            const double u = (double)w / 4294967296.0;              // gsl_rng_uniform, :174
            // sample4 (c_sample_tau.c:72-91)
            const double c0 = p0, c1 = p1 + c0, c2 = p2 + c1;
            const int t = (u < c0) ? 0 : (u < c1) ? 1 : (u < c2) ? 2 : 3;
            if (t != cur) {
                code = code_set(code, g, t);
                flips++;
                if (p.tau_cnt && lane == 0) {
                    const size_t vg = (size_t)v * G + g;
                    p.tau_cnt[vg * 4 + cur] += p.iter - p.tau_last[vg];
                    p.tau_last[vg] = p.iter;
                }
            }
        }
        if (code != code_in && lane < G) p.tau[(size_t)v * G + lane] = (uint8_t)code_get(code, lane);

        if (p.ll_partial) {
            // sum_s sum_b n*log(p_vsb), p = sum_g gamma[s,g]*eta_ll[tau_vg,b]   (HaploSNP_Sampler.py:435,441)
            double acc = 0.0;
            for (int s = lane; s < S; s += 32) {
                const int4 n = tile[s];
                if ((n.x | n.y | n.z | n.w) == 0) continue;
                double b0 = 0.0, b1 = 0.0, b2 = 0.0, b3 = 0.0;
                for (int h = 0; h < G; h++) {
                    const double *e = etall_s + 4 * code_get(code, h);
                    const double gm = gT[h * Sp + s];
                    b0 = fma(e[0], gm, b0); b1 = fma(e[1], gm, b1);
                    b2 = fma(e[2], gm, b2); b3 = fma(e[3], gm, b3);
                }
                if (n.x) acc = fma((double)n.x, log(b0), acc);
                if (n.y) acc = fma((double)n.y, log(b1), acc);
                if (n.z) acc = fma((double)n.z, log(b2), acc);
                if (n.w) acc = fma((double)n.w, log(b3), acc);
            }
            ll_acc += warp_sum(acc);
        }
        __syncwarp();
    }

    if (lane == 0 && flips) atomicAdd(p.nchange, (unsigned long long)flips);
    if (p.ll_partial) {
        if (lane == 0) ll_warp[wib] = ll_acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int i = 0; i < TAU_WARPS; i++) t += ll_warp[i];
            p.ll_partial[blockIdx.x] = t;   // fixed site->warp->block order: deterministic
        }
    }
}

static inline size_t tau_smem_bytes(int S, int G)
{
    size_t Sp = (size_t)((S + 31) & ~31);
    return sizeof(double) * ((size_t)G * Sp + 32) + sizeof(int4) * (size_t)TAU_WARPS * S;
}
