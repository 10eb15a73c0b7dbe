// desman_b200/csrc/mu_kernel.cuh -- K2: mu/E sufficient statistics
// (replaces HaploSNP_Sampler.sampleMu, HaploSNP_Sampler.py:284-309, and the reductions at :266, :276).
//
// The reference materialises E[V,S,4,4] and mu[V,S,4,G] and then only consumes their sums
// sum_mu[S,G] and Esum[4,4].  With one-hot tau the two-stage multinomial (:301, :305-309) is a
// per-read categorical over strains with weights gamma[s,g]*eta[tau_vg,a]; the true base of the
// read is tau_vg of the chosen strain.  Draw contract (identical in oracle/desman_oracle.c):
//   read j of cell (v,s,a) uses word (j&3) of Philox(ctr=(v, j>>2, sweep, STAGE_MU<<28|a<<26|s), key=seed)
//   strain = #{g < G-1 : word >= T_g},  T_g = min(floor(cum_g * (2^32/cum_{G-1})), 2^32-1),
//   cum_g = sum_{h<=g} gamma[s,h]*eta[tau_vh,a] in ascending h, IEEE round-to-nearest, no FMA.
//
// Mapping: lane <-> sample s (fixed for the life of the warp, so gamma[s,:] and the sum_mu[s,:]
// accumulators live in registers), warps stride over sites.  Integer statistics are flushed with
// 64-bit global atomics: order-independent, hence bit-reproducible for any grid / GPU count.
#pragma once
#include "common.cuh"

struct MuParams {
    const int4 *counts;     // [V][S]
    const uint8_t *tau;     // [V][G]
    const double *gamma;    // [S][G]
    const double *eta;      // [16]
    uint64_t seed;
    uint32_t sweep;
    int64_t v0;
    int V, S, G;
    unsigned long long *sum_mu;  // [S][G] +=
    unsigned long long *esum;    // [16]   += (esum[a_obs*4 + b_true])
};

#define MU_WARPS 8
#define MU_FLUSH_SITES 64   // 64 sites * 2^24 max count < 2^31

// (x >= T) as the carry of x + (2^32 - T), evaluated by IMAD.WIDE on the FMA pipe: the kernel is bound by the
// integer-ALU pipe (Philox xors/adds + threshold counting), so the compares are moved off it.
__device__ __forceinline__ uint32_t ge_carry(uint32_t x, unsigned long long neg_t)
{
    unsigned long long r;
    asm("mad.wide.u32 %0, %1, 1, %2;" : "=l"(r) : "r"(x), "l"(neg_t));
    return (uint32_t)(r >> 32);
}

template <int GP>
__global__ void __launch_bounds__(MU_WARPS * 32, 2) mu_stats_kernel(MuParams p)
{
    pdl_enter();
    __shared__ double eta_s[16];
    const int S = p.S, G = p.G;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    if (threadIdx.x < 16) eta_s[threadIdx.x] = p.eta[threadIdx.x];
    __syncthreads();

    const int nch = (S + 31) >> 5;                       // sample chunks of 32
    const int gw = blockIdx.x * MU_WARPS + wib, nw = gridDim.x * MU_WARPS;  // nw % nch == 0 (host)
    const int chunk = gw % nch;
    const int s = chunk * 32 + lane;
    const bool valid = s < S;
    const int vstride = nw / nch;
    const uint32_t k0 = (uint32_t)p.seed, k1 = (uint32_t)(p.seed >> 32);

    double gam[GP];
    int acc[GP];
    int E[16];
#pragma unroll
    for (int g = 0; g < GP; g++) {
        gam[g] = (valid && g < G) ? p.gamma[(size_t)s * G + g] : 0.0;
        acc[g] = 0;
    }
#pragma unroll
    for (int i = 0; i < 16; i++) E[i] = 0;

    int since_flush = 0;
    for (int v = gw / nch; v < p.V; v += vstride) {
        const uint64_t code = load_tau_code(p.tau + (size_t)v * G, G, lane);
        int4 n = make_int4(0, 0, 0, 0);
        if (valid) n = ld_counts(p.counts + (size_t)v * S + s);
        const int nn[4] = {n.x, n.y, n.z, n.w};
#pragma unroll
        for (int a = 0; a < 4; a++) {
            const int na = nn[a];
            if (na <= 0) continue;
            // integer thresholds of the cumulative weights (deterministic IEEE sequence)
            double cums[GP];
            double cum = 0.0;
#pragma unroll
            for (int g = 0; g < GP; g++) {
                const double w = __dmul_rn(gam[g], eta_s[4 * code_get(code, g) + a]);
                cum = __dadd_rn(cum, w);
                cums[g] = cum;
            }
            const double scale = __ddiv_rn(4294967296.0, cum);
            unsigned long long nthr[GP - 1];   // 2^32 - T_g
            uint32_t cge[GP - 1];
#pragma unroll
            for (int g = 0; g < GP - 1; g++) {
                const double t = floor(__dmul_rn(cums[g], scale));
                const uint32_t thr = (t >= 4294967295.0) ? 0xffffffffu : (uint32_t)t;
                nthr[g] = 0x100000000ull - (unsigned long long)thr;
                cge[g] = 0;
            }
            const uint32_t c3 = ((uint32_t)STAGE_MU << 28) | ((uint32_t)a << 26) | (uint32_t)s;
            // full blocks of 4 reads, then one guarded tail block
            const int nfull = na >> 2;
            for (int jb = 0; jb < nfull; jb++) {
                const uint4 o = philox4x32_10((uint32_t)(p.v0 + v), (uint32_t)jb, p.sweep, c3, k0, k1);
#pragma unroll
                for (int g = 0; g < GP - 1; g++) {
                    cge[g] += ge_carry(o.x, nthr[g]) + ge_carry(o.y, nthr[g]);
                    cge[g] += ge_carry(o.z, nthr[g]) + ge_carry(o.w, nthr[g]);
                }
            }
            const int rem = na & 3;
            if (rem) {
                const uint4 o = philox4x32_10((uint32_t)(p.v0 + v), (uint32_t)nfull, p.sweep, c3, k0, k1);
#pragma unroll
                for (int g = 0; g < GP - 1; g++) {
                    cge[g] += ge_carry(o.x, nthr[g]);
                    cge[g] += (rem > 1) ? ge_carry(o.y, nthr[g]) : 0u;
                    cge[g] += (rem > 2) ? ge_carry(o.z, nthr[g]) : 0u;
                }
            }
            // cge[k] = #reads with strain >= k+1 (only k < G-1 is meaningful)
#pragma unroll
            for (int g = 0; g < GP; g++) {
                const uint32_t hi = (g == 0) ? (uint32_t)na : ((g - 1 < G - 1) ? cge[(g > 0) ? g - 1 : 0] : 0u);
                const uint32_t lo = (g < G - 1 && g < GP - 1) ? cge[(g < GP - 1) ? g : 0] : 0u;
                const int cnt = (int)(hi - lo);
                acc[g] += cnt;
                const int b = code_get(code, g);
                E[a * 4 + 0] += (b == 0) ? cnt : 0;
                E[a * 4 + 1] += (b == 1) ? cnt : 0;
                E[a * 4 + 2] += (b == 2) ? cnt : 0;
                E[a * 4 + 3] += (b == 3) ? cnt : 0;
            }
        }
        if (++since_flush == MU_FLUSH_SITES) {
            since_flush = 0;
            if (valid) {
#pragma unroll
                for (int g = 0; g < GP; g++)
                    if (g < G && acc[g]) { atomicAdd(p.sum_mu + (size_t)s * G + g, (unsigned long long)acc[g]); acc[g] = 0; }
            }
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const unsigned long long t = warp_sum_u64((unsigned long long)E[i]);
                if (lane == 0 && t) atomicAdd(p.esum + i, t);
                E[i] = 0;
            }
        }
    }
    if (valid) {
#pragma unroll
        for (int g = 0; g < GP; g++)
            if (g < G && acc[g]) atomicAdd(p.sum_mu + (size_t)s * G + g, (unsigned long long)acc[g]);
    }
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const unsigned long long t = warp_sum_u64((unsigned long long)E[i]);
        if (lane == 0 && t) atomicAdd(p.esum + i, t);
    }
}
