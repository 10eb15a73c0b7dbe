// desman_b200/csrc/misc_kernels.cuh -- small kernels around the two site passes:
//   lgamma_const       sum_vs lgamma(N+1) - sum_b lgamma(n_b+1)    (Desman_Utils.py:28-33, constant in the chain)
//   mt19937_kernel     K9: GSL-compatible MT19937 stream             (c_sample_tau.c:33-40,174)
//   draw_gamma_eta     K3: Dirichlet draws of gamma and eta          (HaploSNP_Sampler.py:263-281)
//   finalize_sweep / copy_tau_if / flush_tau_counts                  K5 (:326-332,:349-358,:444-461)
#pragma once
#include "common.cuh"

// ---------------------------------------------------------------------------------------------
__global__ void lgamma_const_kernel(const int4 *__restrict__ counts, size_t ncell, double *__restrict__ partial)
{
    __shared__ double red[256];
    double acc = 0.0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < ncell; i += (size_t)gridDim.x * blockDim.x) {
        const int4 n = counts[i];
        const double N = (double)n.x + (double)n.y + (double)n.z + (double)n.w;
        acc += lgamma(N + 1.0) - (lgamma((double)n.x + 1.0) + lgamma((double)n.y + 1.0) + lgamma((double)n.z + 1.0) +
                                  lgamma((double)n.w + 1.0));
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int m = 128; m > 0; m >>= 1) {
        if (threadIdx.x < m) red[threadIdx.x] += red[threadIdx.x + m];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

// ---------------------------------------------------------------------------------------------
// K9.  One block.  state[624] holds the current (already regenerated) block of the recurrence and
// `pos` how many of its words were consumed (GSL: mti).  Emits `n` tempered words of the stream,
// storing only those with index in [store_lo, store_hi) (a rank's slice under V-sharding) at
// out[index - store_lo].  Three dependent phases of <= 227 independent words per regeneration.
__device__ __forceinline__ uint32_t mt_twist(uint32_t a, uint32_t b)
{
    const uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
    return (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}

__global__ void __launch_bounds__(256) mt19937_kernel(uint32_t *__restrict__ state, int pos, size_t n,
                                                      size_t store_lo, size_t store_hi, uint32_t *__restrict__ out)
{
    __shared__ uint32_t bufA[624], bufB[624];
    uint32_t *cur = bufA, *nxt = bufB;
    const int t = threadIdx.x;
    for (int i = t; i < 624; i += 256) cur[i] = state[i];
    __syncthreads();
    size_t done = 0;
    while (done < n) {
        if (pos == 624) {
            if (t < 227) nxt[t] = cur[t + 397] ^ mt_twist(cur[t], cur[t + 1]);
            __syncthreads();
            if (t < 227) nxt[t + 227] = nxt[t] ^ mt_twist(cur[t + 227], cur[t + 228]);
            __syncthreads();
            if (t < 169) nxt[t + 454] = nxt[t + 227] ^ mt_twist(cur[t + 454], cur[t + 455]);
            if (t == 255) nxt[623] = nxt[396] ^ mt_twist(cur[623], nxt[0]);
            __syncthreads();
            uint32_t *tmp = cur; cur = nxt; nxt = tmp;
            pos = 0;
        }
        const size_t take = min((size_t)(624 - pos), n - done);
        for (int i = t; i < (int)take; i += 256) {
            const size_t idx = done + i;
            if (idx >= store_lo && idx < store_hi) {
                uint32_t y = cur[pos + i];
                y ^= y >> 11;
                y ^= (y << 7) & 0x9d2c5680u;
                y ^= (y << 15) & 0xefc60000u;
                y ^= y >> 18;
                out[idx - store_lo] = y;
            }
        }
        pos += (int)take;
        done += take;
    }
    __syncthreads();
    for (int i = t; i < 624; i += 256) state[i] = cur[i];
}

// ---------------------------------------------------------------------------------------------
// K3.  Marsaglia-Tsang gamma variates under the Philox contract (see oracle_gamma_variate):
// attempt t of variate idx owns block ctr=(idx, t, sweep, stage<<28): (w0,w1)->53-bit radius uniform,
// w2 -> angle, w3 -> accept; shape < 1 boosts with U^(1/shape), U from ctr=(idx,0,sweep,boost<<28).
__device__ __forceinline__ double u53(uint32_t hi, uint32_t lo)
{
    const unsigned long long m = ((unsigned long long)(hi >> 5) << 26) | (unsigned long long)(lo >> 6);
    return ((double)m + 0.5) * (1.0 / 9007199254740992.0);
}
__device__ __forceinline__ double u32d(uint32_t w) { return ((double)w + 0.5) * (1.0 / 4294967296.0); }

__device__ double gamma_variate(double shape, uint32_t k0, uint32_t k1, uint32_t sweep, uint32_t idx, int stage,
                                int boost_stage)
{
    const double a1 = (shape < 1.0) ? shape + 1.0 : shape;
    const double d = a1 - 1.0 / 3.0;
    const double c = 1.0 / sqrt(9.0 * d);
    double y = d;
    for (uint32_t t = 0; t < (1u << 20); t++) {
        const uint4 o = philox4x32_10(idx, t, sweep, (uint32_t)stage << 28, k0, k1);
        const double r1 = u53(o.x, o.y);
        const double r2 = u32d(o.z);
        const double z = sqrt(-2.0 * log(r1)) * cos(6.283185307179586476925 * r2);
        double vv = 1.0 + c * z;
        if (vv <= 0.0) continue;
        vv = vv * vv * vv;
        const double r3 = u32d(o.w);
        if (log(r3) < 0.5 * z * z + d - d * vv + d * log(vv)) { y = d * vv; break; }
    }
    if (shape < 1.0) {
        const uint4 o = philox4x32_10(idx, 0u, sweep, (uint32_t)boost_stage << 28, k0, k1);
        y *= exp(log(u53(o.x, o.y)) / shape);
    }
    return y;
}

#define DRAW_MAX_THREADS 1024
struct DrawParams {
    const unsigned long long *sum_mu;  // [S][G]
    const unsigned long long *esum;    // [16] esum[a_obs*4+b_true]
    int S, G;
    double alpha, delta, epsilon;
    uint64_t seed;
    uint32_t sweep;
    double *gamma_out;  // [S][G]
    double *eta_out;    // [16]
    unsigned long long *esum_keep;  // if non-null: [16] copy of esum (E_store[i].sum(axis=(0,1)) of this sweep, HaploSNP_Sampler.py:557)
};

__global__ void __launch_bounds__(DRAW_MAX_THREADS) draw_gamma_eta_kernel(DrawParams p)
{
    pdl_enter();
    KPROF_SCOPE(KP_DRAW);
    extern __shared__ double y[];  // [S*G + 16]
    const int S = p.S, G = p.G, nG = S * G;
    const uint32_t k0 = (uint32_t)p.seed, k1 = (uint32_t)(p.seed >> 32);
    for (int i = threadIdx.x; i < nG + 16; i += blockDim.x) {
        if (i < nG) {
            y[i] = gamma_variate(p.alpha + (double)p.sum_mu[i], k0, k1, p.sweep, (uint32_t)i, STAGE_GAMMA,
                                 STAGE_GAMMA_BOOST);
        } else {
            if (p.esum_keep) p.esum_keep[i - nG] = p.esum[i - nG];
            const int j = i - nG, t = j >> 2, o = j & 3;   // eta[t][o] ~ Gamma(delta + Esum[o][t])  (:276-281)
            y[i] = gamma_variate(p.delta + (double)p.esum[o * 4 + t], k0, k1, p.sweep, (uint32_t)j, STAGE_ETA,
                                 STAGE_ETA_BOOST);
        }
    }
    __syncthreads();
    for (int s = threadIdx.x; s < S + 4; s += blockDim.x) {
        if (s < S) {
            double tot = 0.0;
            for (int g = 0; g < G; g++) tot += y[s * G + g];
            double rs = 0.0;
            for (int g = 0; g < G; g++) {
                double x = (tot > 0.0) ? y[s * G + g] / tot : 1.0 / G;
                if (x < p.epsilon) x = p.epsilon;             // :271
                y[s * G + g] = x;
                rs += x;
            }
            for (int g = 0; g < G; g++) p.gamma_out[s * G + g] = y[s * G + g] / rs;   // :272-273
        } else {
            const int t = s - S;
            double tot = 0.0;
            for (int o = 0; o < 4; o++) tot += y[nG + t * 4 + o];
            for (int o = 0; o < 4; o++) p.eta_out[t * 4 + o] = y[nG + t * 4 + o] / tot;
        }
    }
}

// ---------------------------------------------------------------------------------------------
struct FinalParams {
    const long long *red_i;   // [2] fixed-point sum n*log p | nchange (already all-reduced under sharding)
    double ll_const, ll_inv_scale;
    const unsigned int *agg_nslots;   // pattern table bookkeeping: ask for a rebuild when too many slots were handed out
    int *agg_ctl;
    unsigned int agg_limit;
    int *gctl;                // site-group control words (tau_group_kernel.cuh), or null
    long long V_local;
    const double *gamma;      // [S][G] used for the prior
    const double *eta;        // [16]   used for the prior (eta_new in update())
    double *eta_commit;       // if non-null: eta_commit[0..15] = eta (the chain's eta <- eta_new)
    int S, G;
    double V_total;
    double alpha, delta;
    double lg_alphaG, lg_alpha, lg_delta4, lg_delta;   // lgamma(alpha*G), lgamma(alpha), lgamma(4 delta), lgamma(delta)
    int it;                   // iteration index, or -1 for the pre-sweep state (:336-338)
    int star_mode;            // 0: gamma/eta/tau star (update), 1: tau only (updateTau)
    double *ll_store, *lp_store, *nchange_store;   // device [n_iter] or null
    double *gamma_store, *eta_store;               // device [n_iter][S*G], [n_iter][16] or null
    double *gamma_star, *eta_star;                 // device
    double *scal;             // [0]=lp_star [1]=iter_star [2]=ll [3]=lp
    int *flag;                // 1 when the star state must be replaced
};

// logPosterior (HaploSNP_Sampler.py:444-461) + star bookkeeping (:326-332, :351-358), in two parts that the fused
// exchange + finalize kernel of the sharded chain (exchange_kernel.cuh) shares with finalize_sweep_kernel.  Both parts are
// executed by ALL threads of a block of >= 256 threads (they contain block barriers); threads >= 256 only take part in those.
struct FinalEarly { double prior, lp_star; };
// part 1: the log-priors and lp_star -- nothing of it was written by the grid right before (see finalize_sweep_kernel)
__device__ __forceinline__ FinalEarly finalize_early(const FinalParams &p, double *sh /* [256] shared */)
{
    const int nG = p.S * p.G, t = threadIdx.x;
    double acc = 0.0;
    if (t < 256) {
        for (int i = t; i < nG; i += 256) acc += (p.alpha - 1.0) * log(p.gamma[i]);
        for (int i = t; i < 16; i += 256) acc += (p.delta - 1.0) * log(p.eta[i]);
        sh[t] = acc;
    }
    __syncthreads();
    for (int m = 128; m > 0; m >>= 1) {
        if (t < m) sh[t] += sh[t + m];
        __syncthreads();
    }
    FinalEarly e;
    e.prior = 0.0; e.lp_star = 0.0;
    if (t == 0) {
        e.prior = sh[0] + p.S * (p.lg_alphaG - p.G * p.lg_alpha) + 4.0 * (p.lg_delta4 - 4.0 * p.lg_delta) +
                  p.V_total * (double)p.G * log(0.25);
        e.lp_star = p.scal[0];
    }
    return e;
}
// part 2: needs red_i (the log-likelihood pass, or the exchange that summed it over the ranks)
__device__ __forceinline__ void finalize_late(const FinalParams &p, const FinalEarly &e, int *upd /* shared */)
{
    const int nG = p.S * p.G, t = threadIdx.x;
    if (t == 0) {
        const double ll = p.ll_const + (double)p.red_i[0] * p.ll_inv_scale, lp = ll + e.prior;
        // upkeep wishes of the sweep (ll_table_kernel; summed over the ranks of a sharded chain by the exchange): a table rebuild
        // (which implies a regroup) and / or a regroup of the site groups, on every rank in the same sweep
        const unsigned long long wish = (unsigned long long)p.red_i[2];
        if (p.agg_ctl && (wish >> 16)) p.agg_ctl[0] = 1;
        if (p.gctl && p.it >= 0) {
            // the screening pass pays off while few sites flip
            p.gctl[GC_CALM] = (double)p.red_i[1] <= p.V_total / 16.0;
            if (wish & 0xffffull) p.gctl[GC_REGROUP] = 1;
        }
        p.scal[2] = ll; p.scal[3] = lp;
        if (p.it >= 0) {
            if (p.ll_store) p.ll_store[p.it] = ll;
            if (p.lp_store) p.lp_store[p.it] = lp;
            if (p.nchange_store) p.nchange_store[p.it] = (double)p.red_i[1];
        }
        *upd = (p.it < 0) || (lp > e.lp_star);
        if (*upd) { p.scal[0] = lp; p.scal[1] = (double)(p.it < 0 ? 0 : p.it); }
        *p.flag = *upd;
    }
    __syncthreads();
    const bool u = *upd != 0;
    if (t < 256) {
        for (int i = t; i < nG; i += 256) {
            const double x = p.gamma[i];
            if (p.it >= 0 && p.gamma_store) p.gamma_store[(size_t)p.it * nG + i] = x;
            if (u && p.star_mode == 0) p.gamma_star[i] = x;
        }
        if (t < 16) {
            const double x = p.eta[t];
            if (p.it >= 0 && p.eta_store) p.eta_store[(size_t)p.it * 16 + t] = x;
            if (u && p.star_mode == 0) p.eta_star[t] = x;
            if (p.eta_commit) p.eta_commit[t] = x;
        }
    }
}

__global__ void __launch_bounds__(256) finalize_sweep_kernel(FinalParams p)
{
    // The log-priors and the bookkeeping words read below were written at least two grids ago (gamma, eta_new: the draw
    // kernel; lp_star: the previous finalize) -- the grid right before this one is the log-likelihood pass or the exchange,
    // which write red_i only.  So everything but red_i is formed before pdl_enter(), under the tail of that grid (PDL_EARLY,
    // common.cuh).
#if !PDL_EARLY
    pdl_enter();
#endif
    KPROF_SCOPE(KP_FIN);
    __shared__ double sh[256];
    __shared__ int upd;
    const FinalEarly e = finalize_early(p, sh);
#if PDL_EARLY
    pdl_enter();
#endif
    finalize_late(p, e, &upd);
}

// count cells uploaded as 4 x uint16 -> the canonical int32x4 cells (desman_set_counts)
__global__ void widen_counts_kernel(const uint2 *__restrict__ src, int4 *__restrict__ dst, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint2 w = src[i];
        dst[i] = make_int4((int)(w.x & 0xffffu), (int)(w.x >> 16), (int)(w.y & 0xffffu), (int)(w.y >> 16));
    }
}

__global__ void copy_tau_if_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, size_t n,
                                   const int *__restrict__ flag)
{
    pdl_enter();
    KPROF_SCOPE(KP_COPY);
    if (*flag == 0) return;
    // 16-byte words (cudaMalloc'ed buffers are 256-byte aligned), byte tail
    const size_t n16 = n / 16, i0 = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
    uint4 *d4 = reinterpret_cast<uint4 *>(dst);
    for (size_t i = i0; i < n16; i += st) d4[i] = s4[i];
    for (size_t i = n16 * 16 + i0; i < n; i += st) dst[i] = src[i];
}

// close the lazy occupancy counters at the end of an update(): cnt[vg][tau_vg] += n_iter - last[vg]
__global__ void flush_tau_counts_kernel(const uint8_t *__restrict__ tau, uint32_t *__restrict__ cnt,
                                        uint32_t *__restrict__ last, size_t nvg, uint32_t n_iter)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nvg; i += (size_t)gridDim.x * blockDim.x) {
        cnt[i * 4 + (tau[i] & 3)] += n_iter - last[i];
        last[i] = n_iter;
    }
}

__global__ void l2_flush_kernel(uint4 *__restrict__ buf, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        buf[i] = make_uint4((uint32_t)i, 0u, 0u, 0u);
}
