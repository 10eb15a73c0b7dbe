// desman_b200/csrc/tau_kernel.cuh -- K1: the tau Gibbs update (replaces sampletau/c_sample_tau.c:95-204)
// It also keeps the persistent pattern table (mu_agg_kernel.cuh) current: a site whose pattern changed moves its
// counts from the old slot to the new one, so K2b and the log-likelihood never need a fresh aggregation pass.
//
// Mapping: one warp per variant position v (sites are independent, c_sample_tau.c:130), lanes over
// samples s, strains g strictly in order (each draw conditions on the tau just written, :133,:180).
// The site's S count cells (one 128-bit int32x4 word per (v,s)) are staged once into the warp's
// shared-memory tile and re-read for every strain.
//
// Arithmetic: "filtered exact".  The reference evaluates 16*S logs in FP64 per (v,g); under that
// arithmetic the kernel is bound by the FP64/log rate, ~100x above its HBM floor.  What has to be
// reproduced, though, is only the DRAW t = sample4(softmax(L), u).  So per (v,g):
//   tier 1  log-likelihood differences D_a = L_a - L_cur in FP32 (MUFU lg2.approx) with a rigorous
//           running error bound B_a.  If one candidate leads every other by more than TAU_GAP = 26 nats even
//           after the bounds, its probability differs from 1 by < 1.6e-11 < 2^-33 <= what u can come within,
//           so the draw is decided without evaluating a single exp.
//   tier 2  otherwise the CDF boundaries are bracketed in FP64 from D_a +- B_a; if u lies outside
//           every bracket (plus 1e-9 slack) the draw is decided.
//   tier 3  otherwise (u within the bracket of a boundary, or u == 0) the (v,g) step is recomputed
//           with the reference's FP64 arithmetic and operation order (tau_exact_logp below).
// Cancellation-free by construction: the mixture P[s][b] = sum_h eta[tau_h][b]*gamma[s][h] is kept in
// FP64 and base = P - eta[cur][b]*gamma[s][g] is formed in FP64 before rounding to FP32; every FP32
// quantity that enters a log is then a sum of non-negative terms (relative error a few ulp).
#pragma once
#include "common.cuh"
#include "mu_agg_kernel.cuh"

struct TauParams {
    const int4 *counts;      // [V][S] int32x4
    uint8_t *tau;            // [V][G] base index, updated in place
    const double *gamma;     // [S][G]
    const double *eta;       // [16] eta used for the draw (row = true base)
    const uint32_t *words;   // MT19937 words [V*G] (u = w/2^32), or nullptr -> Philox
    uint64_t seed;
    uint32_t sweep;
    int64_t v0;              // global index of local site 0 (Philox counter / sharding)
    int V, S, G;
    unsigned long long *nchange;  // += flips
    AggTable agg;            // persistent pattern table to keep current (agg.N == nullptr: none)
    uint32_t *tau_cnt;       // [V][G][4] lazy per-base occupancy counters, or nullptr
    uint32_t *tau_last;      // [V][G] iteration at which the current base was adopted
    uint32_t iter;           // iteration index inside the current update() call
    int exact_only;          // 1: every (v,g) step takes the FP64 reference-order path (validation)
    // grouped mode (tau_group_kernel.cuh): walk only the listed sites and, until the first flip of a site, only the
    // strains the screening pass left undecided.  Active iff gctl[GC_HAVE] && gctl[GC_CALM] (same test as the screening pass).
    const uint2 *work;       // {site, mask of undecided strains}, gctl[GC_NWORK] entries; nullptr: every site, every strain
    const int *singles;      // sites alone in their pattern, gctl[GC_NSINGLES] entries
    int *gctl;
    int *site_slot;          // [V] slot of the site's pattern, kept current on flips
    // tensor-memory screening pass (tau_group_tc_kernel.cuh): a site that leaves its group is marked in the row table of the
    // count image (img_site[row] = ~site), so that pass needs no gather to recognise orphans; need_img: the work list is
    // only valid if the image was built (gctl[GC_IMG_OK])
    int *img_site;
    const int *site_row;
    int need_img;
    // sampleTauFixTau (HaploSNP_Sampler.py:196-222): strains below g_begin keep their base; logp_out [V][4] (or nullptr) gets the
    // normalised log-probabilities (normaliseLogProb, :186-194) of strain g_begin's four bases before its draw
    int g_begin;
    double *logp_out;
    int skip_listed;         // 1: when the work list is valid this launch does nothing (tau_open_kernel walks the list)
    // batched small problems (Eta_Sampler.sampleTauC: one call per gene, each with its own masked gamma; Eta_Sampler.py:355-369):
    // the sites of all problems are concatenated, prob_off[k] .. prob_off[k+1] are the sites of problem k, gamma holds one
    // [S][G] matrix per problem and blockIdx.y is the problem a CTA works on; nullptr: one problem
    const int *prob_off;
    unsigned long long *tier_counts;  // [3] += draws decided by tier 1 / 2 / 3 (or nullptr)
};

#define TAU_WARPS 8
#ifndef TAU_GAP
// nats.  The uniforms live on a 2^-32 grid and are never 0 on the paths that use the gap test (Philox: u = (w + 1/2) / 2^32;
// MT19937: a zero word goes to the reference-order path), so u is in [2^-33, 1 - 2^-32].  If one base leads every other by more
// than g after the error bounds, every CDF boundary of sample4 is within 3 e^-g of 0 or 1; 3 e^-g < 2^-33 needs g > 23.97.
// 26 leaves a factor 7.6 (3 e^-26 = 1.5e-11 against 2^-33 = 1.2e-10).  (60 in round 1: same-box A/B at C3, work list 2935 -> 640
// sites, per-site kernel 29 -> 21 us, chains bit-identical.)
#define TAU_GAP 26.0f
#endif
#define TAU_SLACK 1.0e-9         // absolute slack on CDF brackets evaluated in FP64
#define TAU_SLACK32 1.0e-4       // ... and in FP32 fast math (tau_bracket_decide)

// ---------------------------------------------------------------------------------------------
// Reference arithmetic (c_sample_tau.c:136-170) in FP64 for ONE (v,g): base over h ascending skipping
// g from 0.0, candidate term added last, count through float.  Terms with n == 0 are skipped:
// 0*log(p) adds exactly -0.0 there (p > 0 because eta, gamma > 0).  All lanes of the warp call this.
__device__ __noinline__ void tau_exact_logp(const int4 *tile, const double *gT, const double *eta_s, uint64_t code,
                                               int g, int S, int Sp, int G, int lane, double L[4])
{
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
        const double f0 = (double)(float)n.x, f1 = (double)(float)n.y, f2 = (double)(float)n.z, f3 = (double)(float)n.w;
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
    L[0] = warp_sum(L0); L[1] = warp_sum(L1); L[2] = warp_sum(L2); L[3] = warp_sum(L3);
}

// The same sum for ONE candidate base a (same lane / chunk order, same operations: bit-identical to L[a] of tau_exact_logp);
// tau_open_kernel spreads the four candidates of a step over its four warps.
__device__ __noinline__ double tau_exact_logp_cand(const int4 *tile, const double *gT, const double *eta_s, uint64_t code,
                                                   int g, int S, int Sp, int G, int lane, int a)
{
    double L = 0.0;
    const double *e = eta_s + 4 * a;
    for (int s = lane; s < S; s += 32) {
        const int4 n = tile[s];
        if ((n.x | n.y | n.z | n.w) == 0) continue;
        double b0 = 0.0, b1 = 0.0, b2 = 0.0, b3 = 0.0;
        for (int h = 0; h < G; h++) {
            if (h == g) continue;
            const double *eh = eta_s + 4 * code_get(code, h);
            const double gm = gT[h * Sp + s];
            b0 = fma(eh[0], gm, b0); b1 = fma(eh[1], gm, b1);
            b2 = fma(eh[2], gm, b2); b3 = fma(eh[3], gm, b3);
        }
        const double gg = gT[g * Sp + s];
        const double f0 = (double)(float)n.x, f1 = (double)(float)n.y, f2 = (double)(float)n.z, f3 = (double)(float)n.w;
        // (branch-free: the four logs of a sample are independent and overlap; a zero count takes log(1) = 0 and
        // fma(0, 0, L) == L, the value the skipped term leaves in tau_exact_logp)
        const double l0 = log(n.x ? fma(e[0], gg, b0) : 1.0), l1 = log(n.y ? fma(e[1], gg, b1) : 1.0),
                     l2 = log(n.z ? fma(e[2], gg, b2) : 1.0), l3 = log(n.w ? fma(e[3], gg, b3) : 1.0);
        L = fma(f0, l0, L); L = fma(f1, l1, L); L = fma(f2, l2, L); L = fma(f3, l3, L);
    }
    return warp_sum(L);
}

// normaliseLog4 + sample4 (c_sample_tau.c:48-91)
__device__ __noinline__ int tau_exact_pick(const double L[4], double u)
{
    double mx = L[0];
    if (L[1] > mx) mx = L[1];
    if (L[2] > mx) mx = L[2];
    if (L[3] > mx) mx = L[3];
    const double e0 = exp(L[0] - mx), e1 = exp(L[1] - mx), e2 = exp(L[2] - mx), e3 = exp(L[3] - mx);
    const double sum = ((0.0 + e0) + e1) + e2 + e3;
    const double p0 = e0 / sum, p1 = e1 / sum, p2 = e2 / sum;
    const double c0 = p0, c1 = p1 + c0, c2 = p2 + c1;
    return (u < c0) ? 0 : (u < c1) ? 1 : (u < c2) ? 2 : 3;
}

// Sum four per-lane floats over the warp with 10 shuffles instead of 20 (transpose-reduce), result in all lanes.
__device__ __forceinline__ void warp_sum4(float &a, float &b, float &c, float &d, int lane)
{
    const bool up16 = lane & 16;
    float s0 = up16 ? a : c, s1 = up16 ? b : d;          // what I send
    float k0 = up16 ? c : a, k1 = up16 ? d : b;          // what I keep (lanes <16 keep a,b; >=16 keep c,d)
    k0 += __shfl_xor_sync(DESMAN_FULL_MASK, s0, 16);
    k1 += __shfl_xor_sync(DESMAN_FULL_MASK, s1, 16);
    const bool up8 = lane & 8;
    float s = up8 ? k0 : k1, k = up8 ? k1 : k0;          // lanes with bit3 clear keep k0, set keep k1
    k += __shfl_xor_sync(DESMAN_FULL_MASK, s, 8);
    k += __shfl_xor_sync(DESMAN_FULL_MASK, k, 4);
    k += __shfl_xor_sync(DESMAN_FULL_MASK, k, 2);
    k += __shfl_xor_sync(DESMAN_FULL_MASK, k, 1);
    // value index held by lane: (lane>>4)*2 + ((lane>>3)&1)  -> a:0-7, b:8-15, c:16-23, d:24-31
    a = __shfl_sync(DESMAN_FULL_MASK, k, 0);
    b = __shfl_sync(DESMAN_FULL_MASK, k, 8);
    c = __shfl_sync(DESMAN_FULL_MASK, k, 16);
    d = __shfl_sync(DESMAN_FULL_MASK, k, 24);
}

// error-model constants (log2 units per read); u = 2^-24
//   q = base32 + eta32*gamma32: base32 carries 1 rounding (from FP64), the product 3, the sum 1 -> <= 5u relative
//   lg2.approx: abs error <= 2^-22 * max(1, |lg2 q|)   (CUDA math API, __log2f)
//   lP = lg2((float)P64): 1u relative + the same lg2 bound
#define TAU_C0 (6.0f * 5.9604645e-8f * 1.4426950f + 2.0f * 2.3841858e-7f)   // relative parts + the two max(1,.) floors
// per unit of |lg2|: 2^-22 (lg2.approx) and, for the FP32 accumulation of sum n*lg2 q and sum n*lg2 P over
// 4*nch terms per lane + 5 reduction levels, (4*nch+8)*2^-24  (every partial sum is <= reads * max|lg2|)
#define TAU_C1(nch) (2.3841858e-7f + (float)(4 * (nch) + 8) * 5.9604645e-8f)
// FP64 mixture P carries <= (2G+2)*2^-53*P absolute error; relative to a candidate q that is amplified by P/q = 2^(lP-lq)
#define TAU_CANCEL(G) ((float)(2 * (G) + 2) * 1.1102230e-16f * 1.4426950f)
// The FP32 path needs every candidate q = base + eta*gamma to be a NORMAL float with a finite log even where the
// count is 0 (0*lg2 q must be 0).  q >= min(eta)*min(gamma) =: qmin, evaluated per launch; if qmin < TAU_QMIN every
// draw of the launch takes the FP64 path.  qmin also gives the a-priori bounds |lg2 q| <= -lg2(qmin)+1 and
// P/q <= 1/qmin used by the cheap first-tier error bound.
#define TAU_QMIN 1.0e-36f

__device__ __forceinline__ float lg2_fast(float x)
{
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));   // bare MUFU.LG2: no denormal rescaling code around it
    return y;
}

// FP32 evaluation of one (v,g) step for the three candidates a_j = (cur+1+j)&3:
//   E_j = sum_s sum_b n_sb * lg2(base_sb + eta[a_j][b]*gamma[s][g]),   KK = sum_s sum_b n_sb * lg2 P_sb   (log2 units)
// TRACK additionally returns max |lg2 q| over the lane's terms (tight error bound for the bracket test).
template <bool TRACK>
__device__ __forceinline__ void tau_fp32_terms(const int4 *tile, const double2 *Pw, const float *Kw, const double *gTg,
                                               const float *gT32g, const double *eta_cur, const float4 *eta32, int cur,
                                               int nch, int lane, float &E0, float &E1, float &E2, float &KK, float &mq)
{
    const double2 *ecp = reinterpret_cast<const double2 *>(eta_cur);
    const double2 ec01 = ecp[0], ec23 = ecp[1];
    const float4 ea = eta32[(cur + 1) & 3], eb4 = eta32[(cur + 2) & 3], ecc = eta32[(cur + 3) & 3];
    E0 = 0.f; E1 = 0.f; E2 = 0.f; KK = 0.f; mq = 1.0f;
    for (int c = 0; c < nch; c++) {
        const int s = c * 32 + lane;
        const int4 n = tile[s];
        const double gg = gTg[s];
        const float gf = gT32g[s];
        const double2 P01 = Pw[s * 2], P23 = Pw[s * 2 + 1];
        KK += Kw[s];
        // base = P - eta[cur][b]*gamma (FP64: no cancellation error), then one rounding to FP32
        const float q0 = fmaxf((float)fma(-ec01.x, gg, P01.x), 0.f), q1 = fmaxf((float)fma(-ec01.y, gg, P01.y), 0.f),
                    q2 = fmaxf((float)fma(-ec23.x, gg, P23.x), 0.f), q3 = fmaxf((float)fma(-ec23.y, gg, P23.y), 0.f);
        const float f0 = (float)n.x, f1 = (float)n.y, f2 = (float)n.z, f3 = (float)n.w;
#define TAU_TERM(fb, qb, eab, E)                                               \
    {                                                                          \
        const float lq = lg2_fast(fmaf(eab, gf, qb));                          \
        E = fmaf(fb, lq, E);                                                   \
        if (TRACK) mq = fmaxf(mq, fabsf(lq));                                  \
    }
        // 12 independent FFMA -> MUFU.LG2 -> FFMA chains per chunk (straight-line: the scheduler interleaves them)
        TAU_TERM(f0, q0, ea.x, E0) TAU_TERM(f1, q1, ea.y, E0) TAU_TERM(f2, q2, ea.z, E0) TAU_TERM(f3, q3, ea.w, E0)
        TAU_TERM(f0, q0, eb4.x, E1) TAU_TERM(f1, q1, eb4.y, E1) TAU_TERM(f2, q2, eb4.z, E1) TAU_TERM(f3, q3, eb4.w, E1)
        TAU_TERM(f0, q0, ecc.x, E2) TAU_TERM(f1, q1, ecc.y, E2) TAU_TERM(f2, q2, ecc.z, E2) TAU_TERM(f3, q3, ecc.w, E2)
#undef TAU_TERM
    }
}

// tier 2 (rare, kept out of line so the hot loop stays small in the instruction cache)
__device__ __noinline__ int tau_bracket_decide(const int4 *tile, const double2 *Pw, const float *Kw, const double *gTg,
                                               const float *gT32g, const double *eta_cur, const float4 *eta32, int cur,
                                               int nch, int lane, float nlane, float mlP, float c1, float ccan, double u)
{
    float E0, E1, E2, KK, mq;
    tau_fp32_terms<true>(tile, Pw, Kw, gTg, gT32g, eta_cur, eta32, cur, nch, lane, E0, E1, E2, KK, mq);
    const float LN2 = 0.69314718f;
    float D0 = E0 - KK, D1 = E1 - KK, D2 = E2 - KK;
    float eb = nlane * (TAU_C0 + c1 * (mq + mlP) + ccan * exp2f(fminf(mq + mlP, 120.f)));
    warp_sum4(D0, D1, D2, eb, lane);
    const float Bn = eb * LN2 * 1.0001f + 1e-6f;
    if (!((fabsf(D0) + fabsf(D1) + fabsf(D2) + Bn) < 1.0e30f)) return -1;
    // The brackets themselves are evaluated in FP32 (ex2.approx, fast division): every bracket end carries a relative error
    // below 1e-5 (argument rounding |x| 2^-24 log2e for |x| <= 100, ex2.approx 2^-22, three adds, one division), absorbed
    // by the absolute slack TAU_SLACK32 = 1e-4 on the CDF.  u inside a widened bracket (probability ~2e-4 per test) goes to
    // tier 3 like any other undecided draw.
    // the exponents are differenced in FP64 first (|D| can be 1e5 nats, where a float ulp is 8e-3), then exponentiated in FP32
    const double x0 = (double)(D0 * LN2), x1 = (double)(D1 * LN2), x2 = (double)(D2 * LN2), B = (double)Bn;
    double dh[4], dl[4];
    double M = -1.0e300;
    // a_j = (cur+1+j)&3  <=>  j = (a-cur-1)&3 ; j == 3 is cur itself (exactly 0, no error)
#pragma unroll
    for (int a = 0; a < 4; a++) {
        const int j = (a - cur - 1) & 3;
        const double d = (j == 0) ? x0 : (j == 1) ? x1 : (j == 2) ? x2 : 0.0;
        const double bb = (j == 3) ? 0.0 : B;
        dh[a] = d + bb; dl[a] = d - bb;
        M = fmax(M, dh[a]);
    }
    float eh[4], el[4];
#pragma unroll
    for (int a = 0; a < 4; a++) {
        // FP32 range: terms below e^-80 of the largest are rounded in the SAFE direction -- the upper ends up (to e^-80), the
        // lower ends down (to 0) -- which can only widen a bracket (cplus = ah/(ah+rl) grows, cminus = al/(al+rh) shrinks)
        eh[a] = __expf((float)fmax(dh[a] - M, -80.0));
        el[a] = (dl[a] - M < -80.0) ? 0.0f : __expf((float)(dl[a] - M));
    }
    int below = 0, above = 0;     // number of boundaries certainly > u / certainly <= u
    float ah = 0.0f, al = 0.0f;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        ah += eh[k]; al += el[k];
        float rh = 0.0f, rl = 0.0f;
#pragma unroll
        for (int a = k + 1; a < 4; a++) { rh += eh[a]; rl += el[a]; }
        const float cplus = __fdividef(ah, ah + rl), cminus = __fdividef(al, al + rh);
        if (u < (double)cminus - TAU_SLACK32) below++;
        else if (u >= (double)cplus + TAU_SLACK32) above++;
    }
    return (below + above == 3) ? above : -1;    // boundaries are ordered: t = #boundaries <= u
}

__global__ void __launch_bounds__(TAU_WARPS * 32, 3) tau_sample_kernel(TauParams p)
{
    // After the screening pass, gamma and eta are operands written at least two grids ago: the prologue that stages them runs
    // before pdl_enter() (PDL_EARLY, common.cuh).  Without the screening pass the preceding grid may be the one that drew gamma.
#if PDL_EARLY
    const bool early = p.work != nullptr;
#else
    const bool early = false;
#endif
    // the work list is walked by tau_open_kernel: nothing to do here when it is valid (the control words that say so were
    // written at least two grids ago; the wait still has to happen: the next grid of the chain orders itself after THIS one)
    if (p.skip_listed && grp_active(p.gctl, p.need_img)) { pdl_enter(); return; }
    if (!early) pdl_enter();
    KPROF_SCOPE(KP_TAU);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int S = p.S, G = p.G;
    const int Sp = (S + 31) & ~31;
    double *gT = reinterpret_cast<double *>(smem_raw);           // [G][Sp] gamma transposed (FP64)
    double *eta_s = gT + (size_t)G * Sp;                         // [16]
    double *etall_s = eta_s + 16;                                // [16]
    double2 *P64 = reinterpret_cast<double2 *>(etall_s + 16);    // [TAU_WARPS][Sp][2] mixture probabilities (b0,b1),(b2,b3)
    float *gT32 = reinterpret_cast<float *>(P64 + (size_t)TAU_WARPS * Sp * 2);   // [G][Sp]
    float4 *eta32 = reinterpret_cast<float4 *>(gT32 + (size_t)G * Sp);           // [4] rows
    float *K = reinterpret_cast<float *>(eta32 + 4);             // [TAU_WARPS][Sp] sum_b n_b*lg2 P_b
    int4 *tiles = reinterpret_cast<int4 *>(K + (size_t)TAU_WARPS * Sp);          // [TAU_WARPS][Sp]
    uint32_t *wbuf = reinterpret_cast<uint32_t *>(tiles + (size_t)TAU_WARPS * Sp);   // [TAU_WARPS][32] uniform words
    __shared__ unsigned int gmin_bits, emin_bits;   // min gamma / min eta as float bit patterns (positive floats order like uints)

    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    if (threadIdx.x == 0) { gmin_bits = 0x7f800000u; emin_bits = 0x7f800000u; }
    __syncthreads();
    const double *gamma_k = p.prob_off ? p.gamma + (size_t)blockIdx.y * S * G : p.gamma;   // this CTA's problem
    float gmin_l = __int_as_float(0x7f800000);
    for (int i = threadIdx.x; i < G * Sp; i += blockDim.x) {
        const int g = i / Sp, s = i - g * Sp;
        const double x = (s < S) ? gamma_k[(size_t)s * G + g] : 0.0;
        gT[i] = x;
        gT32[i] = (float)x;
        if (s < S && x > 0.0) gmin_l = fminf(gmin_l, (float)x);     // masked strains (gamma == 0): q = P there
    }
    atomicMin(&gmin_bits, __float_as_uint(gmin_l));
    if (threadIdx.x < 16) {
        eta_s[threadIdx.x] = p.eta[threadIdx.x];
        reinterpret_cast<float *>(eta32)[threadIdx.x] = (float)p.eta[threadIdx.x];
        etall_s[threadIdx.x] = 0.0;
        atomicMin(&emin_bits, __float_as_uint(fmaxf((float)p.eta[threadIdx.x], 0.f)));
    }
    __syncthreads();
    const float qmin = 0.99f * __uint_as_float(gmin_bits) * __uint_as_float(emin_bits);
    const bool fast_ok = !p.exact_only && qmin >= TAU_QMIN;
    const float mq0 = fmaxf(1.0f, 1.0f - log2f(fmaxf(qmin, TAU_QMIN)));      // a-priori max |lg2 q| (q <= 2)

    int4 *tile = tiles + (size_t)wib * Sp;
    double2 *Pw = P64 + (size_t)wib * Sp * 2;
    float *Kw = K + (size_t)wib * Sp;
    uint32_t *ww = wbuf + wib * 32;
    const int nch = Sp >> 5;
    const int gw = blockIdx.x * TAU_WARPS + wib, nw = gridDim.x * TAU_WARPS;
    const uint32_t k0 = (uint32_t)p.seed, k1 = (uint32_t)(p.seed >> 32);
    const float c1 = TAU_C1(nch), ccan = TAU_CANCEL(G);
    const float cancel0 = ccan / fmaxf(qmin, TAU_QMIN);                     // a-priori bound of the P/q amplification
    unsigned int flips = 0, n1 = 0, n2 = 0, n3 = 0;
#ifdef KPROF
    const unsigned long long kp_pro = gtimer();
    unsigned long long kp_stage = 0, kp_steps = 0, kp_move = 0, kp_first = 0;
    int kp_sites = 0;
#define KP_T(x) const unsigned long long x = gtimer()
#else
#define KP_T(x)
#endif

    if (early) pdl_enter();
    int nsite = p.V, nwork = 0;
    bool listed = false;
    if (p.work && grp_active(p.gctl, p.need_img)) {
        listed = true;
        nwork = p.gctl[GC_NWORK];
        nsite = nwork + p.gctl[GC_NSINGLES];
    }

    int i_begin = gw, i_end = nsite;
    if (p.prob_off) { i_begin = p.prob_off[blockIdx.y] + gw; i_end = p.prob_off[blockIdx.y + 1]; }
    for (int i = i_begin; i < i_end; i += nw) {
        int v = i;
        uint32_t todo = 0xffffffffu;                 // strains the screening pass did not decide
        bool screened = false;                       // todo comes from the gap test of the screening pass
        if (listed) {
            if (i < nwork) {
                const uint2 e = p.work[i]; v = (int)e.x; todo = e.y;
                // a full mask is what orphans of their group (and sites with a zero MT word) get without any test: their
                // steps are ordinary ones, mostly settled by the cheap gap test, and must not be sent straight to the brackets
                screened = todo != ((G >= 32) ? 0xffffffffu : ((1u << G) - 1u));
            } else v = p.singles[i - nwork];
        }
        KP_T(kp0);
        const int4 *src = p.counts + (size_t)v * S;
        uint64_t code = load_tau_code(p.tau + (size_t)v * G, G, lane);
        const uint64_t code_in = code;
        // the G uniform words of this site: lane g draws word g (one Philox call per site instead of G per lane)
        {
            uint32_t w = 0;
            if (lane < G) {
                if (p.words) w = p.words[(size_t)v * G + lane];
                else w = philox4x32_10((uint32_t)(p.v0 + v), (uint32_t)lane, p.sweep, (uint32_t)STAGE_TAU << 28, k0, k1).x;
            }
            ww[lane] = w;
        }
        // stage counts; mixture P (FP64, ascending h); K = sum_b n_b lg2 P_b; per-lane read count and max |lg2 P|
        float nlane = 0.0f, mlP = 1.0f;
        for (int c = 0; c < nch; c++) {
            const int s = c * 32 + lane;
            int4 n = make_int4(0, 0, 0, 0);
            if (s < S) n = ld_counts(src + s);
            tile[s] = n;
            double b0 = 0.0, b1 = 0.0, b2 = 0.0, b3 = 0.0;
            for (int h = 0; h < G; h++) {
                const double2 *e = reinterpret_cast<const double2 *>(eta_s + 4 * code_get(code, h));
                const double2 e01 = e[0], e23 = e[1];
                const double gm = gT[h * Sp + s];
                b0 = fma(e01.x, gm, b0); b1 = fma(e01.y, gm, b1);
                b2 = fma(e23.x, gm, b2); b3 = fma(e23.y, gm, b3);
            }
            if (s >= S) { b0 = 1.0; b1 = 1.0; b2 = 1.0; b3 = 1.0; }   // padding lanes: finite logs, zero counts
            Pw[s * 2] = make_double2(b0, b1); Pw[s * 2 + 1] = make_double2(b2, b3);
            const float l0 = lg2_fast((float)b0), l1 = lg2_fast((float)b1), l2 = lg2_fast((float)b2), l3 = lg2_fast((float)b3);
            Kw[s] = fmaf((float)n.x, l0, fmaf((float)n.y, l1, fmaf((float)n.z, l2, (float)n.w * l3)));
            nlane += (float)(n.x + n.y + n.z + n.w);
            mlP = fmaxf(mlP, fmaxf(fmaxf(fabsf(l0), fabsf(l1)), fmaxf(fabsf(l2), fabsf(l3))));
        }
        __syncwarp();
        KP_T(kp1);

        for (int g = p.g_begin; g < G; g++) {
            if (!((todo >> g) & 1u) && code == code_in) { n1++; continue; }   // decided "stay" by the screening pass
            const int cur = code_get(code, g);
            const uint32_t w = ww[g];
            // MT19937 mode: gsl_rng_uniform, c_sample_tau.c:174 (u = 0 possible).  Philox mode: mid-point of the word's cell.
            const double u = p.words ? (double)w / 4294967296.0 : ((double)w + 0.5) / 4294967296.0;
            int t = -1;
            const bool record = p.logp_out != nullptr && g == p.g_begin;      // this step's log-probabilities are an output: FP64 path
            const bool usable = fast_ok && (w != 0u || !p.words) && !record;
            if (usable && screened && code == code_in) {
                // a step the screening pass left open: its gap test already failed on the same kind of sums, so go straight
                // to the tracked sums and the brackets (which also settle a decisive flip)
                t = tau_bracket_decide(tile, Pw, Kw, gT + g * Sp, gT32 + g * Sp, eta_s + 4 * cur, eta32, cur, nch, lane, nlane, mlP,
                                       c1, ccan, u);
                if (t >= 0) n2++;
            } else if (usable) {
                // ---- tier 1: D_j = (E_j - KK)*ln2 = L(a_j) - L(cur), j = 0..2, with the a-priori bounds from qmin
                const double *eta_cur = eta_s + 4 * cur;
                float E0, E1, E2, KK, mq;
                tau_fp32_terms<false>(tile, Pw, Kw, gT + g * Sp, gT32 + g * Sp, eta_cur, eta32, cur, nch, lane, E0, E1, E2, KK, mq);
                const float LN2 = 0.69314718f;
                float D0 = E0 - KK, D1 = E1 - KK, D2 = E2 - KK;
                float eb = nlane * (TAU_C0 + c1 * (mq0 + mlP) + cancel0);
                warp_sum4(D0, D1, D2, eb, lane);
                float Bn = eb * LN2 * 1.0001f + 1e-6f;              // nats
                float x0 = D0 * LN2, x1 = D1 * LN2, x2 = D2 * LN2;
                // leader among {cur (exactly 0), x0, x1, x2} and the best of the rest
                float top = fmaxf(fmaxf(x0, x1), x2);
                int jm = (x0 == top) ? 0 : (x1 == top) ? 1 : 2;
                const bool finite = (fabsf(x0) + fabsf(x1) + fabsf(x2) + Bn) < 1.0e30f;   // false for inf / NaN
                if (finite) {
                    if (top + Bn < -TAU_GAP) { t = cur; n1++; }     // cur leads everything by more than the gap
                    else {
                        const float rest = fmaxf(fmaxf(jm == 0 ? 0.f : x0, jm == 1 ? 0.f : x1), fmaxf(jm == 2 ? 0.f : x2, 0.f));
                        if ((top - Bn) - (rest + Bn) > TAU_GAP) { t = (cur + 1 + jm) & 3; n1++; }
                    }
                }
                if (t < 0 && finite) {
                    // ---- tier 2 (rare): tracked FP32 sums -> tight bound -> FP64 brackets of the CDF boundaries
                    t = tau_bracket_decide(tile, Pw, Kw, gT + g * Sp, gT32 + g * Sp, eta_cur, eta32, cur, nch, lane, nlane, mlP,
                                           c1, ccan, u);
                    if (t >= 0) n2++;
                }
            }
            if (t < 0) {
                // ---- tier 3: the reference's FP64 arithmetic and order
                double L[4];
                tau_exact_logp(tile, gT, eta_s, code, g, S, Sp, G, lane, L);
                t = tau_exact_pick(L, u);
                n3++;
                if (record && lane == 0) {
                    double mx = L[0];
                    for (int b = 1; b < 4; b++) if (L[b] > mx) mx = L[b];
                    double sum = 0.0;
                    for (int b = 0; b < 4; b++) sum += exp(L[b] - mx);
                    const double ls = log(sum);
                    for (int b = 0; b < 4; b++) p.logp_out[(size_t)v * 4 + b] = (L[b] - mx) - ls;
                }
            }
            if (t != cur) {
                // P += (eta[t][b] - eta[cur][b]) * gamma[s][g]; refresh K = sum_b n_b lg2 P_b
                const double *et = eta_s + 4 * t, *ec2 = eta_s + 4 * cur;
                const double e0 = et[0] - ec2[0], e1 = et[1] - ec2[1], e2 = et[2] - ec2[2], e3 = et[3] - ec2[3];
                mlP = 1.0f;
                for (int c = 0; c < nch; c++) {
                    const int s = c * 32 + lane;
                    const double gg = gT[g * Sp + s];
                    const double2 P01 = Pw[s * 2], P23 = Pw[s * 2 + 1];
                    const double b0 = fma(e0, gg, P01.x), b1 = fma(e1, gg, P01.y), b2 = fma(e2, gg, P23.x), b3 = fma(e3, gg, P23.y);
                    Pw[s * 2] = make_double2(b0, b1); Pw[s * 2 + 1] = make_double2(b2, b3);
                    const int4 n = tile[s];
                    const float l0 = lg2_fast((float)b0), l1 = lg2_fast((float)b1), l2 = lg2_fast((float)b2), l3 = lg2_fast((float)b3);
                    Kw[s] = fmaf((float)n.x, l0, fmaf((float)n.y, l1, fmaf((float)n.z, l2, (float)n.w * l3)));
                    mlP = fmaxf(mlP, fmaxf(fmaxf(fabsf(l0), fabsf(l1)), fmaxf(fabsf(l2), fabsf(l3))));
                }
                code = code_set(code, g, t);
                flips++;
                if (p.tau_cnt && lane == 0) {
                    const size_t vg = (size_t)v * G + g;
                    p.tau_cnt[vg * 4 + cur] += p.iter - p.tau_last[vg];
                    p.tau_last[vg] = p.iter;
                }
            }
        }
        KP_T(kp2);
        if (code != code_in) {
            if (lane < G) p.tau[(size_t)v * G + lane] = (uint8_t)code_get(code, lane);
            if (p.agg.N) {
                const int sn = agg_move_site(p.agg, code_in, code, tile, lane);
                if (p.site_slot && lane == 0) {
                    p.site_slot[v] = sn; atomicAdd(p.gctl + GC_ORPHANS, 1);
                    if (p.site_row) { const int r = p.site_row[v]; if (r >= 0) p.img_site[r] = ~v; }
                }
            }
        }

        __syncwarp();
#ifdef KPROF
        { const unsigned long long kp3 = gtimer(); if (!kp_sites) kp_first = kp0; kp_sites++; kp_stage += kp1 - kp0; kp_steps += kp2 - kp1; kp_move += kp3 - kp2; }
#endif
    }
#ifdef KPROF
    if (lane == 0 && blockIdx.x % 8 == 0) krec_put(KP_TAU_WARP,   /* a sample of the CTAs: the records cost atomics at the end */ (int)blockIdx.x, wib, (int)(kp_sites | (n2 << 8) | (n3 << 20) | (flips << 24)), kp_pro, gtimer(), kp_first, kp_stage, kp_steps, kp_move);
#endif

    if (lane == 0 && flips) atomicAdd(p.nchange, (unsigned long long)flips);
    if (lane == 0 && p.tier_counts) {
        if (n1) atomicAdd(p.tier_counts + 0, (unsigned long long)n1);
        if (n2) atomicAdd(p.tier_counts + 1, (unsigned long long)n2);
        if (n3) atomicAdd(p.tier_counts + 2, (unsigned long long)n3);
    }
}

static inline size_t tau_smem_bytes(int S, int G)
{
    const size_t Sp = (size_t)((S + 31) & ~31);
    return sizeof(double) * ((size_t)G * Sp + 32 + (size_t)TAU_WARPS * Sp * 4) + sizeof(float) * ((size_t)G * Sp + 16) +
           (sizeof(float) + sizeof(int4)) * (size_t)TAU_WARPS * Sp + sizeof(uint32_t) * TAU_WARPS * 32;
}
