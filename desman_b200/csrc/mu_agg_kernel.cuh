// desman_b200/csrc/mu_agg_kernel.cuh -- K2b: mu/E sufficient statistics, pattern-aggregated form
// (replaces HaploSNP_Sampler.sampleMu, HaploSNP_Sampler.py:284-309, and the reductions at :266, :276).
//
// The per-read kernel (mu_kernel.cuh) spends one Philox word and G-1 compares on each of the ~V*S*depth reads
// of a sweep (6.4e8 at BASELINE config C3) and is bound by integer issue.  But the category probabilities of a
// read observed as base a in sample s depend on the site only through its haplotype pattern tau_v: reads of all
// sites with the same pattern are exchangeable, and the sum of their multinomials is ONE multinomial with the summed
// count.  So:
//   aggregation pass      (table_maintain_kernel) one warp per site: hash the 2G-bit pattern code into a slot (open addressing, atomicCAS),
//                         add the site's S count cells into N[slot][s][a] (64-bit reductions); one HBM pass, needed only
//                         after a state upload: afterwards the tau kernel moves the counts of the sites it flips
//   mu_binomial_kernel    one warp per (slot, 32-sample chunk): per (s,a) a chain of conditional binomials over the
//                         strains (what numpy's RandomState.multinomial does), each by inversion when
//                         n*min(p,q) < 10 and by Hoermann's BTRS transformed rejection otherwise: O(1) per draw,
//                         independent of the count
// In a converged chain of biallelic sites the number of patterns is ~12*2^G (3e3 at G=8) << V, so the work drops
// by the average number of sites per pattern; with all-distinct patterns it degrades to one chain per cell.
// Draw contract: identical, operation for operation, to oracle_mu_stats_agg (oracle/desman_oracle.c); every
// decision outside the three logs of the BTRS slow path is made with +,*,/,sqrt,floor in IEEE double without
// contraction, so the integer statistics are reproducible between CPU and GPU.
#pragma once
#include "common.cuh"

#define STAGE_MUB 7
#define STAGE_MUC 8
#define MUC_MAX_G 16              // up to here the within-class splits are merged over all patterns sharing the class (dense table)
#define STAGE_MUC 8
#define MUC_MAX_G 16              // up to here the within-class splits are merged over all patterns sharing the class (dense table)
#ifndef MUB_WARPS
#define MUB_WARPS 8
#endif
#define MUB_EMPTY 0xffffffffffffffffull
#define AGG_CTL_WORDS 16          // ctl[4 + chunk]: work cursor of mu_binomial_kernel for sample chunk `chunk` (< MUB_CURSORS)
#define MUB_CURSORS 12
#ifndef MUB_MIN_BLOCKS
#define MUB_MIN_BLOCKS 2
#endif

// Persistent pattern table.  N[slot][s][a] = sum of the counts of all sites whose haplotype pattern is slot_code[slot].
// Built by mu_aggregate_kernel, then kept current by the tau kernel (a site that changes pattern moves its counts),
// so that in steady state (a handful of flips per sweep) no aggregation pass is needed.  ctl[0] = rebuild wanted
// (host on state upload; finalize_sweep when stale slots pile up; served and cleared by table_maintain_kernel,
// maintain_kernel.cuh), ctl[1] unused, ctl[2] = overflow, ctl[3] = flips since the last rebuild (upper bound of the
// number of stale slots).
struct AggTable {
    unsigned long long *keys;        // [H] pattern codes (MUB_EMPTY = free)
    int *ids;                        // [H] slot id of the key (-1 until published)
    unsigned int hmask;              // H - 1
    unsigned long long *slot_code;   // [cap_slots]
    unsigned int *nslots;            // slots handed out so far
    unsigned long long *N;           // [cap_slots][S][4]
    unsigned int cap_slots;
    int S;
    int *ctl;
};

struct MuAggParams {
    const int4 *counts;          // [V][S]
    const uint8_t *tau;          // [V][G]
    const double *gamma;         // [S][G]
    const double *eta;           // [16]
    uint64_t seed;
    uint32_t sweep;
    uint32_t shard;              // global index of local site 0 (keys the streams of a rank's partial aggregates)
    int V, S, G;
    AggTable t;
    unsigned long long *sum_mu;  // [S][G] +=
    unsigned long long *esum;    // [16]   += (esum[a_obs*4 + b_true])
    unsigned long long *classM;  // [2^G][S] reads per (set of strains, sample) awaiting their within-class split (G <= MUC_MAX_G), or null
    double ll_scale;             // 2^k of the fixed-point log-likelihood accumulator
    unsigned long long *ll_fx;   // += llrint(sum n*log p * 2^k)  (two's complement)
    double *eta_commit;          // ll_table_kernel: if non-null, eta_commit[0..15] = eta (the chain's eta <- eta_new, :347)
    // Upkeep wishes of this rank, formed by ll_table_kernel at the end of a sweep and written to ll_fx[2] (the word travels
    // with [ll, nchange] through the per-sweep exchange, so that every rank of a sharded chain regroups / rebuilds its table in
    // the SAME sweep: unsynchronised, some rank was regrouping in every other sweep at 8 ranks and all the others waited for it):
    // + 1 if the orphans of the site groups piled up, + 65536 if the pattern table wants a rebuild.  upkeep = 0: none.
    int upkeep;
    int *gctl_u;                 // site-group control words or null
    long long V_local_u;
    unsigned int agg_limit_u;
};

__device__ __forceinline__ unsigned int mix_code(unsigned long long x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return (unsigned int)x;
}

// Slot of a pattern code (one thread per call).  insert = false: the code is known to be present.
__device__ __forceinline__ int agg_slot(const AggTable &t, unsigned long long code, bool insert)
{
    unsigned int h = mix_code(code) & t.hmask;
    int id;
    while (true) {
        const unsigned long long prev = insert ? atomicCAS(t.keys + h, MUB_EMPTY, code) : *((volatile unsigned long long *)(t.keys + h));
        if (insert && prev == MUB_EMPTY) {                  // first site of this pattern: publish a new slot
            id = (int)atomicAdd(t.nslots, 1u);
            if ((unsigned int)id >= t.cap_slots) { t.ctl[2] = 1; id = (int)t.cap_slots - 1; }
            t.slot_code[id] = code;
            __threadfence();
            atomicExch(t.ids + h, id);
            return id;
        }
        if (prev == code) {                                 // known pattern: wait until its slot id is visible
            while ((id = *((volatile int *)(t.ids + h))) < 0) {}
            return id;
        }
        if (!insert && prev == MUB_EMPTY) { t.ctl[2] = 1; return 0; }   // inconsistent table (never expected)
        h = (h + 1) & t.hmask;
    }
}

// Move one site's counts between patterns (called by all lanes of the warp that owns the site).
__device__ __forceinline__ int agg_move_site(const AggTable &t, unsigned long long code_old, unsigned long long code_new,
                                             const int4 *tile, int lane)
{
    int so = 0, sn = 0;
    if (lane == 0) { so = agg_slot(t, code_old, false); sn = agg_slot(t, code_new, true); }
    so = __shfl_sync(DESMAN_FULL_MASK, so, 0);
    sn = __shfl_sync(DESMAN_FULL_MASK, sn, 0);
    unsigned long long *po = t.N + (size_t)so * t.S * 4, *pn = t.N + (size_t)sn * t.S * 4;
    for (int s = lane; s < t.S; s += 32) {
        const int4 n = tile[s];
        if (n.x) { atomicAdd(pn + s * 4 + 0, (unsigned long long)n.x); atomicAdd(po + s * 4 + 0, 0ull - (unsigned long long)n.x); }
        if (n.y) { atomicAdd(pn + s * 4 + 1, (unsigned long long)n.y); atomicAdd(po + s * 4 + 1, 0ull - (unsigned long long)n.y); }
        if (n.z) { atomicAdd(pn + s * 4 + 2, (unsigned long long)n.z); atomicAdd(po + s * 4 + 2, 0ull - (unsigned long long)n.z); }
        if (n.w) { atomicAdd(pn + s * 4 + 3, (unsigned long long)n.w); atomicAdd(po + s * 4 + 3, 0ull - (unsigned long long)n.w); }
    }
    return sn;
}

// K4 on the table: sum_v sum_s sum_b n*log p = sum_slots sum_s sum_a N[slot][s][a]*log(sum_g gamma[s,g]*eta[tau_g,a])
// (HaploSNP_Sampler.py:435,441).  Per (slot, chunk) the terms are summed in FP64 in a fixed order; the per-item sums are
// accumulated in 64-bit fixed point, so the total does not depend on slot numbering or scheduling (bitwise reproducible).
__global__ void __launch_bounds__(256) ll_table_kernel(MuAggParams p)
{
    pdl_enter();
    KPROF_SCOPE(KP_LL);
    __shared__ double eta_s[16];
    if (threadIdx.x < 16) {
        eta_s[threadIdx.x] = p.eta[threadIdx.x];
        if (blockIdx.x == 0 && p.eta_commit) p.eta_commit[threadIdx.x] = eta_s[threadIdx.x];
    }
    if (p.upkeep && blockIdx.x == 0 && threadIdx.x == 32) {
        // every flipped (v,g) of this sweep (ll_fx[1], still this rank's own count) may have left a stale slot behind: ask for a
        // rebuild when the stale slots could outnumber half the live ones, or when the slot array is close to its capacity
        // bound; ask for a regroup when the orphans piled up
        const long long stale = (long long)p.t.ctl[3] + (long long)p.ll_fx[1], used = (long long)*p.t.nslots;
        const long long live = used > stale ? used - stale : 0;
        p.t.ctl[3] = (int)(stale > 0x3fffffff ? 0x3fffffff : stale);
        unsigned long long wish = (stale > live / 2 + 512 || used > (long long)p.agg_limit_u) ? 65536ull : 0ull;
        if (p.gctl_u) {
            const long long lim = p.V_local_u / 128 > 64 ? p.V_local_u / 128 : 64;
            if ((long long)p.gctl_u[GC_ORPHANS] > lim) wish += 1ull;
        }
        p.ll_fx[2] = wish;
    }
    __syncthreads();
    const int S = p.S, G = p.G, lane = threadIdx.x & 31;
    const int nch = (S + 31) >> 5;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    unsigned int P = *p.t.nslots;
    if (P > p.t.cap_slots) P = p.t.cap_slots;
    long long fx = 0;
    for (long long item = gw; item < (long long)P * nch; item += nw) {
        const int slot = (int)(item / nch), s = (int)(item % nch) * 32 + lane;
        const unsigned long long code = p.t.slot_code[slot];
        double acc = 0.0;
        if (s < S) {
            const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(p.t.N + ((size_t)slot * S + s) * 4);
            const ulonglong2 lo = src[0], hi = src[1];
            if (lo.x | lo.y | hi.x | hi.y) {
                double b0 = 0.0, b1 = 0.0, b2 = 0.0, b3 = 0.0;
                for (int g = 0; g < G; g++) {
                    const double *e = eta_s + 4 * code_get(code, g);
                    const double gm = p.gamma[(size_t)s * G + g];
                    b0 = fma(e[0], gm, b0); b1 = fma(e[1], gm, b1); b2 = fma(e[2], gm, b2); b3 = fma(e[3], gm, b3);
                }
                if (lo.x) acc = fma((double)lo.x, log(b0), acc);
                if (lo.y) acc = fma((double)lo.y, log(b1), acc);
                if (hi.x) acc = fma((double)hi.x, log(b2), acc);
                if (hi.y) acc = fma((double)hi.y, log(b3), acc);
            }
        }
        acc = warp_sum(acc);
        fx += __double2ll_rn(acc * p.ll_scale);
    }
    if (lane == 0 && fx) atomicAdd(p.ll_fx, (unsigned long long)fx);
}

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double mub_u53(uint32_t hi, uint32_t lo)
{
    const unsigned long long m = ((unsigned long long)(hi >> 5) << 26) | (unsigned long long)(lo >> 6);
    return __dmul_rn(__dadd_rn((double)m, 0.5), 1.0 / 9007199254740992.0);
}

__device__ double stirling_tail_d(double k)
{
    if (k <= 9.0) {
        const int i = (int)k;
        return i == 0 ? 0.0810614667953272 : i == 1 ? 0.0413406959554092 : i == 2 ? 0.0276779256849983 :
               i == 3 ? 0.02079067210376509 : i == 4 ? 0.0166446911898211 : i == 5 ? 0.0138761288230707 :
               i == 6 ? 0.0118967099458917 : i == 7 ? 0.0104112652619720 : i == 8 ? 0.00925546218271273 :
                        0.00833056343336287;
    }
    const double t = __ddiv_rn(1.0, __dadd_rn(k, 1.0)), t2 = __dmul_rn(t, t);
    return __dmul_rn(__dadd_rn(1.0 / 12.0, -__dmul_rn(__dadd_rn(1.0 / 360.0, -__dmul_rn(1.0 / 1260.0, t2)), t2)), t);
}

struct BinStream { uint32_t c0, c1, c2, c3, k0, k1; };

// 1/x for x = 1..64, correctly rounded (filled by the host with 1.0 / x in IEEE double)
__constant__ double c_inv_small[65];

__device__ __forceinline__ void bin_uniforms_d(const BinStream &st, int g, uint32_t attempt, double &u1, double &u2)
{
    const uint4 o = philox4x32_10(st.c0, st.c1, st.c2, st.c3, st.k0 ^ (((uint32_t)(g + 1) << 20) | attempt), st.k1);
    u1 = mub_u53(o.x, o.y);
    u2 = mub_u53(o.z, o.w);
}

// Bin(n, p) with q = 1-p supplied separately.  Mirrors binomial_draw() of oracle/desman_oracle.c step for step.
__device__ __noinline__ long long binomial_draw_d(long long n, double p, double q, const BinStream &st, int g)
{
    if (n <= 0 || !(p > 0.0)) return 0;
    if (!(q > 0.0)) return n;
    const bool flip = p > q;
    const double pp = flip ? q : p, qq = flip ? p : q;
    const double dn = (double)n;
    long long x;
    if (__dmul_rn(dn, pp) < 10.0) {
        double r = 1.0, base = qq;
        for (long long e = n; e; e >>= 1) { if (e & 1) r = __dmul_rn(r, base); base = __dmul_rn(base, base); }
        const double s = __ddiv_rn(pp, qq);
        double u, dummy;
        bin_uniforms_d(st, g, 0u, u, dummy);
        x = 0;
        while (u >= r) {
            u = __dadd_rn(u, -r);
            x++;
            if (x > n) { x = n; break; }
            // r_x = r_{x-1} * s * (n-x+1) / x; for x <= 64 the division is a multiplication by the correctly rounded 1/x
            // (a table; the oracle forms the same 1.0/x), which takes ~80 cycles out of every step of the search
            const double num = __dmul_rn(r, __dmul_rn(s, (double)(n - x + 1)));
            r = (x <= 64) ? __dmul_rn(num, c_inv_small[x]) : __ddiv_rn(num, (double)x);
            if (x > 4096) break;
        }
    } else {
        const double spq = __dsqrt_rn(__dmul_rn(__dmul_rn(dn, pp), qq));
        const double b = __dadd_rn(1.15, __dmul_rn(2.53, spq));
        const double a = __dadd_rn(__dadd_rn(-0.0873, __dmul_rn(0.0248, b)), __dmul_rn(0.01, pp));
        const double c = __dadd_rn(__dmul_rn(dn, pp), 0.5);
        const double vr = __dadd_rn(0.92, -__ddiv_rn(4.2, b));
        const double r = __ddiv_rn(pp, qq);
        const double alpha = __dmul_rn(__dadd_rn(2.83, __ddiv_rn(5.1, b)), spq);
        const double m = floor(__dmul_rn(__dadd_rn(dn, 1.0), pp));
        double k = m;
        for (uint32_t t = 0; t < (1u << 20); t++) {
            double u1, v;
            bin_uniforms_d(st, g, t, u1, v);
            const double u = __dadd_rn(u1, -0.5);
            const double us = __dadd_rn(0.5, -fabs(u));
            k = floor(__dadd_rn(__dmul_rn(__dadd_rn(__ddiv_rn(__dmul_rn(2.0, a), us), b), u), c));
            if (us >= 0.07 && v <= vr) break;
            if (k < 0.0 || k > dn) continue;
            const double lv = log(__ddiv_rn(__dmul_rn(v, alpha), __dadd_rn(__ddiv_rn(a, __dmul_rn(us, us)), b)));
            const double nm1 = __dadd_rn(__dadd_rn(dn, -m), 1.0), nk1 = __dadd_rn(__dadd_rn(dn, -k), 1.0);
            double ub = __dmul_rn(__dadd_rn(m, 0.5), log(__ddiv_rn(__dadd_rn(m, 1.0), __dmul_rn(r, nm1))));
            ub = __dadd_rn(ub, __dmul_rn(__dadd_rn(dn, 1.0), log(__ddiv_rn(nm1, nk1))));
            ub = __dadd_rn(ub, __dmul_rn(__dadd_rn(k, 0.5), log(__ddiv_rn(__dmul_rn(r, nk1), __dadd_rn(k, 1.0)))));
            ub = __dadd_rn(ub, stirling_tail_d(m));
            ub = __dadd_rn(ub, stirling_tail_d(__dadd_rn(dn, -m)));
            ub = __dadd_rn(ub, -stirling_tail_d(k));
            ub = __dadd_rn(ub, -stirling_tail_d(__dadd_rn(dn, -k)));
            if (lv <= ub) break;
        }
        if (k < 0.0) k = 0.0;
        if (k > dn) k = dn;
        x = (long long)k;
    }
    return flip ? n - x : x;
}

// shared memory per warp: w[G][32], suf[G+1][32] (double), acc[G][32], E[16][32] (unsigned long long)
static inline size_t mub_smem_bytes(int G) { return (size_t)MUB_WARPS * 32 * 8 * (size_t)(3 * G + 1 + 16) + 16 * 8; }

// One work item: the pattern `slot`, this lane's sample s.  The multinomial of a cell (slot, s, a) over the strains, weights
// gamma[s,g]*eta[tau_g,a], factorises exactly over the CLASSES of the pattern (class b = the strains with tau_g == b): first
// the reads are split over the classes with weights eta[b,a]*Gamma_b (Gamma_b = sum of gamma over the class) -- these are the
// E statistics (HaploSNP_Sampler.py:301) -- and then, inside a class, over its strains with weights gamma[s,g], which do not
// depend on the observed base a: the four class totals are merged before that second split (:309 sums over a anyway).
// A biallelic pattern costs 4 + (G - 2) binomial draws per (slot, s) instead of 4 (G - 1).
//   phase A  cell (slot,s,a), classes present in ascending b: X_b ~ Bin(rem, W_b/suf, suf'/suf), W_b = eta[b,a]*Gamma_b, stream
//            ctr = (code, sweep, STAGE_MUB<<28 | a<<26 | s), draw index b
//   phase B  class b with M_b = sum_a X reads, strains ascending: X_g ~ Bin(rem, gamma_g/suf, suf'/suf), stream
//            ctr = (code, sweep, STAGE_MUC<<28 | b<<26 | s), draw index g
// The weights of phase B depend on the class only through its SET of strains, so for G <= MUC_MAX_G the class totals of all
// patterns are first merged per (set, sample) in classM and split once (mu_class_kernel: ctr = (set, 0, sweep, STAGE_MUC<<28 | s));
// a one-strain class needs no draw.  Identical, operation for operation, to oracle_mu_stats_agg.
__device__ __forceinline__ void mub_item(const MuAggParams &p, BinStream &st, int slot, bool valid, int s, int lane, int G,
                                         const double *eta_s, double *wS, double *sufS, unsigned long long *accS,
                                         unsigned long long *eS)
{
    const int S = p.S;
    const unsigned long long code = p.t.slot_code[slot];
    st.c0 = (uint32_t)code; st.c1 = (uint32_t)(code >> 32);
    long long n[4] = {0, 0, 0, 0};
    if (valid) {
        const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(p.t.N + ((size_t)slot * S + s) * 4);
        const ulonglong2 lo = src[0], hi = src[1];
        n[0] = (long long)lo.x; n[1] = (long long)lo.y; n[2] = (long long)hi.x; n[3] = (long long)hi.y;
    }
    if ((n[0] | n[1] | n[2] | n[3]) <= 0) return;
    // class masks (warp-uniform) and class abundances Gamma_b (ascending g, rounded adds)
    uint32_t cmask[4] = {0u, 0u, 0u, 0u};
    double Gm[4] = {0.0, 0.0, 0.0, 0.0};
    for (int g = 0; g < G; g++) {
        const int b = code_get(code, g);
        const double gm = p.gamma[(size_t)s * G + g];
        wS[g * 32 + lane] = gm;
#pragma unroll
        for (int k = 0; k < 4; k++) if (b == k) { cmask[k] |= 1u << g; Gm[k] = __dadd_rn(Gm[k], gm); }
    }
    const int lastk = cmask[3] ? 3 : cmask[2] ? 2 : cmask[1] ? 1 : 0;
    long long M[4] = {0, 0, 0, 0};

    // ---- phase A
#pragma unroll
    for (int a = 0; a < 4; a++) {
        if (n[a] <= 0) continue;
        st.c3 = ((uint32_t)STAGE_MUB << 28) | ((uint32_t)a << 26) | (uint32_t)s;
        double W[4], suf[5];
        suf[4] = 0.0;
#pragma unroll
        for (int k = 3; k >= 0; k--) {
            W[k] = cmask[k] ? __dmul_rn(eta_s[4 * k + a], Gm[k]) : 0.0;
            suf[k] = cmask[k] ? __dadd_rn(W[k], suf[k + 1]) : suf[k + 1];
        }
        long long rem = n[a];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (!cmask[k]) continue;
            const bool last = (k == lastk);
            long long x;
            if (last) x = rem;
            else if (rem == 0) x = 0;
            else x = binomial_draw_d(rem, __ddiv_rn(W[k], suf[k]), __ddiv_rn(suf[k + 1], suf[k]), st, k);
            rem -= x;
            if (x) { M[k] += x; eS[(a * 4 + k) * 32 + lane] += (unsigned long long)x; }
        }
    }
    // ---- phase B
    if (p.classM) {                                            // merged over patterns: mu_class_kernel does the split
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (!cmask[k] || M[k] <= 0) continue;
            if ((cmask[k] & (cmask[k] - 1u)) == 0u) accS[(31 - __clz(cmask[k])) * 32 + lane] += (unsigned long long)M[k];
            else atomicAdd(p.classM + (size_t)cmask[k] * S + s, (unsigned long long)M[k]);
        }
        return;
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (!cmask[k] || M[k] <= 0) continue;
        st.c3 = ((uint32_t)STAGE_MUC << 28) | ((uint32_t)k << 26) | (uint32_t)s;
        const int gl = 31 - __clz(cmask[k]);                 // last strain of the class
        double suf = 0.0;
        for (int g = gl; g >= 0; g--)
            if ((cmask[k] >> g) & 1u) { suf = __dadd_rn(wS[g * 32 + lane], suf); sufS[g * 32 + lane] = suf; }
        long long rem = M[k];
        double sg = suf;                                      // suffix sum at the current strain
        for (int g = 0; g <= gl; g++) {
            if (!((cmask[k] >> g) & 1u)) continue;
            long long x;
            if (g == gl) x = rem;
            else {
                // suffix sum after this strain = suffix sum at the next strain of the class
                const uint32_t higher = cmask[k] & ~((2u << g) - 1u);
                const double sn = sufS[(__ffs(higher) - 1) * 32 + lane];
                x = (rem == 0) ? 0 : binomial_draw_d(rem, __ddiv_rn(wS[g * 32 + lane], sg), __ddiv_rn(sn, sg), st, g);
                sg = sn;
            }
            rem -= x;
            if (x) accS[g * 32 + lane] += (unsigned long long)x;
        }
    }
}

// The same phase A for ONE observed base a of (slot, s), n = N[slot][s][a] reads (merged phase B only): the class totals of the
// four bases meet in classM / accS by addition, so the bases of a (slot, sample) are independent work items -- one draw per lane
// and item instead of four in a row.  Operation for operation the arithmetic of mub_item.
__device__ __forceinline__ void mub_item_base(const MuAggParams &p, BinStream &st, unsigned long long code, int a, long long n, int s,
                                              int lane, int G, const double *eta_s, unsigned long long *accS, unsigned long long *eS)
{
    if (n <= 0) return;
    const int S = p.S;
    st.c0 = (uint32_t)code; st.c1 = (uint32_t)(code >> 32);
    uint32_t cmask[4] = {0u, 0u, 0u, 0u};
    double Gm[4] = {0.0, 0.0, 0.0, 0.0};
    for (int g = 0; g < G; g++) {
        const int b = code_get(code, g);
        const double gm = p.gamma[(size_t)s * G + g];
#pragma unroll
        for (int k = 0; k < 4; k++) if (b == k) { cmask[k] |= 1u << g; Gm[k] = __dadd_rn(Gm[k], gm); }
    }
    const int lastk = cmask[3] ? 3 : cmask[2] ? 2 : cmask[1] ? 1 : 0;
    st.c3 = ((uint32_t)STAGE_MUB << 28) | ((uint32_t)a << 26) | (uint32_t)s;
    double W[4], suf[5];
    suf[4] = 0.0;
#pragma unroll
    for (int k = 3; k >= 0; k--) {
        W[k] = cmask[k] ? __dmul_rn(eta_s[4 * k + a], Gm[k]) : 0.0;
        suf[k] = cmask[k] ? __dadd_rn(W[k], suf[k + 1]) : suf[k + 1];
    }
    long long rem = n;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (!cmask[k]) continue;
        long long x;
        if (k == lastk) x = rem;
        else if (rem == 0) x = 0;
        else x = binomial_draw_d(rem, __ddiv_rn(W[k], suf[k]), __ddiv_rn(suf[k + 1], suf[k]), st, k);
        rem -= x;
        if (x) {
            eS[(a * 4 + k) * 32 + lane] += (unsigned long long)x;
            if ((cmask[k] & (cmask[k] - 1u)) == 0u) accS[(31 - __clz(cmask[k])) * 32 + lane] += (unsigned long long)x;
            else atomicAdd(p.classM + (size_t)cmask[k] * S + s, (unsigned long long)x);
        }
    }
}

__global__ void __launch_bounds__(MUB_WARPS * 32, MUB_MIN_BLOCKS) mu_binomial_kernel(MuAggParams p)
{
#if !PDL_EARLY
    pdl_enter();
#endif
    KPROF_SCOPE(KP_MUB);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int S = p.S, G = p.G;
    double *eta_s = reinterpret_cast<double *>(smem_raw);                    // [16]
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    double *wS = eta_s + 16 + (size_t)wib * 32 * (3 * G + 1 + 16);            // [G][32]
    double *sufS = wS + G * 32;                                              // [G+1][32]
    unsigned long long *accS = reinterpret_cast<unsigned long long *>(sufS + (G + 1) * 32);   // [G][32]
    unsigned long long *eS = accS + G * 32;                                  // [16][32]
    // (eta: committed by the finalize kernel of the previous sweep, several grids ago -- read under the predecessor's tail)
    if (threadIdx.x < 16) eta_s[threadIdx.x] = p.eta[threadIdx.x];
    for (int i = lane; i < (G + 16) * 32; i += 32) accS[i] = 0ull;
#if PDL_EARLY
    pdl_enter();
#endif
    __syncthreads();

    const int nch = (S + 31) >> 5;
    const int gw = blockIdx.x * MUB_WARPS + wib, nw = gridDim.x * MUB_WARPS;      // nw % nch == 0 (host)
    const int chunk = gw % nch, s = chunk * 32 + lane;
    const bool valid = s < S;
    unsigned int Pu = *p.t.nslots;
    if (Pu > p.t.cap_slots) Pu = p.t.cap_slots;
    const int P = (int)Pu;
    BinStream st;
    st.c2 = p.sweep; st.k0 = (uint32_t)p.seed; st.k1 = (uint32_t)(p.seed >> 32) ^ p.shard;

    // work item = pattern slot, for this warp's sample chunk.  Items are handed out dynamically (one cursor per sample
    // chunk, cleared by table_maintain_kernel at the start of the sweep): their cost varies with the counts (inversion vs.
    // rejection, number of attempts), and a static deal left a quarter of the SM time idle at the tail.  The statistics are
    // integer sums, so the schedule does not change the result.
    const bool dyn = nch <= MUB_CURSORS;
    if (dyn && p.classM) {
        // Merged within-class split: an item is (slot, observed base) -- a lane's four draws of a slot were a sequential chain of
        // ~10 us and a warp saw 1.3 such items per sweep at C3 (first CTA done at 19 us, last at 34 us, tools/kprof.py).  The
        // ticket -> code -> count chain of an item (three dependent memory latencies, ~3 us) is taken off the critical path: the
        // ticket of item i+2 is drawn and the operands of item i+1 are fetched before item i is processed.
        const int nitem = 4 * P;
        // the first item of a warp is its own number within the chunk (no atomic storm at the start); tickets follow from there
        const int first = gw / nch, nfirst = nw / nch;
        auto draw_ticket = [&]() { int t = 0; if (lane == 0) t = nfirst + atomicAdd(p.t.ctl + 4 + chunk, 1); return t; };
        auto fetch = [&](int it, unsigned long long &code, long long &n) {
            code = 0ull; n = 0;
            if (it < nitem) {
                code = p.t.slot_code[it >> 2];
                if (valid) n = (long long)p.t.N[((size_t)(it >> 2) * S + s) * 4 + (it & 3)];
            }
        };
        int it1 = first;
        int raw2 = draw_ticket();
        unsigned long long code1; long long n1;
        fetch(it1, code1, n1);
#ifdef KPROF
        unsigned long long kp_items = 0; const unsigned long long kp_t0 = gtimer();
#endif
        while (it1 < nitem) {
            const int cur = it1;
            const unsigned long long code = code1;
            const long long n = n1;
            it1 = __shfl_sync(DESMAN_FULL_MASK, raw2, 0);
            fetch(it1, code1, n1);
            raw2 = draw_ticket();
            mub_item_base(p, st, code, cur & 3, n, s, lane, G, eta_s, accS, eS);
#ifdef KPROF
            kp_items++;
#endif
        }
#ifdef KPROF
        if (lane == 0 && blockIdx.x % 8 == 0) krec_put(KP_MUB_WARP, (int)blockIdx.x, wib, (int)kp_items, kp_t0, gtimer(), 0, 0, 0, 0);
#endif
    } else {
        int item = gw / nch;
        while (true) {
            if (dyn) {
                if (lane == 0) item = atomicAdd(p.t.ctl + 4 + chunk, 1);
                item = __shfl_sync(DESMAN_FULL_MASK, item, 0);
            }
            if (item >= P) break;
            const int item_cur = item;
            if (!dyn) item += nw / nch;
            mub_item(p, st, item_cur, valid, s, lane, G, eta_s, wS, sufS, accS, eS);
        }
    }
    // Flush: the accumulators of the CTA's warps are added up in shared memory first (row j of the (G + 16) x 32 accumulator
    // block of a warp: strain j < G, else E entry j - G) -- one global atomic per (CTA, sample chunk, row, lane) instead of one
    // per warp: 2368 warps flushing S*G + 16 words each put ~1200 (mu) and ~2400 (E) same-address atomics in a row in front of
    // the grid's completion (5-8 us between the last CTA's exit and the start of the dependent launch, tools/kprof.py)
    __syncthreads();
    {
        const size_t wstride = (size_t)32 * (3 * G + 1 + 16);                   // doubles (= 64-bit words) per warp region
        const unsigned long long *acc0 = reinterpret_cast<const unsigned long long *>(eta_s + 16 + (size_t)(2 * G + 1) * 32);
        const int nrow = G + 16;
        // warps with the same sample chunk: wib' = wib0, wib0 + nch, ...  (chunk = (blockIdx * MUB_WARPS + wib') % nch)
        for (int row = wib; row < nrow * min(nch, MUB_WARPS); row += MUB_WARPS) {
            const int r = row % nrow, k = row / nrow;                            // k-th distinct chunk of this CTA: warps k, k + nch, ...
            unsigned long long x = 0ull;
            for (int w = k; w < MUB_WARPS; w += nch) x += acc0[(size_t)w * wstride + (size_t)r * 32 + lane];
            const int ch = (int)(((long long)blockIdx.x * MUB_WARPS + k) % nch), s2 = ch * 32 + lane;
            if (r < G) {
                if (x && s2 < S) atomicAdd(p.sum_mu + (size_t)s2 * G + r, x);
            } else {
                x = warp_sum_u64(x);
                if (lane == 0 && x) atomicAdd(p.esum + (r - G), x);
            }
        }
    }
}

// Within-class split, merged over patterns (G <= MUC_MAX_G).  The M reads of the set `mask` at sample s are dealt to its strains
// by a balanced binary tree of binomial splits: node over the (ascending) strain positions [lo, hi), n = hi - lo >= 2, splits at
// mid = lo + (n+1)/2:  X_left ~ Bin(M_node, L/(L+R), R/(L+R)),  L, R = sums of gamma[s,g] over the halves (ascending, rounded
// adds).  The tree has depth <= 4 and the nodes of a level are independent, so a quad of lanes works on one sample and the latency
// is that of <= 4 draws (a chain over the strains costs up to 15).  Stream: ctr = (mask, 0, sweep, STAGE_MUC<<28 | s), draw index
// = heap number of the node - 1 (root 1, children 2i, 2i+1).  One warp per (mask, 8 samples); mirrors oracle_mu_stats_agg.
__global__ void __launch_bounds__(MUB_WARPS * 32, MUB_MIN_BLOCKS) mu_class_kernel(MuAggParams p)
{
    pdl_enter();
    KPROF_SCOPE(KP_MUC);
    __shared__ long long nodeM_s[MUB_WARPS][8][32];          // reads at the nodes of the current item, per sample
    __shared__ int pos2g_s[MUB_WARPS][32];
    const int S = p.S, G = p.G;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    long long (*nodeM)[32] = nodeM_s[wib];
    int *pos2g = pos2g_s[wib];
    const int gw = blockIdx.x * MUB_WARPS + wib, nw = gridDim.x * MUB_WARPS;
    const int sl = lane >> 2, q4 = lane & 3;                 // sample within the chunk, lane within the quad
    const int nch = (S + 7) >> 3;
    BinStream st;
    st.c1 = 0u; st.c2 = p.sweep; st.k0 = (uint32_t)p.seed; st.k1 = (uint32_t)(p.seed >> 32) ^ p.shard;
    const long long nitems = ((long long)1 << G) * nch;
    for (long long item = gw; item < nitems; item += nw) {
        const uint32_t mask = (uint32_t)(item / nch);
        if ((mask & (mask - 1u)) == 0u) continue;              // empty or one strain: settled in phase A
        const int s = (int)(item % nch) * 8 + sl;
        const bool valid = s < S;
        long long M = 0;
        if (valid && q4 == 0) {
            M = (long long)p.classM[(size_t)mask * S + s];
            if (M > 0) p.classM[(size_t)mask * S + s] = 0ull;   // consumed
        }
        if (!__any_sync(DESMAN_FULL_MASK, M > 0)) continue;
        const int m = __popc(mask);
        __syncwarp();
        if (lane < m) pos2g[lane] = (int)__fns(mask, 0, lane + 1);
        if (q4 == 0) nodeM[sl][1] = M;
        __syncwarp();
        st.c0 = mask;
        st.c3 = ((uint32_t)STAGE_MUC << 28) | (uint32_t)s;
        int depth = 0;
        while ((1 << depth) < m) depth++;
        for (int level = 0; level < depth; level++) {
            for (int id = (1 << level) + q4; id < (2 << level); id += 4) {
                // range of node `id`: descend from the root along the bits below the leading one
                int lo = 0, hi = m;
                for (int b = level - 1; b >= 0; b--) {
                    const int mid = lo + (hi - lo + 1) / 2;
                    if ((id >> b) & 1) lo = mid; else hi = mid;
                }
                const int n = hi - lo;
                if (n < 2 || !valid) continue;                  // (a range of one strain was credited by its parent)
                const long long Mn = nodeM[sl][id];
                const int mid = lo + (n + 1) / 2;
                long long xl = 0;
                if (Mn > 0) {
                    double L = 0.0, R = 0.0;
                    for (int i = lo; i < mid; i++) L = __dadd_rn(L, p.gamma[(size_t)s * G + pos2g[i]]);
                    for (int i = mid; i < hi; i++) R = __dadd_rn(R, p.gamma[(size_t)s * G + pos2g[i]]);
                    const double T = __dadd_rn(L, R);
                    xl = binomial_draw_d(Mn, __ddiv_rn(L, T), __ddiv_rn(R, T), st, id - 1);
                }
                const long long xr = Mn - xl;
                if (mid - lo >= 2) nodeM[sl][2 * id] = xl;
                else if (xl) atomicAdd(p.sum_mu + (size_t)s * G + pos2g[lo], (unsigned long long)xl);
                if (hi - mid >= 2) nodeM[sl][2 * id + 1] = xr;
                else if (xr) atomicAdd(p.sum_mu + (size_t)s * G + pos2g[mid], (unsigned long long)xr);
            }
            __syncwarp();
        }
    }
}
static inline size_t muc_smem_bytes(int G) { (void)G; return 0; }
