/* oracle/desman_oracle.c -- TEST INFRASTRUCTURE ONLY, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the DESMAN haplotype-inference hot path, used as the
 * parity checker for the CUDA engine and as the timed CPU baseline ("port").
 * Nothing under desman_b200/ may include, link or load this file.
 *
 * Build: oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp).  -ffp-contract=off is part
 * of the contract: the integer thresholds of oracle_mu_stats must round exactly
 * like the CUDA kernel's __dmul_rn/__dadd_rn/__ddiv_rn sequence.
 *
 * Reference citations are file:line into the reference checkout
 * (sampletau/c_sample_tau.c, desman/HaploSNP_Sampler.py, desman/Init_NMFT.py,
 * desman/Desman_Utils.py).
 */
#include "desman_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ======================================================================== */
/* RNG primitives                                                           */
/* ======================================================================== */

/* GSL gsl_rng_mt19937: gsl_rng_set (seed 0 -> 4357, init_genrand recurrence) as used at
 * c_sample_tau.c:33,39. */
void oracle_mt_seed(oracle_mt19937 *r, unsigned long seed)
{
    uint32_t s = (uint32_t)(seed & 0xffffffffUL);
    if (seed == 0) s = 4357u;
    r->mt[0] = s;
    for (int i = 1; i < 624; i++)
        r->mt[i] = 1812433253u * (r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) + (uint32_t)i;
    r->mti = 624;
}

static void mt_regenerate(oracle_mt19937 *r)
{
    uint32_t *m = r->mt;
    for (int k = 0; k < 624; k++) {
        uint32_t y = (m[k] & 0x80000000u) | (m[(k + 1) % 624] & 0x7fffffffu);
        m[k] = m[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    r->mti = 0;
}

uint32_t oracle_mt_next(oracle_mt19937 *r)
{
    if (r->mti >= 624) mt_regenerate(r);
    uint32_t y = r->mt[r->mti++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

void oracle_mt_fill(oracle_mt19937 *r, uint32_t *out, int64_t n)
{
    for (int64_t i = 0; i < n; i++) out[i] = oracle_mt_next(r);
}

void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int round = 0; round < 10; round++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static inline void philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint64_t seed, uint32_t out[4])
{
    uint32_t ctr[4] = {c0, c1, c2, c3};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    oracle_philox4x32_10(ctr, key, out);
}

/* ======================================================================== */
/* tau update  (c_sample_tau.c:48-204)                                      */
/* ======================================================================== */

/* normaliseLog4, c_sample_tau.c:48-70 */
static void softmax4(double *lp)
{
    double mx = lp[0], sum = 0.0;
    for (int b = 1; b < 4; b++) if (lp[b] > mx) mx = lp[b];
    for (int b = 0; b < 4; b++) { lp[b] = lp[b] - mx; sum += exp(lp[b]); }
    for (int b = 0; b < 4; b++) lp[b] = exp(lp[b]) / sum;
}

/* sample4, c_sample_tau.c:72-91: strict '<' against partial sums, hard c3 = 1 */
static int pick4(const double *p, double u)
{
    double c0 = p[0], c1 = p[1] + c0, c2 = p[2] + c1;
    if (u < c0) return 0;
    if (u < c1) return 1;
    if (u < c2) return 2;
    return 3;
}

/* One (v,g) step: c_sample_tau.c:136-170.  idx[h] is the current base of strain h.
 * Operation order: base accumulates over h ascending (skipping g) from 0.0 (:138-147);
 * the candidate term eta[a][b]*pi[s][g] is added last (:157); logp[a] accumulates s-major,
 * b-minor, with the count passed through float (:164). */
static void tau_step_logp(const int *idx, const double *pi, const double *eta,
                          const int64_t *nv, int G, int S, int g, double *scratch, double logp[4])
{
    double *base = scratch; /* [S][4] */
    for (int s = 0; s < S; s++)
        for (int b = 0; b < 4; b++) {
            double acc = 0.0;
            for (int h = 0; h < G; h++)
                if (h != g) acc += eta[idx[h] * 4 + b] * pi[s * G + h];
            base[s * 4 + b] = acc;
        }
    for (int a = 0; a < 4; a++) {
        double L = 0.0;
        for (int s = 0; s < S; s++)
            for (int b = 0; b < 4; b++) {
                double p = base[s * 4 + b];
                p += eta[a * 4 + b] * pi[s * G + g];
                double term = ((float)nv[s * 4 + b]) * log(p);
                L += term;
            }
        logp[a] = L;
    }
}

void oracle_tau_step_probs(const int64_t *tau_index_v, const double *pi, const double *eta,
                           const int64_t *variants_v, int G, int S, int g,
                           double logp[4], double prob[4])
{
    int *idx = (int *)malloc(sizeof(int) * (size_t)G);
    double *scratch = (double *)malloc(sizeof(double) * 4 * (size_t)S);
    for (int h = 0; h < G; h++) idx[h] = (int)tau_index_v[h];
    tau_step_logp(idx, pi, eta, variants_v, G, S, g, scratch, logp);
    for (int a = 0; a < 4; a++) prob[a] = logp[a];
    softmax4(prob);
    free(idx); free(scratch);
}

/* midpoint = 0: u = w / 2^32 (gsl_rng_uniform, c_sample_tau.c:174; u = 0 possible);
 * midpoint = 1: u = (w + 0.5) / 2^32 (the Philox contract of the GPU chain: u is never 0, so a step whose
 * current base holds all but < 2^-33 of the mass is decided without looking at the word). */
/* g_begin > 0 / logp_out != NULL: sampleTauFixTau (HaploSNP_Sampler.py:196-222): strains below g_begin keep their base, and
 * the normalised log-probabilities (normaliseLogProb, :186-194) of the four bases of strain g_begin are recorded per site
 * before its draw. */
static int sample_tau_words_impl2(int64_t *tau, const double *pi, const double *eta,
                                  const int64_t *variants, int V, int G, int S,
                                  const uint32_t *words, int midpoint, int g_begin, double *logp_out);
static int sample_tau_words_impl(int64_t *tau, const double *pi, const double *eta,
                                 const int64_t *variants, int V, int G, int S,
                                 const uint32_t *words, int midpoint)
{
    return sample_tau_words_impl2(tau, pi, eta, variants, V, G, S, words, midpoint, 0, NULL);
}
static int sample_tau_words_impl2(int64_t *tau, const double *pi, const double *eta,
                                  const int64_t *variants, int V, int G, int S,
                                  const uint32_t *words, int midpoint, int g_begin, double *logp_out)
{
    int nchange = 0;
#pragma omp parallel reduction(+ : nchange)
    {
        int *idx = (int *)malloc(sizeof(int) * (size_t)G);
        double *scratch = (double *)malloc(sizeof(double) * 4 * (size_t)S);
#pragma omp for schedule(static)
        for (int v = 0; v < V; v++) {
            int64_t *tv = tau + (size_t)v * G * 4;
            /* one-hot -> index, c_sample_tau.c:115-123 (first b with tau == 1) */
            for (int g = 0; g < G; g++) {
                idx[g] = 0;
                for (int b = 0; b < 4; b++) if (tv[g * 4 + b] == 1) { idx[g] = b; break; }
            }
            const int64_t *nv = variants + (size_t)v * S * 4;
            for (int g = g_begin; g < G; g++) {
                double p[4];
                tau_step_logp(idx, pi, eta, nv, G, S, g, scratch, p);
                if (logp_out && g == g_begin) {                             /* normaliseLogProb, HaploSNP_Sampler.py:186-194 */
                    double mx = p[0], sum = 0.0;
                    for (int b = 1; b < 4; b++) if (p[b] > mx) mx = p[b];
                    for (int b = 0; b < 4; b++) sum += exp(p[b] - mx);
                    for (int b = 0; b < 4; b++) logp_out[(size_t)v * 4 + b] = (p[b] - mx) - log(sum);
                }
                softmax4(p);                                                /* :172 */
                double u = midpoint ? ((double)words[(size_t)v * G + g] + 0.5) / 4294967296.0
                                    : words[(size_t)v * G + g] / 4294967296.0;  /* :174 gsl_rng_uniform */
                int t = pick4(p, u);                                        /* :176 */
                if (t != idx[g]) {                                          /* :178-185 */
                    tv[g * 4 + idx[g]] = 0;
                    tv[g * 4 + t] = 1;
                    idx[g] = t;
                    nchange++;
                }
            }
        }
        free(idx); free(scratch);
    }
    return nchange;
}

int oracle_sample_tau_words(int64_t *tau, const double *pi, const double *eta,
                            const int64_t *variants, int V, int G, int S,
                            const uint32_t *words)
{
    return sample_tau_words_impl(tau, pi, eta, variants, V, G, S, words, 0);
}

int oracle_sample_tau_mt(int64_t *tau, const double *pi, const double *eta,
                         const int64_t *variants, int V, int G, int S, oracle_mt19937 *rng)
{
    size_t n = (size_t)V * G;
    uint32_t *w = (uint32_t *)malloc(sizeof(uint32_t) * (n ? n : 1));
    oracle_mt_fill(rng, w, (int64_t)n);
    int r = oracle_sample_tau_words(tau, pi, eta, variants, V, G, S, w);
    free(w);
    return r;
}

int oracle_sample_tau_philox(int64_t *tau, const double *pi, const double *eta,
                             const int64_t *variants, int V, int G, int S,
                             uint64_t seed, uint32_t sweep, int64_t v0)
{
    size_t n = (size_t)V * G;
    uint32_t *w = (uint32_t *)malloc(sizeof(uint32_t) * (n ? n : 1));
#pragma omp parallel for schedule(static)
    for (int v = 0; v < V; v++)
        for (int g = 0; g < G; g++) {
            uint32_t o[4];
            philox((uint32_t)(v0 + v), (uint32_t)g, sweep, (uint32_t)ORACLE_STAGE_TAU << 28, seed, o);
            w[(size_t)v * G + g] = o[0];
        }
    int r = sample_tau_words_impl(tau, pi, eta, variants, V, G, S, w, 1);
    free(w);
    return r;
}

/* sampleTauFixTau (HaploSNP_Sampler.py:196-222) under the Philox contract of the tau draws: strains [H, G) are redrawn in
 * order with u = (w + 0.5) / 2^32, w = Philox(ctr = (v0 + v, g, sweep, STAGE_TAU << 28)).x, by sample4(softmax(L), u) exactly
 * like a tau step (the reference draws from numpy's sequential multinomial, which has no parallel form); logp_out[v][b] =
 * normalised log-probabilities of strain H's four bases before its draw. */
int oracle_sample_tau_fix_philox(int64_t *tau, const double *pi, const double *eta,
                                 const int64_t *variants, int V, int G, int S,
                                 uint64_t seed, uint32_t sweep, int64_t v0, int H, double *logp_out)
{
    size_t n = (size_t)V * G;
    uint32_t *w = (uint32_t *)malloc(sizeof(uint32_t) * (n ? n : 1));
#pragma omp parallel for schedule(static)
    for (int v = 0; v < V; v++)
        for (int g = 0; g < G; g++) {
            uint32_t o[4];
            philox((uint32_t)(v0 + v), (uint32_t)g, sweep, (uint32_t)ORACLE_STAGE_TAU << 28, seed, o);
            w[(size_t)v * G + g] = o[0];
        }
    int r = sample_tau_words_impl2(tau, pi, eta, variants, V, G, S, w, 1, H, logp_out);
    free(w);
    return r;
}

/* ======================================================================== */
/* mu / E sufficient statistics  (HaploSNP_Sampler.py:284-309)              */
/* ======================================================================== */
/* The reference draws E[v,s,a,:] ~ Mult(n_vsa, P(true b | obs a)) (:301) and then
 * mu[v,s,a,:] += Mult(E_vsab, P(g | b, a)) (:305-309); only sum_mu = mu.sum(axis=(0,2))
 * (:266) and Esum = E.sum(axis=(0,1)) (:276) are consumed.  With one-hot tau the joint law
 * of (b, g) for a read observed as a is P(g) ~ gamma[s,g]*eta[tau_vg,a], b = tau_vg.
 * Contract: read j of cell (v,s,a) uses word (j & 3) of
 *   Philox(ctr = (v, j >> 2, sweep, STAGE_MU<<28 | a<<26 | s), key = seed)
 * and picks strain  #{g < G-1 : word >= T_g},  T_g = min(floor(cum_g * (2^32 / cum_{G-1})), 2^32-1),
 * cum_g = sum_{h<=g} gamma[s,h]*eta[tau_vh,a] accumulated in ascending h (round-to-nearest,
 * no FMA contraction). */
void oracle_mu_stats(const int64_t *tau, const double *gamma, const double *eta,
                     const int64_t *variants, int V, int G, int S,
                     uint64_t seed, uint32_t sweep, int64_t v0,
                     int64_t *sum_mu, int64_t *esum)
{
#pragma omp parallel
    {
        int64_t *lmu = (int64_t *)calloc((size_t)S * G + 16, sizeof(int64_t));
        int64_t *le = lmu + (size_t)S * G;
        int *idx = (int *)malloc(sizeof(int) * (size_t)G);
        uint32_t *thr = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)G);
        int64_t *cnt = (int64_t *)malloc(sizeof(int64_t) * (size_t)G);
#pragma omp for schedule(static)
        for (int v = 0; v < V; v++) {
            const int64_t *tv = tau + (size_t)v * G * 4;
            for (int g = 0; g < G; g++) {
                idx[g] = 0;
                for (int b = 0; b < 4; b++) if (tv[g * 4 + b] == 1) { idx[g] = b; break; }
            }
            for (int s = 0; s < S; s++)
                for (int a = 0; a < 4; a++) {
                    int64_t n = variants[((size_t)v * S + s) * 4 + a];
                    if (n <= 0) continue;
                    double cum = 0.0;
                    double cums[64];
                    for (int g = 0; g < G; g++) {
                        double w = gamma[s * G + g] * eta[idx[g] * 4 + a];
                        cum = cum + w;
                        cums[g] = cum;
                    }
                    double scale = 4294967296.0 / cum;
                    for (int g = 0; g < G - 1; g++) {
                        double t = floor(cums[g] * scale);
                        thr[g] = (t >= 4294967295.0) ? 4294967295u : (uint32_t)t;
                    }
                    for (int g = 0; g < G; g++) cnt[g] = 0;
                    uint32_t c3 = ((uint32_t)ORACLE_STAGE_MU << 28) | ((uint32_t)a << 26) | (uint32_t)s;
                    for (int64_t j0 = 0; j0 < n; j0 += 4) {
                        uint32_t o[4];
                        philox((uint32_t)(v0 + v), (uint32_t)(j0 >> 2), sweep, c3, seed, o);
                        int lim = (n - j0 < 4) ? (int)(n - j0) : 4;
                        for (int k = 0; k < lim; k++) {
                            int cat = 0;
                            for (int g = 0; g < G - 1; g++) cat += (o[k] >= thr[g]);
                            cnt[cat]++;
                        }
                    }
                    for (int g = 0; g < G; g++) {
                        lmu[s * G + g] += cnt[g];
                        le[a * 4 + idx[g]] += cnt[g];
                    }
                }
        }
#pragma omp critical
        {
            for (size_t i = 0; i < (size_t)S * G; i++) sum_mu[i] += lmu[i];
            for (int i = 0; i < 16; i++) esum[i] += le[i];
        }
        free(lmu); free(idx); free(thr); free(cnt);
    }
}

/* ------------------------------------------------------------------------ */
/* Aggregated form of the same statistics ("pattern" contract).
 * Reads of cells (v,s,a) whose sites carry the SAME haplotype pattern tau_v have identical category
 * probabilities and are exchangeable, so their multinomials add up to ONE multinomial with the summed
 * count N[pattern][s][a].  That multinomial is drawn as a chain of conditional binomials in ascending g
 * (the construction numpy's RandomState.multinomial itself uses), each binomial by inversion when
 * n*min(p,q) < 10 and by Hoermann's BTRS transformed rejection otherwise.
 *   weights   w_g = gamma[s,g]*eta[tau_g,a]; suffix sums suf_g = w_g + suf_{g+1} (descending, rounded adds)
 *   draw g    X_g ~ Bin(n_rem, p = w_g/suf_g) with q = suf_{g+1}/suf_g;  X_{G-1} = n_rem
 *   uniforms  attempt t of draw (pattern code, s, a, g): Philox(ctr = (code_lo, code_hi, sweep,
 *             STAGE_MUB<<28 | a<<26 | s), key = (seed_lo ^ ((g+1)<<20 | t), seed_hi ^ shard)):
 *             U1 = u53(w0,w1), U2 = u53(w2,w3)
 * All arithmetic outside the three logs of the BTRS slow path is +,*,/,sqrt,floor in IEEE double without
 * contraction, identical in the CUDA kernel. */
#define ORACLE_STAGE_MUB 7
#define ORACLE_STAGE_MUC 8
#define ORACLE_MUC_MAX_G 16

static inline double u53w(uint32_t hi, uint32_t lo)
{
    uint64_t m = ((uint64_t)(hi >> 5) << 26) | (uint64_t)(lo >> 6);
    return ((double)m + 0.5) * (1.0 / 9007199254740992.0);
}

static double stirling_tail(double k)
{
    static const double t[10] = {0.0810614667953272, 0.0413406959554092, 0.0276779256849983, 0.02079067210376509,
                                 0.0166446911898211, 0.0138761288230707, 0.0118967099458917, 0.0104112652619720,
                                 0.00925546218271273, 0.00833056343336287};
    if (k <= 9.0) return t[(int)k];
    double r1 = 1.0 / (k + 1.0), r2 = r1 * r1;   /* one division; Horner in 1/(k+1)^2 */
    return (1.0 / 12.0 - (1.0 / 360.0 - (1.0 / 1260.0) * r2) * r2) * r1;
}

typedef struct { uint32_t c0, c1, c2, c3; uint64_t seed; uint32_t shard; int g; } bin_stream;

static void bin_uniforms(const bin_stream *st, uint32_t attempt, double *u1, double *u2)
{
    uint32_t ctr[4] = {st->c0, st->c1, st->c2, st->c3};
    uint32_t key[2] = {(uint32_t)st->seed ^ (((uint32_t)(st->g + 1) << 20) | attempt), (uint32_t)(st->seed >> 32) ^ st->shard};
    uint32_t o[4];
    oracle_philox4x32_10(ctr, key, o);
    *u1 = u53w(o[0], o[1]);
    *u2 = u53w(o[2], o[3]);
}

/* Bin(n, p) with q = 1-p supplied separately (no cancellation).  n >= 0. */
static int64_t binomial_draw(int64_t n, double p, double q, const bin_stream *st)
{
    if (n <= 0 || !(p > 0.0)) return 0;
    if (!(q > 0.0)) return n;
    const int flip = p > q;
    const double pp = flip ? q : p, qq = flip ? p : q;
    const double dn = (double)n;
    int64_t x;
    if (dn * pp < 10.0) {
        /* inversion: r = qq^n by binary powering, then sequential search */
        double r = 1.0, base = qq;
        for (int64_t e = n; e; e >>= 1) { if (e & 1) r = r * base; base = base * base; }
        const double s = pp / qq;
        double u, dummy;
        bin_uniforms(st, 0, &u, &dummy);
        x = 0;
        while (u >= r) {
            u = u - r;
            x++;
            if (x > n) { x = n; break; }
            {   /* for x <= 64 the GPU multiplies by the correctly rounded reciprocal instead of dividing */
                const double num = r * (s * (double)(n - x + 1));
                if (x <= 64) { const double inv = 1.0 / (double)x; r = num * inv; }
                else r = num / (double)x;
            }
            if (x > 4096) break;        /* numerical guard: mass beyond here is < 1e-300 */
        }
    } else {
        /* BTRS (Hoermann 1993) */
        const double spq = sqrt(dn * pp * qq);
        const double b = 1.15 + 2.53 * spq;
        const double a = -0.0873 + 0.0248 * b + 0.01 * pp;
        const double c = dn * pp + 0.5;
        const double vr = 0.92 - 4.2 / b;
        const double r = pp / qq;
        const double alpha = (2.83 + 5.1 / b) * spq;
        const double m = floor((dn + 1.0) * pp);
        double k = m;
        for (uint32_t t = 0; t < (1u << 20); t++) {
            double u1, v;
            bin_uniforms(st, t, &u1, &v);
            const double u = u1 - 0.5;
            const double us = 0.5 - fabs(u);
            k = floor((2.0 * a / us + b) * u + c);
            if (us >= 0.07 && v <= vr) break;
            if (k < 0.0 || k > dn) continue;
            const double lv = log(v * alpha / (a / (us * us) + b));
            const double ub = (m + 0.5) * log((m + 1.0) / (r * (dn - m + 1.0))) +
                              (dn + 1.0) * log((dn - m + 1.0) / (dn - k + 1.0)) +
                              (k + 0.5) * log(r * (dn - k + 1.0) / (k + 1.0)) +
                              stirling_tail(m) + stirling_tail(dn - m) - stirling_tail(k) - stirling_tail(dn - k);
            if (lv <= ub) break;
        }
        if (k < 0.0) k = 0.0;
        if (k > dn) k = dn;
        x = (int64_t)k;
    }
    return flip ? n - x : x;
}

static int cmp_u64(const void *a, const void *b)
{
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : x > y;
}

void oracle_mu_stats_agg(const int64_t *tau, const double *gamma, const double *eta,
                         const int64_t *variants, int V, int G, int S,
                         uint64_t seed, uint32_t sweep, int64_t v0,
                         int64_t *sum_mu, int64_t *esum)
{
    /* pattern code of every site: 2 bits per strain */
    uint64_t *code = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(V ? V : 1));
    uint64_t *uniq = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(V ? V : 1));
    for (int v = 0; v < V; v++) {
        uint64_t c = 0;
        for (int g = 0; g < G; g++) {
            int idx = 0;
            for (int b = 0; b < 4; b++) if (tau[((size_t)v * G + g) * 4 + b] == 1) { idx = b; break; }
            c |= (uint64_t)idx << (2 * g);
        }
        code[v] = c; uniq[v] = c;
    }
    qsort(uniq, (size_t)V, sizeof(uint64_t), cmp_u64);
    int P = 0;
    for (int v = 0; v < V; v++) if (v == 0 || uniq[v] != uniq[v - 1]) uniq[P++] = uniq[v];
    int64_t *N = (int64_t *)calloc((size_t)(P ? P : 1) * S * 4, sizeof(int64_t));
    for (int v = 0; v < V; v++) {
        uint64_t *hit = (uint64_t *)bsearch(&code[v], uniq, (size_t)P, sizeof(uint64_t), cmp_u64);
        int64_t *dst = N + (size_t)(hit - uniq) * S * 4;
        const int64_t *src = variants + (size_t)v * S * 4;
        for (int i = 0; i < S * 4; i++) dst[i] += src[i];
    }
    int64_t *classM = (G <= ORACLE_MUC_MAX_G) ? (int64_t *)calloc(((size_t)1 << G) * S, sizeof(int64_t)) : NULL;
#pragma omp parallel
    {
        int64_t *lmu = (int64_t *)calloc((size_t)S * G + 16, sizeof(int64_t));
        int64_t *le = lmu + (size_t)S * G;
        double sufg[65];
#pragma omp for schedule(dynamic, 8)
        for (int pi = 0; pi < P; pi++) {
            const uint64_t c = uniq[pi];
            /* classes of the pattern: class b = the strains whose base is b */
            uint32_t cmask[4] = {0u, 0u, 0u, 0u};
            for (int g = 0; g < G; g++) cmask[(c >> (2 * g)) & 3] |= 1u << g;
            const int lastk = cmask[3] ? 3 : cmask[2] ? 2 : cmask[1] ? 1 : 0;
            for (int s = 0; s < S; s++) {
                const int64_t *n = N + ((size_t)pi * S + s) * 4;
                if ((n[0] | n[1] | n[2] | n[3]) <= 0) continue;
                /* class abundances: ascending g, rounded adds from 0.0 */
                double Gm[4] = {0.0, 0.0, 0.0, 0.0};
                for (int g = 0; g < G; g++) { const int b = (int)((c >> (2 * g)) & 3); Gm[b] = Gm[b] + gamma[s * G + g]; }
                int64_t M[4] = {0, 0, 0, 0};
                bin_stream st;
                st.c0 = (uint32_t)c; st.c1 = (uint32_t)(c >> 32); st.c2 = sweep;
                st.seed = seed; st.shard = (uint32_t)v0;
                /* phase A: reads observed as a, split over the classes (the E statistics, HaploSNP_Sampler.py:301) */
                for (int a = 0; a < 4; a++) {
                    if (n[a] <= 0) continue;
                    st.c3 = ((uint32_t)ORACLE_STAGE_MUB << 28) | ((uint32_t)a << 26) | (uint32_t)s;
                    double W[4], suf[5];
                    suf[4] = 0.0;
                    for (int k = 3; k >= 0; k--) {
                        W[k] = cmask[k] ? eta[4 * k + a] * Gm[k] : 0.0;
                        suf[k] = cmask[k] ? W[k] + suf[k + 1] : suf[k + 1];
                    }
                    int64_t rem = n[a];
                    for (int k = 0; k < 4; k++) {
                        if (!cmask[k]) continue;
                        int64_t x;
                        if (k == lastk) x = rem;
                        else if (rem == 0) x = 0;
                        else { st.g = k; x = binomial_draw(rem, W[k] / suf[k], suf[k + 1] / suf[k], &st); }
                        rem -= x;
                        M[k] += x;
                        le[a * 4 + k] += x;
                    }
                }
                /* phase B: the reads of a class, split over its strains with weights gamma (:309, summed over a).
                 * The weights depend on the class only through its set of strains: for G <= ORACLE_MUC_MAX_G the totals of
                 * all patterns are merged per (set, sample) first and split once, below. */
                if (classM) {
                    for (int k = 0; k < 4; k++) {
                        if (!cmask[k] || M[k] <= 0) continue;
                        if ((cmask[k] & (cmask[k] - 1u)) == 0u) { int g = 0; while (!((cmask[k] >> g) & 1u)) g++; lmu[s * G + g] += M[k]; }
                        else {
#pragma omp atomic
                            classM[(size_t)cmask[k] * S + s] += M[k];
                        }
                    }
                    continue;
                }
                for (int k = 0; k < 4; k++) {
                    if (!cmask[k] || M[k] <= 0) continue;
                    st.c3 = ((uint32_t)ORACLE_STAGE_MUC << 28) | ((uint32_t)k << 26) | (uint32_t)s;
                    int gl = 0;
                    for (int g = 0; g < G; g++) if ((cmask[k] >> g) & 1u) gl = g;
                    double suf = 0.0;
                    for (int g = gl; g >= 0; g--) if ((cmask[k] >> g) & 1u) { suf = gamma[s * G + g] + suf; sufg[g] = suf; }
                    int64_t rem = M[k];
                    double sg = suf;
                    for (int g = 0; g <= gl; g++) {
                        if (!((cmask[k] >> g) & 1u)) continue;
                        int64_t x;
                        if (g == gl) x = rem;
                        else {
                            int gn = g + 1;
                            while (!((cmask[k] >> gn) & 1u)) gn++;
                            const double sn = sufg[gn];
                            st.g = g;
                            x = (rem == 0) ? 0 : binomial_draw(rem, gamma[s * G + g] / sg, sn / sg, &st);
                            sg = sn;
                        }
                        rem -= x;
                        lmu[s * G + g] += x;
                    }
                }
            }
        }
#pragma omp critical
        {
            for (size_t i = 0; i < (size_t)S * G; i++) sum_mu[i] += lmu[i];
            for (int i = 0; i < 16; i++) esum[i] += le[i];
        }
        free(lmu);
    }
    if (classM) {
        /* merged within-class split: set of strains `mask`, sample s, M reads, dealt by a balanced binary tree of binomial
         * splits over the ascending strain positions [lo, hi): mid = lo + (n+1)/2, X_left ~ Bin(M_node, L/(L+R), R/(L+R)),
         * L, R = sums of gamma[s,g] over the halves (ascending, rounded adds); stream ctr = (mask, 0, sweep, STAGE_MUC<<28 | s),
         * draw index = heap number of the node - 1 (root 1, children 2i, 2i+1) */
        const uint32_t nmask = 1u << G;
#pragma omp parallel
        {
            int64_t *lmu = (int64_t *)calloc((size_t)S * G, sizeof(int64_t));
#pragma omp for schedule(dynamic, 16)
            for (int64_t mk = 3; mk < (int64_t)nmask; mk++) {
                const uint32_t mask = (uint32_t)mk;
                if ((mask & (mask - 1u)) == 0u) continue;
                int pos2g[32], m = 0;
                for (int g = 0; g < G; g++) if ((mask >> g) & 1u) pos2g[m++] = g;
                int depth = 0;
                while ((1 << depth) < m) depth++;
                for (int s = 0; s < S; s++) {
                    const int64_t M = classM[(size_t)mask * S + s];
                    if (M <= 0) continue;
                    bin_stream st;
                    st.c0 = mask; st.c1 = 0u; st.c2 = sweep;
                    st.c3 = ((uint32_t)ORACLE_STAGE_MUC << 28) | (uint32_t)s;
                    st.seed = seed; st.shard = (uint32_t)v0;
                    int64_t nodeM[64];
                    nodeM[1] = M;
                    for (int level = 0; level < depth; level++)
                        for (int id = 1 << level; id < (2 << level); id++) {
                            int lo = 0, hi = m;
                            for (int b = level - 1; b >= 0; b--) {
                                const int mid = lo + (hi - lo + 1) / 2;
                                if ((id >> b) & 1) lo = mid; else hi = mid;
                            }
                            const int n = hi - lo;
                            if (n < 2) continue;
                            const int64_t Mn = nodeM[id];
                            const int mid = lo + (n + 1) / 2;
                            int64_t xl = 0;
                            if (Mn > 0) {
                                double L = 0.0, R = 0.0;
                                for (int i = lo; i < mid; i++) L = L + gamma[s * G + pos2g[i]];
                                for (int i = mid; i < hi; i++) R = R + gamma[s * G + pos2g[i]];
                                const double T = L + R;
                                st.g = id - 1;
                                xl = binomial_draw(Mn, L / T, R / T, &st);
                            }
                            const int64_t xr = Mn - xl;
                            if (mid - lo >= 2) nodeM[2 * id] = xl; else lmu[s * G + pos2g[lo]] += xl;
                            if (hi - mid >= 2) nodeM[2 * id + 1] = xr; else lmu[s * G + pos2g[mid]] += xr;
                        }
                }
            }
#pragma omp critical
            for (size_t i = 0; i < (size_t)S * G; i++) sum_mu[i] += lmu[i];
            free(lmu);
        }
        free(classM);
    }
    free(code); free(uniq); free(N);
}

/* ======================================================================== */
/* gamma / eta draws  (HaploSNP_Sampler.py:263-281)                          */
/* ======================================================================== */
static inline double u53(uint32_t hi, uint32_t lo)
{
    uint64_t m = ((uint64_t)(hi >> 5) << 26) | (uint64_t)(lo >> 6);
    return ((double)m + 0.5) * (1.0 / 9007199254740992.0);
}
static inline double u32(uint32_t w) { return ((double)w + 0.5) * (1.0 / 4294967296.0); }

/* Marsaglia-Tsang (2000) gamma variate; attempt t owns Philox block
 * ctr = (idx, t, sweep, stage<<28): (w0,w1) -> 53-bit radius uniform, w2 -> angle, w3 -> accept.
 * shape < 1 uses Gamma(shape+1) * U^(1/shape) with U from ctr = (idx, 0, sweep, boost_stage<<28). */
double oracle_gamma_variate(double shape, uint64_t seed, uint32_t sweep, uint32_t idx,
                            int stage, int boost_stage)
{
    double a1 = (shape < 1.0) ? shape + 1.0 : shape;
    double d = a1 - 1.0 / 3.0;
    double c = 1.0 / sqrt(9.0 * d);
    double y = d;
    for (uint32_t t = 0; t < (1u << 20); t++) {
        uint32_t o[4];
        philox(idx, t, sweep, (uint32_t)stage << 28, seed, o);
        double r1 = u53(o[0], o[1]);
        double r2 = u32(o[2]);
        double z = sqrt(-2.0 * log(r1)) * cos(6.283185307179586476925 * r2);
        double vv = 1.0 + c * z;
        if (vv <= 0.0) continue;
        vv = vv * vv * vv;
        double r3 = u32(o[3]);
        if (log(r3) < 0.5 * z * z + d - d * vv + d * log(vv)) { y = d * vv; break; }
    }
    if (shape < 1.0) {
        uint32_t o[4];
        philox(idx, 0u, sweep, (uint32_t)boost_stage << 28, seed, o);
        y *= exp(log(u53(o[0], o[1])) / shape);
    }
    return y;
}

/* sampleGamma, HaploSNP_Sampler.py:263-273: gamma[s,:] ~ Dir(alpha + sum_mu[s,:]);
 * entries < epsilon set to epsilon (:271), rows renormalised (:272-273). */
void oracle_draw_gamma(const int64_t *sum_mu, int S, int G, double alpha, double epsilon,
                       uint64_t seed, uint32_t sweep, double *gamma)
{
    for (int s = 0; s < S; s++) {
        double tot = 0.0;
        for (int g = 0; g < G; g++) {
            double y = oracle_gamma_variate(alpha + (double)sum_mu[s * G + g], seed, sweep,
                                            (uint32_t)(s * G + g), ORACLE_STAGE_GAMMA, ORACLE_STAGE_GAMMA_BOOST);
            gamma[s * G + g] = y;
            tot += y;
        }
        double rs = 0.0;
        for (int g = 0; g < G; g++) {
            double x = (tot > 0.0) ? gamma[s * G + g] / tot : 1.0 / G;
            if (x < epsilon) x = epsilon;
            gamma[s * G + g] = x;
            rs += x;
        }
        for (int g = 0; g < G; g++) gamma[s * G + g] = gamma[s * G + g] / rs;
    }
}

/* sampleEta, HaploSNP_Sampler.py:275-281: eta[t,:] ~ Dir(delta + Esum[:,t]) with
 * Esum[a_obs, b_true]; row t = true base, columns observed.  No clipping. */
void oracle_draw_eta(const int64_t *esum, double delta, uint64_t seed, uint32_t sweep, double *eta)
{
    for (int t = 0; t < 4; t++) {
        double tot = 0.0;
        for (int o = 0; o < 4; o++) {
            double y = oracle_gamma_variate(delta + (double)esum[o * 4 + t], seed, sweep,
                                            (uint32_t)(t * 4 + o), ORACLE_STAGE_ETA, ORACLE_STAGE_ETA_BOOST);
            eta[t * 4 + o] = y;
            tot += y;
        }
        for (int o = 0; o < 4; o++) eta[t * 4 + o] = eta[t * 4 + o] / tot;
    }
}

/* ======================================================================== */
/* log-likelihood / log-posterior  (HaploSNP_Sampler.py:431-461)             */
/* ======================================================================== */
/* logLikelihood: p[v,s,a] = sum_{g,b} tau[v,g,b] gamma[s,g] eta[b,a] (:435);
 * ll += lgamma(N+1) - sum_b lgamma(n_b+1) + sum_b n_b log p_b (Desman_Utils.py:28-33).
 * tau may be non-one-hot here (DIC calls it with tauMean, :494), so the general
 * contraction is kept. */
double oracle_loglik(const int64_t *tau, const double *gamma, const double *eta,
                     const int64_t *variants, int V, int G, int S)
{
    double ll = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : ll)
    for (int v = 0; v < V; v++) {
        double te[32 * 4];
        double *tev = te;
        double *heap = NULL;
        if (G > 32) { heap = (double *)malloc(sizeof(double) * 4 * (size_t)G); tev = heap; }
        for (int g = 0; g < G; g++)
            for (int a = 0; a < 4; a++) {
                double x = 0.0;
                for (int b = 0; b < 4; b++) x += (double)tau[((size_t)v * G + g) * 4 + b] * eta[b * 4 + a];
                tev[g * 4 + a] = x;
            }
        double lv = 0.0;
        for (int s = 0; s < S; s++) {
            const int64_t *n = variants + ((size_t)v * S + s) * 4;
            int64_t N = n[0] + n[1] + n[2] + n[3];
            double r = lgamma((double)N + 1.0);
            double sub = 0.0, dot = 0.0;
            for (int a = 0; a < 4; a++) {
                double p = 0.0;
                for (int g = 0; g < G; g++) p += gamma[s * G + g] * tev[g * 4 + a];
                sub += lgamma((double)n[a] + 1.0);
                dot += (double)n[a] * log(p);
            }
            lv += r - sub + dot;
        }
        ll += lv;
        free(heap);
    }
    return ll;
}

/* log_dirichlet_pdf, Desman_Utils.py:35-44 */
static double log_dirichlet(const double *x, int n, double conc)
{
    double ret = lgamma(conc * n);
    for (int i = 0; i < n; i++) {
        ret += (conc - 1.0) * log(x[i]);
        ret -= lgamma(conc);
    }
    return ret;
}

/* logPosterior minus logLL, HaploSNP_Sampler.py:448-459 */
double oracle_logprior(const double *gamma, const double *eta, int V, int G, int S,
                       double alpha, double delta)
{
    double lg = 0.0, le = 0.0;
    for (int s = 0; s < S; s++) lg += log_dirichlet(gamma + (size_t)s * G, G, alpha);
    for (int a = 0; a < 4; a++) le += log_dirichlet(eta + a * 4, 4, delta);
    double lt = (double)V * (double)G * log(1.0 / 4.0);
    return lg + le + lt;
}

/* ======================================================================== */
/* chain drivers                                                            */
/* ======================================================================== */
static void add_tau_sum(int64_t *tau_sum, const int64_t *tau, size_t n)
{
    if (!tau_sum) return;
    for (size_t i = 0; i < n; i++) tau_sum[i] += tau[i];
}

/* update(), HaploSNP_Sampler.py:334-365.  Sweep order mu/E -> gamma -> tau -> eta -> ll/lp
 * (:341-350); eta is drawn from the E sampled with the OLD tau; star initialised from the
 * pre-sweep state (:336-338) and replaced on strict lp > lp_star (:352). */
void oracle_update(const oracle_chain_cfg *cfg, int64_t *tau, double *gamma, double *eta,
                   const int64_t *variants,
                   double *gamma_store, double *eta_store, double *ll_store, double *lp_store,
                   int64_t *nchange_store, int64_t *tau_sum,
                   int64_t *tau_star, double *gamma_star, double *eta_star,
                   double *lp_star_out, int *iter_star_out,
                   int64_t *sum_mu_last, int64_t *esum_last)
{
    int V = cfg->V, G = cfg->G, S = cfg->S;
    size_t nt = (size_t)V * G * 4, ng = (size_t)S * G;
    int64_t *sum_mu = (int64_t *)malloc(sizeof(int64_t) * ng);
    int64_t esum[16];
    double ll = oracle_loglik(tau, gamma, eta, variants, V, G, S);
    double lp = ll + oracle_logprior(gamma, eta, V, G, S, cfg->alpha, cfg->delta);
    double lp_star = lp;
    int iter_star = 0;
    if (tau_star) memcpy(tau_star, tau, sizeof(int64_t) * nt);
    if (gamma_star) memcpy(gamma_star, gamma, sizeof(double) * ng);
    if (eta_star) memcpy(eta_star, eta, sizeof(double) * 16);
    for (int it = 0; it < cfg->n_iter; it++) {
        uint32_t sweep = cfg->sweep0 + (uint32_t)it;
        memset(sum_mu, 0, sizeof(int64_t) * ng);
        memset(esum, 0, sizeof(esum));
        if (cfg->mu_mode == 1) oracle_mu_stats_agg(tau, gamma, eta, variants, V, G, S, cfg->seed, sweep, 0, sum_mu, esum);
        else oracle_mu_stats(tau, gamma, eta, variants, V, G, S, cfg->seed, sweep, 0, sum_mu, esum);
        oracle_draw_gamma(sum_mu, S, G, cfg->alpha, cfg->epsilon, cfg->seed, sweep, gamma);
        int nchange = oracle_sample_tau_philox(tau, gamma, eta, variants, V, G, S, cfg->seed, sweep, 0);
        oracle_draw_eta(esum, cfg->delta, cfg->seed, sweep, eta);
        ll = oracle_loglik(tau, gamma, eta, variants, V, G, S);
        lp = ll + oracle_logprior(gamma, eta, V, G, S, cfg->alpha, cfg->delta);
        if (ll_store) ll_store[it] = ll;
        if (lp_store) lp_store[it] = lp;
        if (nchange_store) nchange_store[it] = nchange;
        if (lp > lp_star) {
            lp_star = lp; iter_star = it;
            if (tau_star) memcpy(tau_star, tau, sizeof(int64_t) * nt);
            if (gamma_star) memcpy(gamma_star, gamma, sizeof(double) * ng);
            if (eta_star) memcpy(eta_star, eta, sizeof(double) * 16);
        }
        add_tau_sum(tau_sum, tau, nt);
        if (gamma_store) memcpy(gamma_store + (size_t)it * ng, gamma, sizeof(double) * ng);
        if (eta_store) memcpy(eta_store + (size_t)it * 16, eta, sizeof(double) * 16);
    }
    if (sum_mu_last) memcpy(sum_mu_last, sum_mu, sizeof(int64_t) * ng);
    if (esum_last) memcpy(esum_last, esum, sizeof(esum));
    if (lp_star_out) *lp_star_out = lp_star;
    if (iter_star_out) *iter_star_out = iter_star;
    free(sum_mu);
}

/* updateTau(), HaploSNP_Sampler.py:383-407: tau-only replay against stored gamma/eta.
 * lp_star starts from logPosterior(gamma_store[0], tau, eta_store[0]) (:386-388). */
void oracle_update_tau(const oracle_chain_cfg *cfg, int use_mt, oracle_mt19937 *rng,
                       int64_t *tau, const double *gamma_store, const double *eta_store,
                       const int64_t *variants, double *ll_store, double *lp_store,
                       int64_t *nchange_store, int64_t *tau_sum, int64_t *tau_star, double *lp_star_out)
{
    int V = cfg->V, G = cfg->G, S = cfg->S;
    size_t nt = (size_t)V * G * 4, ng = (size_t)S * G;
    double lp = oracle_loglik(tau, gamma_store, eta_store, variants, V, G, S) +
                oracle_logprior(gamma_store, eta_store, V, G, S, cfg->alpha, cfg->delta);
    double lp_star = lp;
    if (tau_star) memcpy(tau_star, tau, sizeof(int64_t) * nt);
    for (int it = 0; it < cfg->n_iter; it++) {
        const double *gm = gamma_store + (size_t)it * ng;
        const double *et = eta_store + (size_t)it * 16;
        int nchange = use_mt ? oracle_sample_tau_mt(tau, gm, et, variants, V, G, S, rng)
                             : oracle_sample_tau_philox(tau, gm, et, variants, V, G, S, cfg->seed,
                                                        cfg->sweep0 + (uint32_t)it, 0);
        double ll = oracle_loglik(tau, gm, et, variants, V, G, S);
        lp = ll + oracle_logprior(gm, et, V, G, S, cfg->alpha, cfg->delta);
        if (lp > lp_star) { lp_star = lp; if (tau_star) memcpy(tau_star, tau, sizeof(int64_t) * nt); }
        add_tau_sum(tau_sum, tau, nt);
        if (ll_store) ll_store[it] = ll;
        if (lp_store) lp_store[it] = lp;
        if (nchange_store) nchange_store[it] = nchange;
    }
    if (lp_star_out) *lp_star_out = lp_star;
}

/* ======================================================================== */
/* NMFT  (Init_NMFT.py)                                                      */
/* ======================================================================== */
#define NMFT_EPS DBL_EPSILON /* np.finfo(float64).eps, Desman_Utils.py:16-17, Init_NMFT.py:90-95 */

/* Init_NMFT.__init__, Init_NMFT.py:49-60: (n+1)/sum_b(n_b+1), rows v + a*V */
void oracle_nmft_freq(const int64_t *snps, int V, int S, double *freq)
{
    for (int v = 0; v < V; v++)
        for (int s = 0; s < S; s++) {
            const int64_t *n = snps + ((size_t)v * S + s) * 4;
            double x[4], tot = 0.0;
            for (int a = 0; a < 4; a++) { x[a] = (double)n[a] + 1.0; tot += x[a]; }
            for (int a = 0; a < 4; a++) freq[((size_t)v + (size_t)a * V) * S + s] = x[a] / tot;
        }
}

static inline double nz(double x) { return x == 0.0 ? NMFT_EPS : x; } /* elop zero rule */

/* div_objective, Init_NMFT.py:152-156 */
double oracle_nmft_objective(const double *freq, const double *tau, const double *gamma,
                             int V, int G, int S)
{
    size_t N = (size_t)4 * V;
    double tot = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : tot)
    for (size_t n = 0; n < N; n++) {
        double row = 0.0;
        for (int s = 0; s < S; s++) {
            double pa = 0.0;
            for (int g = 0; g < G; g++) pa += tau[n * G + g] * gamma[(size_t)g * S + s];
            if (pa < NMFT_EPS) pa = NMFT_EPS;                   /* _adjustment_input :93-97 */
            double x = freq[n * S + s];
            row += x * log(nz(x) / nz(pa)) - x + pa;
        }
        tot += row;
    }
    return tot;
}

/* div_update (:158-181), div_update_gamma (:183-190), div_update_tau (:192-205) */
void oracle_nmft_update(const double *freq, double *tau, double *gamma, int V, int G, int S,
                        int update_gamma, int update_tau)
{
    size_t N = (size_t)4 * V;
    if (update_gamma) {
        if (G > 1) {
            double *h1 = (double *)calloc((size_t)G, sizeof(double));
            double *num = (double *)calloc((size_t)G * S, sizeof(double));
            for (size_t n = 0; n < N; n++)
                for (int g = 0; g < G; g++) h1[g] += tau[n * G + g];
#pragma omp parallel
            {
                double *lnum = (double *)calloc((size_t)G * S, sizeof(double));
#pragma omp for schedule(static)
                for (size_t n = 0; n < N; n++)
                    for (int s = 0; s < S; s++) {
                        double pa = 0.0;
                        for (int g = 0; g < G; g++) pa += tau[n * G + g] * gamma[(size_t)g * S + s];
                        double r = nz(freq[n * S + s]) / nz(pa);
                        for (int g = 0; g < G; g++) lnum[(size_t)g * S + s] += tau[n * G + g] * r;
                    }
#pragma omp critical
                for (size_t i = 0; i < (size_t)G * S; i++) num[i] += lnum[i];
                free(lnum);
            }
            for (int g = 0; g < G; g++)
                for (int s = 0; s < S; s++)
                    gamma[(size_t)g * S + s] *= nz(num[(size_t)g * S + s]) / nz(h1[g]);
            for (int s = 0; s < S; s++) {
                double cs = 0.0;
                for (int g = 0; g < G; g++) cs += gamma[(size_t)g * S + s];
                for (int g = 0; g < G; g++) gamma[(size_t)g * S + s] /= cs;
            }
            free(h1); free(num);
        } else {
            for (int s = 0; s < S; s++) gamma[s] = 1.0;                     /* :167-168 */
        }
    }
    if (update_tau) {
        double *t1 = (double *)calloc((size_t)G, sizeof(double));
        for (int g = 0; g < G; g++)
            for (int s = 0; s < S; s++) t1[g] += gamma[(size_t)g * S + s];
#pragma omp parallel for schedule(static)
        for (int v = 0; v < V; v++) {
            double numt[4][64];
            for (int a = 0; a < 4; a++) {
                size_t n = (size_t)v + (size_t)a * V;
                for (int g = 0; g < G; g++) numt[a][g] = 0.0;
                for (int s = 0; s < S; s++) {
                    double pa = 0.0;
                    for (int g = 0; g < G; g++) pa += tau[n * G + g] * gamma[(size_t)g * S + s];
                    double r = nz(freq[n * S + s]) / nz(pa);
                    for (int g = 0; g < G; g++) numt[a][g] += r * gamma[(size_t)g * S + s];
                }
            }
            for (int g = 0; g < G; g++) {
                double sumvg = 0.0;
                for (int a = 0; a < 4; a++) {
                    size_t n = (size_t)v + (size_t)a * V;
                    tau[n * G + g] *= nz(numt[a][g]) / nz(t1[g]);
                    sumvg += tau[n * G + g];
                }
                for (int a = 0; a < 4; a++) {                                /* :174-181 */
                    size_t n = (size_t)v + (size_t)a * V;
                    tau[n * G + g] = tau[n * G + g] / sumvg;
                }
            }
        }
        free(t1);
    }
}

int oracle_nmft_factorize(const double *freq, double *tau, double *gamma, int V, int G, int S,
                          int max_iter, double min_change, int fix_gamma,
                          double *div_trace, double *div_final)
{
    size_t N = (size_t)4 * V;
    if (!fix_gamma) {                                                        /* _adjustment :88-91 */
        for (size_t i = 0; i < N * G; i++) if (tau[i] < NMFT_EPS) tau[i] = NMFT_EPS;
        for (size_t i = 0; i < (size_t)G * S; i++) if (gamma[i] < NMFT_EPS) gamma[i] = NMFT_EPS;
    }
    double divl = 0.0, div = oracle_nmft_objective(freq, tau, gamma, V, G, S);
    int iter = 0;
    while (iter < max_iter && fabs(divl - div) > min_change) {               /* :106 / :140 */
        oracle_nmft_update(freq, tau, gamma, V, G, S, !fix_gamma, 1);
        if (!fix_gamma) {
            for (size_t i = 0; i < N * G; i++) if (tau[i] < NMFT_EPS) tau[i] = NMFT_EPS;
            for (size_t i = 0; i < (size_t)G * S; i++) if (gamma[i] < NMFT_EPS) gamma[i] = NMFT_EPS;
        }
        divl = div;
        div = oracle_nmft_objective(freq, tau, gamma, V, G, S);
        if (div_trace) div_trace[iter] = div;
        iter++;
    }
    if (div_final) *div_final = div;
    return iter;
}

/* get_tau, Init_NMFT.py:230-245: strict '>' from maxt = 0.0, ties -> lowest base */
void oracle_nmft_get_tau(const double *tau, int V, int G, int64_t *tau_onehot)
{
    memset(tau_onehot, 0, sizeof(int64_t) * (size_t)V * G * 4);
    for (int v = 0; v < V; v++)
        for (int g = 0; g < G; g++) {
            double maxt = 0.0; int maxa = 0;
            for (int a = 0; a < 4; a++) {
                double x = tau[((size_t)v + (size_t)a * V) * G + g];
                if (x > maxt) { maxt = x; maxa = a; }
            }
            tau_onehot[((size_t)v * G + g) * 4 + maxa] = 1;
        }
}

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void oracle_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
