/* include/desman_b200.h -- C-ABI of libdesman_b200.so
 *
 * B200 (sm_100a) engine for DESMAN's haplotype-inference Gibbs sweep.  Plain C
 * types only: host pointers and sizes in, host pointers out.  No torch types.
 *
 * Part 1 is the reference's own native ABI, symbol for symbol, so that the
 * reference's Cython module (sampletau/sampletau.pyx:13-19) -- or any other FFI
 * bound to sampletau/c_sample_tau.c -- links against this library unchanged.
 * Part 2 is the handle-based API behind desman_b200.HaploSNP_Sampler /
 * desman_b200.Init_NMFT (device-resident chains; the per-call host<->device
 * traffic of Part 1 disappears).
 *
 * Error convention (Part 2): every function returns 0 on success, a negative
 * DESMAN_E* code otherwise; desman_last_error() returns a thread-local message.
 * Nothing in this library calls exit() (the reference does on malloc failure,
 * c_sample_tau.c:200-203).
 */
#ifndef DESMAN_B200_H
#define DESMAN_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ Part 1
 * Reference ABI.  `long` is 64-bit (LP64), as in the reference build.
 *
 * c_initRNG   replaces sampletau/c_sample_tau.c:26-34  (allocates the process-global stream,
 *             binds the default CUDA device)
 * c_setRNG    replaces c_sample_tau.c:36-40  (GSL gsl_rng_mt19937 seeding; seed 0 -> 4357)
 * c_freeRNG   replaces c_sample_tau.c:42-45
 * c_sample_tau replaces c_sample_tau.c:95-204: anTau int64 one-hot [nV,nG,4] mutated in
 *             place; adPi = gamma [nS,nG]; adEta [4,4] row = true base; anVariants int64
 *             [nV,nS,4].  Consumes nV*nG words of the MT19937 stream in (v,g) order and
 *             returns the number of flipped (v,g) entries.  On failure (no device, counts
 *             out of range, tau not one-hot) returns -1 with desman_last_error() set
 *             instead of exiting the process. */
void c_initRNG(void);
void c_setRNG(unsigned long int seed);
void c_freeRNG(void);
int  c_sample_tau(long *anTau, double *adPi, double *adEta, long *anVariants, int nV, int nG, int nS);

/* ------------------------------------------------------------------ Part 2 */
typedef struct desman_ctx desman_ctx;

enum { DESMAN_OK = 0, DESMAN_EINVAL = -1, DESMAN_ECUDA = -2, DESMAN_ENOMEM = -3, DESMAN_ESTATE = -4,
       DESMAN_ECOMM = -5 };
enum { DESMAN_RNG_MT19937 = 0,  /* tau words from the GSL-compatible MT19937 stream (reference-exact) */
       DESMAN_RNG_PHILOX = 1 }; /* Philox4x32-10 counter contract, required for update() */
#define DESMAN_MAX_G 32
#define DESMAN_MAX_COUNT 16777216 /* 2^24: per-cell count limit (float cast exact, c_sample_tau.c:164) */

const char *desman_last_error(void);
int desman_device_count(int *n);
/* version string + the sm arch the kernels were built for */
const char *desman_build_info(void);

int desman_ctx_create(desman_ctx **out, int device, uint64_t seed, int rng_mode);
int desman_ctx_destroy(desman_ctx *ctx);

/* Counts: int64 [V,S,4] C-order (HaploSNP_Sampler.__init__, HaploSNP_Sampler.py:48-52).  Repacked on
 * device to one 128-bit int32x4 word per (v,s) cell.  v0/V_total describe a shard: this context
 * owns global sites [v0, v0+V) of V_total (single GPU: v0 = 0, V_total = V). */
int desman_set_counts(desman_ctx *ctx, const int64_t *variants, int64_t V, int S, int64_t v0, int64_t V_total);
/* hyper-parameters alpha, delta, epsilon (HaploSNP_Sampler.py:31) */
int desman_set_hyper(desman_ctx *ctx, double alpha, double delta, double epsilon);
/* seed / position of the random streams: philox sweep counter, MT19937 words already consumed */
int desman_set_rng(desman_ctx *ctx, uint64_t seed, uint32_t sweep, uint64_t mt_words_consumed);
int desman_get_rng(desman_ctx *ctx, uint32_t *sweep, uint64_t *mt_words_consumed);

/* State sync with the externally assignable class attributes tau [V,G,4] int64 one-hot, gamma [S,G],
 * eta [4,4] (bin/desman:140-146).  Any pointer may be NULL (left unchanged / not returned). */
int desman_set_state(desman_ctx *ctx, const int64_t *tau, const double *gamma, const double *eta, int G);
int desman_get_state(desman_ctx *ctx, int64_t *tau, double *gamma, double *eta);
/* compact form of tau: uint8 base index [V,G] */
int desman_set_tau_index(desman_ctx *ctx, const uint8_t *tau_idx, int G);
int desman_get_tau_index(desman_ctx *ctx, uint8_t *tau_idx);

/* Single steps (used by the parity tests and by sampletau.sample_tau on a resident context) */
int desman_sample_tau(desman_ctx *ctx, int64_t *nchange);                       /* c_sample_tau.c:130-188 */
int desman_mu_stats(desman_ctx *ctx, int64_t *sum_mu /*S*G*/, int64_t *esum /*16*/); /* HaploSNP_Sampler.py:284-309 */
int desman_draw_gamma_eta(desman_ctx *ctx, const int64_t *sum_mu, const int64_t *esum,
                          double *gamma /*S*G*/, double *eta /*16*/);          /* :263-281 (does not change state) */
int desman_loglik(desman_ctx *ctx, double *ll, double *lp);                      /* :431-461 */
/* logLikelihood under a real-valued tau [V,G,4] (the tauMean of DIC, :479-496) on the counts of the context */
int desman_loglik_general(desman_ctx *ctx, const double *tau /*V*G*4*/, const double *gamma /*S*G*/, const double *eta /*16*/,
                          int G, double *ll);
/* Joint-state enumeration (assignTau :233-261, logTauProb :498-524): for N sites and all T = 4^G states t (strain g = digit
 * G-1-g of t in base 4, the reference's tauStates order) stateLogProb[n][t] = sum_{s,b} n_sb log(sum_g gamma[s,g] eta[t_g,b]).
 * variants: int64 [N,S,4] host counts, or NULL for the counts of the context (then N = V, S = the context's).  Outputs, each
 * may be NULL: logprob [N,T] (all of it: mind the size), maxlp [N], lse [N] = log sum_t exp, argmax [N] (first maximum);
 * lp_at_index [N] = stateLogProb[n][index[n]] when index is given.  4^G * 4S doubles must fit in 4 GiB (G <= 10 at S = 64). */
int desman_state_logprob(desman_ctx *ctx, const int64_t *variants, int64_t N, int S, const double *gamma /*S*G*/,
                         const double *eta /*16*/, int G, const int64_t *index, double *logprob, double *lp_at_index,
                         double *maxlp, double *lse, int64_t *argmax);

/* update(): n_iter full Gibbs sweeps on device (HaploSNP_Sampler.py:334-365).  Output arrays may be
 * NULL; gamma_store [n_iter,S,G], eta_store [n_iter,4,4], ll_store/lp_store/nchange_store [n_iter]. */
int desman_update(desman_ctx *ctx, int n_iter, double *gamma_store, double *eta_store,
                  double *ll_store, double *lp_store, int64_t *nchange_store);
/* updateTau(): tau-only replay against stored gamma/eta (HaploSNP_Sampler.py:383-407) */
int desman_update_tau(desman_ctx *ctx, int n_iter, const double *gamma_store, const double *eta_store,
                      double *ll_store, double *lp_store, int64_t *nchange_store);
/* MAP state tracked by the last update()/update_tau() (storeStarState, :326-332) */
int desman_get_star(desman_ctx *ctx, int64_t *tau_star, double *gamma_star, double *eta_star,
                    double *lp_star, int *iter_star);
int desman_get_star_index(desman_ctx *ctx, uint8_t *tau_star_idx /*V*G*/);   /* tau_star as base indices */
/* E_store[i].sum(axis=(0,1)) for every sweep i of the last desman_update: Esum[a_obs][b_true] (:557) */
int desman_get_esum_store(desman_ctx *ctx, int64_t *esum_store /*n_iter*16*/);
/* sum over the sweeps of the last update()/update_tau() of one-hot tau: tau_store.sum(axis=0)
 * (tauMean :479-483, probabilisticTau :834-840) */
int desman_get_tau_sum(desman_ctx *ctx, int64_t *tau_sum /*V*G*4*/);
int desman_get_tau_sum_u32(desman_ctx *ctx, uint32_t *tau_sum /*V*G*4*/);   /* same counters, as kept on the device */

/* NMFT initialiser (Init_NMFT.py).  snps int64 [V,S,4]; tau [4V,G] (rows v + a*V) and gamma [G,S] hold
 * the random initial factors on entry and the result on exit.  fix_gamma = 0: factorize (:98-115);
 * fix_gamma = 1: factorize_tau (:134-149); fix_gamma = 2: factorize_gamma (:117-132: tau stays, no eps clamps).  max_iter = 0
 * evaluates div_objective (:152-156) of the factors as they are (with fix_gamma != 0); max_iter = 1 with min_change < 0 is one
 * div_update / div_update_tau / div_update_gamma step (:158-205).  div_trace (may be NULL) gets div after each iteration. */
int desman_nmft_factorize(desman_ctx *ctx, const int64_t *snps, int64_t V, int S, int G,
                          double *tau, double *gamma, int max_iter, double min_change, int fix_gamma,
                          int *n_iter_done, double *div_final, double *div_trace);

/* device time (CUDA events) of the iteration loop of the last desman_nmft_factorize of this process, and its iterations */
int desman_nmft_last_timing(double *ms, int *iters);

/* Engine options.  "mu_mode" = 1: mu/E statistics from pattern-aggregated conditional binomials; 0: one categorical
 * draw per read (both exact, different counter contracts); 2 (default): 1 iff 12*2^G <= V/2.  "fixed_tau" = 1 makes desman_update skip the tau draw (update_fixed_tau, HaploSNP_Sampler.py:409-428).
 * "tau_exact" = 1 forces the FP64 reference-order arithmetic for every tau draw (validation;
 * default 0 = filtered-exact FP32 fast path with FP64 fallback, same draws).  desman_get_tier_counts returns how
 * many draws were decided by the FP32 gap test / the FP64 CDF brackets / the FP64 reference-order recompute. */
int desman_set_option(desman_ctx *ctx, const char *name, int64_t value);
int desman_get_tier_counts(desman_ctx *ctx, int64_t out[3], int reset);
/* "tau_group" = 1: the tau update first screens the sites grouped by haplotype pattern (one table of log terms per
 * pattern, a dense FP32 contraction per site) and walks only the undecided sites with the per-site kernel; 0: per-site
 * kernel for every site; 2 (default): 1 iff 12*2^G <= V/2.  Same draws either way.  desman_get_group_stats returns
 * {groups valid, chain calm, work items, single-site patterns, sites left to the per-site kernel in the last sweep,
 *  orphans since the last regroup, pattern slots in use, grouping configured}. */
int desman_get_group_stats(desman_ctx *ctx, int64_t out[8]);
/* "tau_group_tc" = 1 (default): where every count is < 2048 the screening pass runs on the Blackwell tensor path (TMA bulk
 * copies into shared memory, tcgen05.mma with the sums in tensor memory; tau_group_tc_kernel.cuh); 0: mma.sync / FFMA forms.
 * desman_debug_screen (validation only) regroups the current state, runs that pass once and returns per site the 3G sums
 * D[v][3g+j] = L(candidate j of strain g) - L(current base) in log2 units (NaN: site in no group) and the mask of strains
 * the gap test left undecided (0xffffffff: none, the site is not on the work list). */
int desman_debug_screen(desman_ctx *ctx, float *D /*V*3G*/, uint32_t *mask /*V*/);

/* sampleTauFixTau (HaploSNP_Sampler.py:196-222) on the current device state: strains [H, G) of every site are redrawn in
 * order under the Philox contract of the tau draws; logp [V][4] (may be NULL) = normalised log-probabilities
 * (normaliseLogProb, :186-194) of strain H's four bases before its draw.  Option "advance_sweep" = n moves the Philox sweep
 * counter on by n (loops over desman_mu_stats / desman_draw_gamma_eta, which leave it where it is). */
int desman_sample_tau_fix(desman_ctx *ctx, int H, double *logp /*V*4*/, int64_t *nchange);

/* Batched form of the reference ABI for callers that make many small sample_tau calls per iteration (Eta_Sampler.sampleTauC,
 * Eta_Sampler.py:355-369,:430-446: one call per gene, each with its own masked gamma).  The counts of all problems are
 * uploaded once and stay resident; desman_batch_sample_tau is ONE launch over all of them and, draw for draw (one
 * process-global MT19937 stream, problem after problem), equals the sequence of c_sample_tau calls it replaces.  Needs c_initRNG. */
typedef struct desman_batch desman_batch;
int desman_batch_create(desman_batch **out, int nprob, const int64_t *const *variants /*[nprob] -> [V_k][S][4]*/, const int *nV, int nS);
int desman_batch_sample_tau(desman_batch *b, int64_t *const *tau /*[nprob] -> one-hot [V_k][G][4], in place*/,
                            const double *const *pi /*[nprob] -> [S][G]*/, const double *eta /*[4][4]*/, int nG, int *nchange /*[nprob]*/);
int desman_batch_destroy(desman_batch *b);

/* Multi-GPU: one context per process/GPU, sites sharded by desman_set_counts(v0, V_total).
 * desman_comm_unique_id fills a 128-byte NCCL id on rank 0; every rank calls desman_comm_init. */
int desman_comm_unique_id(char id[128]);
int desman_comm_init(desman_ctx *ctx, const char id[128], int rank, int nranks);
/* data plane of the per-sweep exchange: 0 none (one rank), 1 NCCL all-reduce, 2 one-shot peer-memory mailboxes over NVLink */
int desman_comm_kind(desman_ctx *ctx);

/* Measurement hooks (bench.py): device-timed sweeps with CUDA events on the engine's stream. */
enum { DESMAN_K_TAU = 0, DESMAN_K_MU = 1, DESMAN_K_DRAW = 2, DESMAN_K_FINAL = 3, DESMAN_K_MT = 4,
       DESMAN_K_NMFT = 5, DESMAN_K_OTHER = 6, DESMAN_K_TAU_GROUP = 7, DESMAN_K_MAINT = 8, DESMAN_K_TAU_UPDATE = 9,
       DESMAN_K_COUNT = 10 };
/* per_kernel_events: 0 none (sweeps only); 1 an event pair around every kernel launch (an event between two launches also
 * undoes their programmatic overlap, so these times add up to more than the sweep); 2 ONE pair around the whole tau update
 * (screening pass + the kernels that walk its work list, launched as the dependent chain they are in production): reported
 * as DESMAN_K_TAU_UPDATE and nothing else. */
int desman_set_profiling(desman_ctx *ctx, int per_kernel_events, int flush_l2_between_sweeps);
/* elapsed_ms: sum over sweeps of the event-timed sweep durations of the last update()/update_tau() */
int desman_get_timing(desman_ctx *ctx, double *elapsed_ms, double kernel_ms[DESMAN_K_COUNT],
                      int64_t kernel_launches[DESMAN_K_COUNT]);
int desman_synchronize(desman_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif
