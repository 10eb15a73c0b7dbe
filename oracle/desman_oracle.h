/* oracle/desman_oracle.h -- TEST INFRASTRUCTURE ONLY, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C99) of the DESMAN haplotype-inference hot path.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library; the product (desman_b200/) never does.
 *
 * Every function cites the reference file:line (paths relative to the
 * reference checkout) whose arithmetic it restates.  Parity pins: see
 * oracle/README.md and tests/golden/ (vectors generated from the UNMODIFIED
 * reference run in the build container).
 */
#ifndef DESMAN_ORACLE_H
#define DESMAN_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- RNG primitives ------------------------------------------------------ */
/* GSL-compatible MT19937 stream (c_sample_tau.c:24-45,174 via gsl_rng_mt19937). */
typedef struct { uint32_t mt[624]; int mti; } oracle_mt19937;
void     oracle_mt_seed(oracle_mt19937 *r, unsigned long seed);
uint32_t oracle_mt_next(oracle_mt19937 *r);
void     oracle_mt_fill(oracle_mt19937 *r, uint32_t *out, int64_t n);

/* Philox4x32-10 (Salmon et al. 2011), the counter-based contract of the device chain. */
void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* counter "stage" tags (c3 = stage<<28 | ...), shared by contract with the CUDA path */
enum { ORACLE_STAGE_TAU = 1, ORACLE_STAGE_MU = 2, ORACLE_STAGE_GAMMA = 3, ORACLE_STAGE_ETA = 4,
       ORACLE_STAGE_GAMMA_BOOST = 5, ORACLE_STAGE_ETA_BOOST = 6 };

/* ---- tau update ---------------------------------------------------------- */
/* c_sample_tau.c:95-204 with the V*G uniform words supplied by the caller
 * (u = word / 2^32, one word per (v,g), v outer, g inner).  tau is int64 one-hot
 * [V,G,4] mutated in place.  Returns the number of flips. */
int oracle_sample_tau_words(int64_t *tau, const double *pi, const double *eta,
                            const int64_t *variants, int V, int G, int S,
                            const uint32_t *words);
/* same, drawing the words from a GSL-compatible MT19937 stream (reference behaviour) */
int oracle_sample_tau_mt(int64_t *tau, const double *pi, const double *eta,
                         const int64_t *variants, int V, int G, int S, oracle_mt19937 *rng);
/* same, words = Philox(ctr=(v0+v, g, sweep, STAGE_TAU<<28), key=seed).out[0] */
int oracle_sample_tau_philox(int64_t *tau, const double *pi, const double *eta,
                             const int64_t *variants, int V, int G, int S,
                             uint64_t seed, uint32_t sweep, int64_t v0);
int oracle_sample_tau_fix_philox(int64_t *tau, const double *pi, const double *eta,
                                 const int64_t *variants, int V, int G, int S,
                                 uint64_t seed, uint32_t sweep, int64_t v0, int H, double *logp_out);
/* the four candidate log-likelihoods and normalised probabilities of one (v,g) step
 * (c_sample_tau.c:136-172), for inspection in tests */
void oracle_tau_step_probs(const int64_t *tau_index_v, const double *pi, const double *eta,
                           const int64_t *variants_v, int G, int S, int g,
                           double logp[4], double prob[4]);

/* ---- mu / E sufficient statistics (HaploSNP_Sampler.py:284-309, :266, :276) */
/* Per read categorical draw over strains with weights gamma[s,g]*eta[tau_vg,a]
 * (same joint law as the reference's two-stage multinomial).
 * sum_mu[S*G] and esum[16] (esum[a_obs*4+b_true]) are ADDED to. */
void oracle_mu_stats(const int64_t *tau, const double *gamma, const double *eta,
                     const int64_t *variants, int V, int G, int S,
                     uint64_t seed, uint32_t sweep, int64_t v0,
                     int64_t *sum_mu, int64_t *esum);

/* same statistics under the pattern-aggregated conditional-binomial contract (see desman_oracle.c) */
void oracle_mu_stats_agg(const int64_t *tau, const double *gamma, const double *eta,
                         const int64_t *variants, int V, int G, int S,
                         uint64_t seed, uint32_t sweep, int64_t v0,
                         int64_t *sum_mu, int64_t *esum);

/* ---- gamma / eta Dirichlet draws (HaploSNP_Sampler.py:263-281) ----------- */
void oracle_draw_gamma(const int64_t *sum_mu, int S, int G, double alpha, double epsilon,
                       uint64_t seed, uint32_t sweep, double *gamma);
void oracle_draw_eta(const int64_t *esum, double delta, uint64_t seed, uint32_t sweep, double *eta);
double oracle_gamma_variate(double shape, uint64_t seed, uint32_t sweep, uint32_t idx,
                            int stage, int boost_stage);

/* ---- log-likelihood / log-posterior (HaploSNP_Sampler.py:431-461) -------- */
double oracle_loglik(const int64_t *tau, const double *gamma, const double *eta,
                     const int64_t *variants, int V, int G, int S);
double oracle_logprior(const double *gamma, const double *eta, int V, int G, int S,
                       double alpha, double delta);

/* ---- whole sweep chain (HaploSNP_Sampler.py:334-365), Philox contract ---- */
typedef struct {
    int V, G, S, n_iter;
    double alpha, delta, epsilon;
    uint64_t seed;
    uint32_t sweep0;
    int mu_mode;   /* 0: per-read categorical contract, 1: pattern-aggregated binomial contract */
} oracle_chain_cfg;
/* Runs n_iter sweeps in place.  Outputs (any may be NULL): gamma_store[n_iter*S*G],
 * eta_store[n_iter*16], ll_store[n_iter], lp_store[n_iter], nchange_store[n_iter],
 * tau_sum[V*G*4] (int64, += one-hot per sweep), star state + lp_star/iter_star. */
void oracle_update(const oracle_chain_cfg *cfg, int64_t *tau, double *gamma, double *eta,
                   const int64_t *variants,
                   double *gamma_store, double *eta_store, double *ll_store, double *lp_store,
                   int64_t *nchange_store, int64_t *tau_sum,
                   int64_t *tau_star, double *gamma_star, double *eta_star,
                   double *lp_star, int *iter_star,
                   int64_t *sum_mu_last, int64_t *esum_last);
/* tau-only replay (HaploSNP_Sampler.py:383-407) with MT19937 or Philox words */
void oracle_update_tau(const oracle_chain_cfg *cfg, int use_mt, oracle_mt19937 *rng,
                       int64_t *tau, const double *gamma_store, const double *eta_store,
                       const int64_t *variants, double *ll_store, double *lp_store,
                       int64_t *nchange_store, int64_t *tau_sum, int64_t *tau_star, double *lp_star);

/* ---- NMFT (Init_NMFT.py) -------------------------------------------------- */
/* freq_matrix[4V,S] base-major rows (Init_NMFT.py:49-60) */
void   oracle_nmft_freq(const int64_t *snps, int V, int S, double *freq);
double oracle_nmft_objective(const double *freq, const double *tau, const double *gamma,
                             int V, int G, int S);
void   oracle_nmft_update(const double *freq, double *tau, double *gamma, int V, int G, int S,
                          int update_gamma, int update_tau);
/* factorize (:98-115) when fix_gamma==0, factorize_tau (:134-149) when fix_gamma==1.
 * tau[4V,G], gamma[G,S] hold the random initial factors on entry.
 * div_trace (may be NULL) receives div after every iteration, up to max_iter. */
int    oracle_nmft_factorize(const double *freq, double *tau, double *gamma, int V, int G, int S,
                             int max_iter, double min_change, int fix_gamma,
                             double *div_trace, double *div_final);
void   oracle_nmft_get_tau(const double *tau, int V, int G, int64_t *tau_onehot);

int oracle_num_threads(void);
void oracle_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
