/* oracle/gsl_shim/gsl/gsl_rng.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Stand-in for the seven GSL symbols that the reference's
 * sampletau/c_sample_tau.c (lines 24-44 and 174) uses, so that the reference
 * file compiles UNMODIFIED into oracle/_ref/ on an image without GSL.
 *
 * GSL (not vendored by the reference; setup.py:21 links "gsl", no version pin)
 * documents gsl_rng_mt19937 as the 2002 Matsumoto-Nishimura MT19937 with
 * init_genrand seeding, seed 0 replaced by 4357, and gsl_rng_uniform =
 * 32-bit output / 2^32.  This file restates that published algorithm; it is
 * checked in tests/test_oracle_cpu.py against numpy's MT19937 legacy seeding.
 */
#ifndef ORACLE_GSL_RNG_SHIM_H
#define ORACLE_GSL_RNG_SHIM_H
#include <stdlib.h>

typedef struct { unsigned long mt[624]; int mti; } gsl_rng;
typedef struct { int unused; } gsl_rng_type;

static const gsl_rng_type gsl_rng_mt19937_obj = {0};
#define gsl_rng_mt19937 (&gsl_rng_mt19937_obj)

static inline void gsl_rng_env_setup(void) {}

static inline void gsl_rng_set(gsl_rng *r, unsigned long int s)
{
    int i;
    if (s == 0) s = 4357;                      /* GSL convention for seed 0 */
    r->mt[0] = s & 0xffffffffUL;
    for (i = 1; i < 624; i++)
        r->mt[i] = (1812433253UL * (r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) + (unsigned long)i) & 0xffffffffUL;
    r->mti = 624;
}

static inline gsl_rng *gsl_rng_alloc(const gsl_rng_type *T)
{
    gsl_rng *r = (gsl_rng *)malloc(sizeof(gsl_rng));
    (void)T;
    if (r) gsl_rng_set(r, 0);                  /* gsl_rng_alloc seeds with the default seed 0 */
    return r;
}

static inline void gsl_rng_free(gsl_rng *r) { free(r); }

static inline unsigned long gsl_rng_shim_next(gsl_rng *r)
{
    unsigned long y;
    if (r->mti >= 624) {
        int k;
        for (k = 0; k < 624; k++) {
            unsigned long a = r->mt[k], b = r->mt[(k + 1) % 624], c = r->mt[(k + 397) % 624];
            y = (a & 0x80000000UL) | (b & 0x7fffffffUL);
            r->mt[k] = c ^ (y >> 1) ^ ((y & 1UL) ? 0x9908b0dfUL : 0UL);
        }
        r->mti = 0;
    }
    y = r->mt[r->mti++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680UL;
    y ^= (y << 15) & 0xefc60000UL;
    y ^= (y >> 18);
    return y & 0xffffffffUL;
}

static inline double gsl_rng_uniform(gsl_rng *r) { return gsl_rng_shim_next(r) / 4294967296.0; }

#endif
