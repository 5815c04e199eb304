/* gsl/gsl_rng.h -- minimal shim of the GSL random-number handle that CosmoPMC
 * threads through its PMC driver (exec/cosmo_pmc.c:530,604,739;
 * exec/exec_helper.c:14-36).  GSL is absent from this image; the host only
 * needs a seeded handle and a few uniform draws (initial proposal shifts,
 * revive_comp).  Host draws use SplitMix64; device draws use Philox keyed by
 * the handle's seed and a per-call stream counter. */
#ifndef PMCB200_GSL_RNG_H
#define PMCB200_GSL_RNG_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct { const char *name; } gsl_rng_type;
typedef struct {
  const gsl_rng_type *type;
  uint64_t seed;        /* as given to gsl_rng_set */
  uint64_t state;       /* SplitMix64 state of the host stream */
  uint32_t stream;      /* number of device sampling calls made with this handle */
} gsl_rng;
extern const gsl_rng_type *gsl_rng_default;
extern const gsl_rng_type *gsl_rng_mt19937;
gsl_rng *gsl_rng_alloc(const gsl_rng_type *T);
void gsl_rng_set(gsl_rng *r, unsigned long seed);
void gsl_rng_free(gsl_rng *r);
double gsl_rng_uniform(const gsl_rng *r);            /* [0,1) */
unsigned long gsl_rng_get(const gsl_rng *r);
const gsl_rng_type *gsl_rng_env_setup(void);
#ifdef __cplusplus
}
#endif
#endif
