/* pmclib/mcmc.h -- error codes of pmclib's Metropolis-Hastings module that the
 * reference's wrappers use as generic "MC" codes (wrappers/src/param.c,
 * exec/cosmo_pmc.c:482-492).  The MCMC sampler itself is out of scope. */
#ifndef PMCLIB_MCMC_H
#define PMCLIB_MCMC_H
#include "pmctools/errorlist.h"
#include "pmctools/mvdens.h"
#define mcmc_base      (-1300)
#define mcmc_allocate  (-1 + mcmc_base)
#define mcmc_infnan    (-2 + mcmc_base)
#define mcmc_negative  (-3 + mcmc_base)
#define mcmc_prior     (-4 + mcmc_base)
#define mcmc_dimension (-5 + mcmc_base)
#define mcmc_file      (-6 + mcmc_base)
#define mcmc_unknown   (-7 + mcmc_base)
#define mcmc_outOfBound (-8 + mcmc_base)
#define mcmc_tooManySteps (-9 + mcmc_base)
#define mcmc_undef     (-10 + mcmc_base)
#define MC_AF_ACCEPT 1
#endif
