/* gsl/gsl_randist.h -- shim: gsl_ran_flat (wrappers/src/param.c:481,670,678),
 * gsl_ran_gaussian. */
#ifndef PMCB200_GSL_RANDIST_H
#define PMCB200_GSL_RANDIST_H
#include "gsl_rng.h"
#ifdef __cplusplus
extern "C" {
#endif
double gsl_ran_flat(const gsl_rng *r, double a, double b);
double gsl_ran_gaussian(const gsl_rng *r, double sigma);
#ifdef __cplusplus
}
#endif
#endif
