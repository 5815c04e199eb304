/* generated: functions of exec_helper.c's environment that the post-processing tests never reach */
#include <stdio.h>
#include <stdlib.h>
void estimate_param_covar_weight(void) { fprintf(stderr, "oracle/_ref: stub estimate_param_covar_weight called\n"); abort(); }
void gsl_rng_alloc(void) { fprintf(stderr, "oracle/_ref: stub gsl_rng_alloc called\n"); abort(); }
void gsl_rng_default(void) { fprintf(stderr, "oracle/_ref: stub gsl_rng_default called\n"); abort(); }
void gsl_rng_set(void) { fprintf(stderr, "oracle/_ref: stub gsl_rng_set called\n"); abort(); }
void mean_from_psim(void) { fprintf(stderr, "oracle/_ref: stub mean_from_psim called\n"); abort(); }
void mvdens_alloc(void) { fprintf(stderr, "oracle/_ref: stub mvdens_alloc called\n"); abort(); }
void mvdens_dump(void) { fprintf(stderr, "oracle/_ref: stub mvdens_dump called\n"); abort(); }
void mvdens_from_meanvar(void) { fprintf(stderr, "oracle/_ref: stub mvdens_from_meanvar called\n"); abort(); }
void mvdens_inverse(void) { fprintf(stderr, "oracle/_ref: stub mvdens_inverse called\n"); abort(); }
