/* pmctools/mvdens.h -- multivariate Gaussian / Student-t densities and their
 * mixtures: the `mvdens` / `mix_mvdens` structs of pmclib whose FIELDS the
 * reference pokes directly (wrappers/src/param.c:531-549,610,671-695;
 * exec/cosmo_pmc.c:212-235,275), their text formats
 * (Manual/manual.tex:3204-3255) and the scalar single-point helpers.
 * The batched (N x K) evaluations run on the GPU through include/pmcb200.h;
 * the functions here are set-up / I/O / single-point utilities of the host API. */
#ifndef PMCTOOLS_MVDENS_H
#define PMCTOOLS_MVDENS_H

#include <stdio.h>
#include <stddef.h>
#include "errorlist.h"
#include "gsl/gsl_rng.h"
#include "gsl/gsl_vector.h"

/* significant digits of mvdens_dump / mix_mvdens_dump: the reference writes %g
 * (6 digits, cf. Demo/MC_Demo/COSMOS-S10+SN+BAO/fisher) */
#ifndef PMC_DUMP_DIGITS
#define PMC_DUMP_DIGITS 6
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  size_t ndim;
  void  *buf;              /* one lump: mean[ndim], std[ndim*ndim], x_tmp[ndim] */
  int    own_buf;
  double *mean;            /* [ndim] */
  double *std;             /* [ndim*ndim] row-major: covariance, or its lower Cholesky
                              factor in place when chol == 1 */
  double *x_tmp;
  gsl_vector_view mean_view_container, x_tmp_view_container;
  gsl_matrix_view std_view_container;
  gsl_vector *mean_view, *x_tmp_view;
  gsl_matrix *std_view;
  int    band_limit;       /* number of secondary diagonals updated (B in the file header) */
  int    df;               /* -1 Gaussian, > 0 Student-t degrees of freedom */
  int    chol;             /* 1 if std holds the Cholesky factor */
  double detL;             /* determinant of L (valid when chol == 1) */
} mvdens;

typedef struct {
  size_t ncomp, ndim;
  void  *buf;
  int    own_buf;
  mvdens **comp;           /* [ncomp] */
  double *wght;            /* [ncomp] component weights alpha_d */
  double *cwght;           /* cumulative weights */
  gsl_vector_view wght_view_container, cwght_view_container;
  gsl_vector *wght_view, *cwght_view;
  int    init_cwght;
} mix_mvdens;

typedef double (posterior_log_pdf_func)(void *, const double *, error **);
typedef posterior_log_pdf_func log_pdf_func;
typedef void (retrieve_ded_func)(const void *, double *, error **);

/* mvdens */
mvdens *mvdens_alloc(size_t ndim, error **err);
void    mvdens_free(mvdens **m);
void    mvdens_empty(mvdens *m);
void    mvdens_from_meanvar(mvdens *m, const double *mean, const double *var, double scale);
void    mvdens_set_band_limit(mvdens *m, int band_limit);
void    mvdens_print(FILE *where, mvdens *m);
void    mvdens_dump(FILE *where, mvdens *m);                       /* text format, covariance */
void    mvdens_chdump(const char *name, mvdens *m, error **err);   /* dump to a named file */
mvdens *mvdens_dwnp(FILE *where, error **err);                     /* read */
void    mvdens_cholesky_decomp(mvdens *m, error **err);
double  mvdens_inverse(mvdens *m, error **err);                    /* in place; returns det of the input */
double  mvdens_log_pdf(mvdens *m, const double *x, error **err);
double  mvdens_log_pdf_void(void *m, const double *x, error **err);
double *mvdens_ran(double *dest, mvdens *m, gsl_rng *r, error **err);
double  determinant(const double *L, size_t ndim);

/* mix_mvdens */
mix_mvdens *mix_mvdens_alloc(size_t ncomp, size_t ndim, error **err);
void    mix_mvdens_free(mix_mvdens **m);
void    mix_mvdens_free_void(void **m);
void    mix_mvdens_copy(mix_mvdens *target, const mix_mvdens *source, error **err);
void    mix_mvdens_print(FILE *where, mix_mvdens *m);
void    mix_mvdens_dump(FILE *where, mix_mvdens *m);
mix_mvdens *mix_mvdens_dwnp(FILE *where, error **err);
void    mix_mvdens_cholesky_decomp(mix_mvdens *m, error **err);
double  mix_mvdens_log_pdf(mix_mvdens *m, const double *x, error **err);
double  mix_mvdens_log_pdf_void(void *m, const double *x, error **err);
double  effective_number_of_components(const mix_mvdens *m, error **err);
size_t  mix_mvdens_size(size_t ncomp, size_t ndim);

#ifdef __cplusplus
}
#endif
#endif
