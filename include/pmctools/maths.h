/* pmctools/maths.h -- Numerical-Recipes style helpers of pmclib used by the
 * reference host code (wrappers/src/param.c:566-621,1215-1357;
 * exec/exec_helper.c; wrappers/src/sn.c:80-96).  Host-side set-up only. */
#ifndef PMCTOOLS_MATHS_H
#define PMCTOOLS_MATHS_H
#include <stdio.h>
#include <stdlib.h>
#include "errorlist.h"
#include "maths_base.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef error sm2_error;

typedef double my_complex[2];
typedef double (funcwithpars)(double, void *, error **);

/* equidistant 1-d interpolation table (e.g. z -> A_V, wrappers/src/sn.c:80-96) */
typedef struct { double *table; double a, b, dx, lower, upper; int n; } interTable;
interTable *init_interTable(int n, double a, double b, double dx, double lower, double upper, error **err);
void del_interTable(interTable **self);
double interpol_wr(interTable *self, double x, error **err);

/* NR-style offset vectors / matrices */
double *sm2_vector(long nl, long nh, error **err);
void    sm2_free_vector(double *v, long nl, long nh);
double **sm2_matrix(long nrl, long nrh, long ncl, long nch, error **err);
void    sm2_free_matrix(double **m, long nrl, long nrh, long ncl, long nch);
/* in-place inverse of a dense n x n row-major matrix; returns the determinant */
double  sm2_inverse(double *C, int N, error **err);
/* Jacobi eigen-decomposition of a symmetric n x n row-major matrix a (destroyed):
 * eigenvalues d[1..n], eigenvector k in row v[k][1..n] (NR offset arrays), as the
 * reference indexes them (wrappers/src/param.c:630-632) */
void    jacobi_transform(double *a, int n, double *d, double **v, int *nrot, error **err);
/* NR indexx: 1-based arrays arr[1..n], indx[1..n]; ascending */
void    indexx(unsigned long n, double arr[], unsigned long indx[], error **err);
/* mixed second derivative d^2 f / dx_a dx_b by Ridders' extrapolation of the 4-point
 * central stencil (exec/go_fishing.c:26); errn = error estimate */
double  nd_dfridr2(double (*func)(void *, const double *, error **), int a, int b, double *x, double ha, double hb,
                   void *extra, double *errn, error **err);
#ifdef __cplusplus
}
#endif
#endif
