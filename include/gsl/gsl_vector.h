/* gsl/gsl_vector.h -- shim: the strided vector view used for
 * `gsl_vector_scale(proposal->wght_view, 1.0/wght_sum)` (exec/cosmo_pmc.c:275). */
#ifndef PMCB200_GSL_VECTOR_H
#define PMCB200_GSL_VECTOR_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct { size_t size, stride; double *data; void *block; int owner; } gsl_vector;
typedef struct { gsl_vector vector; } gsl_vector_view;
typedef struct { size_t size1, size2, tda; double *data; void *block; int owner; } gsl_matrix;
typedef struct { gsl_matrix matrix; } gsl_matrix_view;
gsl_vector_view gsl_vector_view_array(double *base, size_t n);
gsl_matrix_view gsl_matrix_view_array(double *base, size_t n1, size_t n2);
int gsl_vector_scale(gsl_vector *a, const double x);
double gsl_vector_get(const gsl_vector *v, size_t i);
void gsl_vector_set(gsl_vector *v, size_t i, double x);
typedef void gsl_error_handler_t(const char *, const char *, int, int);
gsl_error_handler_t *gsl_set_error_handler_off(void);
#ifdef __cplusplus
}
#endif
#endif
