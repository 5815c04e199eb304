/* gsl_shim.c -- see include/gsl/gsl_rng.h. */
#include "gsl/gsl_rng.h"
#include "gsl/gsl_randist.h"
#include "gsl/gsl_vector.h"
#include <math.h>
#include <stdlib.h>

static const gsl_rng_type t_default = {"pmcb200-splitmix64"};
const gsl_rng_type *gsl_rng_default = &t_default;
const gsl_rng_type *gsl_rng_mt19937 = &t_default;

const gsl_rng_type *gsl_rng_env_setup(void) { return gsl_rng_default; }

gsl_rng *gsl_rng_alloc(const gsl_rng_type *T)
{
   gsl_rng *r = (gsl_rng *)calloc(1, sizeof(gsl_rng));
   if (r) { r->type = T; gsl_rng_set(r, 0); }
   return r;
}
void gsl_rng_set(gsl_rng *r, unsigned long seed)
{
   r->seed = seed; r->state = 0x9E3779B97F4A7C15ull ^ (uint64_t)seed; r->stream = 0;
}
void gsl_rng_free(gsl_rng *r) { free(r); }

static uint64_t splitmix(gsl_rng *r)
{
   uint64_t z = (r->state += 0x9E3779B97F4A7C15ull);
   z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
   z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
   return z ^ (z >> 31);
}
unsigned long gsl_rng_get(const gsl_rng *r) { return (unsigned long)(splitmix((gsl_rng *)r) >> 32); }
double gsl_rng_uniform(const gsl_rng *r) { return (double)(splitmix((gsl_rng *)r) >> 11) * (1.0 / 9007199254740992.0); }
double gsl_ran_flat(const gsl_rng *r, double a, double b) { return a + (b - a) * gsl_rng_uniform(r); }
double gsl_ran_gaussian(const gsl_rng *r, double sigma)
{
   double u1 = 1.0 - gsl_rng_uniform(r), u2 = gsl_rng_uniform(r);
   return sigma * sqrt(-2.0 * log(u1)) * cos(2.0 * M_PI * u2);
}

gsl_vector_view gsl_vector_view_array(double *base, size_t n)
{
   gsl_vector_view v; v.vector.size = n; v.vector.stride = 1; v.vector.data = base; v.vector.block = 0; v.vector.owner = 0;
   return v;
}
gsl_matrix_view gsl_matrix_view_array(double *base, size_t n1, size_t n2)
{
   gsl_matrix_view m; m.matrix.size1 = n1; m.matrix.size2 = n2; m.matrix.tda = n2; m.matrix.data = base;
   m.matrix.block = 0; m.matrix.owner = 0;
   return m;
}
int gsl_vector_scale(gsl_vector *a, const double x)
{
   for (size_t i = 0; i < a->size; i++) a->data[i * a->stride] *= x;
   return 0;
}
double gsl_vector_get(const gsl_vector *v, size_t i) { return v->data[i * v->stride]; }
void gsl_vector_set(gsl_vector *v, size_t i, double x) { v->data[i * v->stride] = x; }
gsl_error_handler_t *gsl_set_error_handler_off(void) { return 0; }
