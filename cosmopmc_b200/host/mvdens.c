/* mvdens.c -- host side of the mvdens / mix_mvdens API (include/pmctools/mvdens.h):
 * allocation, text I/O in the reference's formats, Cholesky / inverse for the
 * small (d <= 32) matrices, and the scalar single-point log-pdf that the
 * reference's host code calls for one point at a time (wrappers/src/param.c:1023,
 * 1489, 1534).  The batched N x K evaluations of the PMC iteration never come
 * here: they go through pmcb200_* (CUDA). */
#include "pmctools/mvdens.h"
#include <math.h>

#define LN2PI 1.8378770664093454836

static void set_views(mvdens *m)
{
   m->mean_view_container = gsl_vector_view_array(m->mean, m->ndim);
   m->x_tmp_view_container = gsl_vector_view_array(m->x_tmp, m->ndim);
   m->std_view_container = gsl_matrix_view_array(m->std, m->ndim, m->ndim);
   m->mean_view = &m->mean_view_container.vector;
   m->x_tmp_view = &m->x_tmp_view_container.vector;
   m->std_view = &m->std_view_container.matrix;
}

mvdens *mvdens_alloc(size_t ndim, error **err)
{
   mvdens *m = (mvdens *)calloc_err(1, sizeof(mvdens), err);
   forwardError(*err, __LINE__, NULL);
   m->ndim = ndim;
   m->buf = calloc_err(ndim * (ndim + 2), sizeof(double), err);
   forwardError(*err, __LINE__, NULL);
   m->own_buf = 1;
   m->mean = (double *)m->buf;
   m->std = m->mean + ndim;
   m->x_tmp = m->std + ndim * ndim;
   m->band_limit = (int)ndim;
   m->df = -1;
   m->chol = 0;
   m->detL = 0.0;
   set_views(m);
   return m;
}

void mvdens_empty(mvdens *m)
{
   memset(m->buf, 0, m->ndim * (m->ndim + 2) * sizeof(double));
   m->chol = 0; m->detL = 0.0;
}

void mvdens_free(mvdens **m)
{
   if (!m || !*m) return;
   if ((*m)->own_buf) free((*m)->buf);
   free(*m);
   *m = NULL;
}

/* mean and scale*var into m (not Cholesky-decomposed afterwards) */
void mvdens_from_meanvar(mvdens *m, const double *mean, const double *var, double scale)
{
   size_t d = m->ndim;
   if (mean) memcpy(m->mean, mean, d * sizeof(double)); else memset(m->mean, 0, d * sizeof(double));
   if (var) for (size_t i = 0; i < d * d; i++) m->std[i] = var[i] * scale;
   else for (size_t i = 0; i < d; i++) for (size_t j = 0; j < d; j++) m->std[i * d + j] = (i == j) ? scale : 0.0;
   m->chol = 0; m->detL = 0.0;
}

void mvdens_set_band_limit(mvdens *m, int band_limit) { m->band_limit = band_limit; }

double determinant(const double *L, size_t d)
{
   double det = 1.0;
   for (size_t i = 0; i < d; i++) det *= L[i * d + i];
   return det;
}

static int cholesky_inplace(double *A, size_t d)
{
   for (size_t j = 0; j < d; j++) {
      double s = A[j * d + j];
      for (size_t k = 0; k < j; k++) s -= A[j * d + k] * A[j * d + k];
      if (!(s > 0.0) || !isfinite(s)) return -1;
      double ljj = sqrt(s);
      A[j * d + j] = ljj;
      for (size_t i = j + 1; i < d; i++) {
         double t = A[i * d + j];
         for (size_t k = 0; k < j; k++) t -= A[i * d + k] * A[j * d + k];
         A[i * d + j] = t / ljj;
      }
   }
   for (size_t i = 0; i < d; i++) for (size_t j = i + 1; j < d; j++) A[i * d + j] = 0.0;
   return 0;
}

void mvdens_cholesky_decomp(mvdens *m, error **err)
{
   if (m->chol == 1) return;
   testErrorRet(cholesky_inplace(m->std, m->ndim) != 0, mv_cholesky,
                "Cholesky decomposition failed, matrix not positive definite", *err, __LINE__, );
   m->chol = 1;
   m->detL = determinant(m->std, m->ndim);
}

/* covariance (L L^T) of a density whose std may hold the Cholesky factor */
static void covariance_of(const mvdens *m, double *cov)
{
   size_t d = m->ndim;
   if (!m->chol) { memcpy(cov, m->std, d * d * sizeof(double)); return; }
   for (size_t i = 0; i < d; i++)
      for (size_t j = 0; j < d; j++) {
         double s = 0.0;
         size_t q = i < j ? i : j;
         for (size_t k = 0; k <= q; k++) s += m->std[i * d + k] * m->std[j * d + k];
         cov[i * d + j] = s;
      }
}

/* in-place inverse of the (symmetric positive definite) matrix in std via
 * Cholesky; returns the determinant of the input matrix.  Used on Fisher
 * matrices and inverse-covariance data files (param.c:521-528, bao.c:62, wmap.c:928). */
double mvdens_inverse(mvdens *m, error **err)
{
   size_t d = m->ndim;
   double *A = (double *)malloc_err(2 * d * d * sizeof(double), err);
   forwardError(*err, __LINE__, 0.0);
   double *Li = A + d * d;
   covariance_of(m, A);
   if (cholesky_inplace(A, d) != 0) {
      free(A);
      *err = addError(mv_cholesky, "Cannot invert: matrix not positive definite", *err, __LINE__);
      return 0.0;
   }
   double det = determinant(A, d); det *= det;
   /* Li = L^-1 (lower), inverse = Li^T Li */
   memset(Li, 0, d * d * sizeof(double));
   for (size_t c = 0; c < d; c++) {
      Li[c * d + c] = 1.0 / A[c * d + c];
      for (size_t i = c + 1; i < d; i++) {
         double s = 0.0;
         for (size_t k = c; k < i; k++) s -= A[i * d + k] * Li[k * d + c];
         Li[i * d + c] = s / A[i * d + i];
      }
   }
   for (size_t i = 0; i < d; i++)
      for (size_t j = 0; j < d; j++) {
         double s = 0.0;
         size_t q = i > j ? i : j;
         for (size_t k = q; k < d; k++) s += Li[k * d + i] * Li[k * d + j];
         m->std[i * d + j] = s;
      }
   m->chol = 0; m->detL = 0.0;
   free(A);
   return det;
}

/* single-point log-pdf (Cholesky on demand, as in the reference) */
double mvdens_log_pdf(mvdens *m, const double *x, error **err)
{
   size_t d = m->ndim;
   mvdens_cholesky_decomp(m, err);
   forwardError(*err, __LINE__, 0.0);
   double q = 0.0, logdet = 0.0;
   for (size_t i = 0; i < d; i++) {
      double t = x[i] - m->mean[i];
      for (size_t k = 0; k < i; k++) t -= m->std[i * d + k] * m->x_tmp[k];
      m->x_tmp[i] = t / m->std[i * d + i];
      q += m->x_tmp[i] * m->x_tmp[i];
      logdet += log(m->std[i * d + i]);
   }
   if (m->df <= 0) return -0.5 * (q + d * LN2PI) - logdet;
   double nu = (double)m->df;
   return lgamma(0.5 * (nu + d)) - lgamma(0.5 * nu) - 0.5 * d * log(nu * M_PI) - logdet
          - 0.5 * (nu + d) * log1p(q / nu);
}

double mvdens_log_pdf_void(void *m, const double *x, error **err)
{
   double r = mvdens_log_pdf((mvdens *)m, x, err);
   forwardError(*err, __LINE__, 0.0);
   return r;
}

/* one draw (host; used for set-up only, the PMC sampler is pmcb200_simulate) */
double *mvdens_ran(double *dest, mvdens *m, gsl_rng *r, error **err)
{
   size_t d = m->ndim;
   mvdens_cholesky_decomp(m, err);
   forwardError(*err, __LINE__, NULL);
   extern double gsl_ran_gaussian(const gsl_rng *, double);
   double u = 1.0;
   for (size_t i = 0; i < d; i++) m->x_tmp[i] = gsl_ran_gaussian(r, 1.0);
   if (m->df > 0) {
      double chi2 = 0.0;
      for (int i = 0; i < m->df; i++) { double g = gsl_ran_gaussian(r, 1.0); chi2 += g * g; }
      u = sqrt((double)m->df / chi2);
   }
   for (size_t i = 0; i < d; i++) {
      double t = 0.0;
      for (size_t k = 0; k <= i; k++) t += m->std[i * d + k] * m->x_tmp[k];
      dest[i] = m->mean[i] + u * t;
   }
   return dest;
}

/* ---- text format: header "p nu B c", mean row, p rows (manual.tex:3204-3234) ---- */
void mvdens_print(FILE *where, mvdens *m)
{
   FILE *F = where ? where : stdout;
   size_t d = m->ndim;
   fprintf(F, "%zu %d %d %d\n", d, m->df, m->band_limit, m->chol);
   for (size_t i = 0; i < d; i++) fprintf(F, "%g ", m->mean[i]);
   fprintf(F, "\n");
   for (size_t i = 0; i < d; i++) {
      for (size_t j = 0; j < d; j++) fprintf(F, "%g ", m->std[i * d + j]);
      fprintf(F, "\n");
   }
}

/* dump always writes the covariance (c = 0), never the Cholesky factor */
void mvdens_dump(FILE *where, mvdens *m)
{
   size_t d = m->ndim;
   double *cov = (double *)malloc(d * d * sizeof(double));
   covariance_of(m, cov);
   fprintf(where, "%zu %d %d %d\n", d, m->df, m->band_limit, 0);
   for (size_t i = 0; i < d; i++) fprintf(where, "%.*g ", PMC_DUMP_DIGITS, m->mean[i]);
   fprintf(where, "\n");
   for (size_t i = 0; i < d; i++) {
      for (size_t j = 0; j < d; j++) fprintf(where, "%.*g ", PMC_DUMP_DIGITS, cov[i * d + j]);
      fprintf(where, "\n");
   }
   free(cov);
}

void mvdens_chdump(const char *name, mvdens *m, error **err)
{
   FILE *F = fopen_err(name, "w", err);
   forwardError(*err, __LINE__, );
   mvdens_dump(F, m);
   fclose(F);
}

static int read_double(FILE *F, double *v) { return fscanf(F, "%lg", v) == 1; }

static void mvdens_read_body(FILE *F, mvdens *m, int chol, error **err)
{
   size_t d = m->ndim;
   for (size_t i = 0; i < d; i++)
      testErrorRet(!read_double(F, &m->mean[i]), mv_file, "Cannot read mvdens mean", *err, __LINE__, );
   for (size_t i = 0; i < d * d; i++)
      testErrorRet(!read_double(F, &m->std[i]), mv_file, "Cannot read mvdens matrix", *err, __LINE__, );
   m->chol = chol;
   if (chol) m->detL = determinant(m->std, d);
}

mvdens *mvdens_dwnp(FILE *F, error **err)
{
   long p; int df, B, c;
   testErrorRet(fscanf(F, "%ld %d %d %d", &p, &df, &B, &c) != 4, mv_file, "Cannot read mvdens header 'p nu B c'",
                *err, __LINE__, NULL);
   testErrorRetVA(p < 1 || p > 100000, mv_dimension, "Invalid mvdens dimension %ld", *err, __LINE__, NULL, p);
   mvdens *m = mvdens_alloc((size_t)p, err);
   forwardError(*err, __LINE__, NULL);
   m->df = df; m->band_limit = B;
   mvdens_read_body(F, m, c, err);
   if (isError(*err)) { mvdens_free(&m); forwardError(*err, __LINE__, NULL); }
   return m;
}

/* ---- mixtures --------------------------------------------------------------- */
size_t mix_mvdens_size(size_t ncomp, size_t ndim)
{
   return ncomp * (2 + ndim * (ndim + 2)) * sizeof(double);
}

mix_mvdens *mix_mvdens_alloc(size_t ncomp, size_t ndim, error **err)
{
   mix_mvdens *m = (mix_mvdens *)calloc_err(1, sizeof(mix_mvdens), err);
   forwardError(*err, __LINE__, NULL);
   m->ncomp = ncomp; m->ndim = ndim;
   /* one contiguous lump (the reference MPI-sends a mixture as a blob):
      wght[K], cwght[K], then per component mean, std, x_tmp */
   m->buf = calloc_err(1, mix_mvdens_size(ncomp, ndim), err);
   forwardError(*err, __LINE__, NULL);
   m->own_buf = 1;
   m->wght = (double *)m->buf;
   m->cwght = m->wght + ncomp;
   m->comp = (mvdens **)calloc_err(ncomp, sizeof(mvdens *), err);
   forwardError(*err, __LINE__, NULL);
   double *p = m->cwght + ncomp;
   for (size_t k = 0; k < ncomp; k++) {
      mvdens *c = (mvdens *)calloc_err(1, sizeof(mvdens), err);
      forwardError(*err, __LINE__, NULL);
      c->ndim = ndim; c->buf = p; c->own_buf = 0;
      c->mean = p; c->std = p + ndim; c->x_tmp = c->std + ndim * ndim;
      c->band_limit = (int)ndim; c->df = -1; c->chol = 0;
      set_views(c);
      m->comp[k] = c;
      p += ndim * (ndim + 2);
   }
   m->wght_view_container = gsl_vector_view_array(m->wght, ncomp);
   m->cwght_view_container = gsl_vector_view_array(m->cwght, ncomp);
   m->wght_view = &m->wght_view_container.vector;
   m->cwght_view = &m->cwght_view_container.vector;
   m->init_cwght = 0;
   return m;
}

void mix_mvdens_free(mix_mvdens **m)
{
   if (!m || !*m) return;
   for (size_t k = 0; k < (*m)->ncomp; k++) free((*m)->comp[k]);
   free((*m)->comp);
   if ((*m)->own_buf) free((*m)->buf);
   free(*m);
   *m = NULL;
}
void mix_mvdens_free_void(void **m) { mix_mvdens_free((mix_mvdens **)m); }

void mix_mvdens_copy(mix_mvdens *t, const mix_mvdens *s, error **err)
{
   testErrorRetVA(t->ncomp != s->ncomp || t->ndim != s->ndim, mv_dimension,
                  "Incompatible mixtures (%zu,%zu) vs (%zu,%zu)", *err, __LINE__, , t->ncomp, t->ndim, s->ncomp, s->ndim);
   memcpy(t->buf, s->buf, mix_mvdens_size(s->ncomp, s->ndim));
   for (size_t k = 0; k < s->ncomp; k++) {
      t->comp[k]->df = s->comp[k]->df; t->comp[k]->chol = s->comp[k]->chol;
      t->comp[k]->band_limit = s->comp[k]->band_limit; t->comp[k]->detL = s->comp[k]->detL;
   }
   t->init_cwght = s->init_cwght;
}

void mix_mvdens_cholesky_decomp(mix_mvdens *m, error **err)
{
   for (size_t k = 0; k < m->ncomp; k++) {
      mvdens_cholesky_decomp(m->comp[k], err);
      forwardError(*err, __LINE__, );
   }
}

void mix_mvdens_print(FILE *where, mix_mvdens *m)
{
   FILE *F = where ? where : stdout;
   fprintf(F, "%zu %zu\n", m->ncomp, m->ndim);
   for (size_t k = 0; k < m->ncomp; k++) {
      fprintf(F, "%g\n", m->wght[k]);
      mvdens_print(F, m->comp[k]);
   }
}

/* header "D p", then per component: weight line + mvdens (manual.tex:3236-3255) */
void mix_mvdens_dump(FILE *where, mix_mvdens *m)
{
   fprintf(where, "%zu %zu\n", m->ncomp, m->ndim);
   for (size_t k = 0; k < m->ncomp; k++) {
      fprintf(where, "%.*g\n", PMC_DUMP_DIGITS, m->wght[k]);
      mvdens_dump(where, m->comp[k]);
   }
}

mix_mvdens *mix_mvdens_dwnp(FILE *F, error **err)
{
   long D, p;
   testErrorRet(fscanf(F, "%ld %ld", &D, &p) != 2, mv_file, "Cannot read mix_mvdens header 'D p'", *err, __LINE__, NULL);
   testErrorRetVA(D < 1 || p < 1, mv_dimension, "Invalid mix_mvdens header %ld %ld", *err, __LINE__, NULL, D, p);
   mix_mvdens *m = mix_mvdens_alloc((size_t)D, (size_t)p, err);
   forwardError(*err, __LINE__, NULL);
   for (long k = 0; k < D; k++) {
      long pp; int df, B, c;
      if (!read_double(F, &m->wght[k]) || fscanf(F, "%ld %d %d %d", &pp, &df, &B, &c) != 4 || pp != p) {
         mix_mvdens_free(&m);
         *err = addErrorVA(mv_file, "Cannot read component %ld of mix_mvdens", *err, __LINE__, k);
         return NULL;
      }
      m->comp[k]->df = df; m->comp[k]->band_limit = B;
      mvdens_read_body(F, m->comp[k], c, err);
      if (isError(*err)) { mix_mvdens_free(&m); forwardError(*err, __LINE__, NULL); }
   }
   return m;
}

/* single-point mixture log-pdf: log sum_d alpha_d exp(log phi_d), no max-shift */
double mix_mvdens_log_pdf(mix_mvdens *m, const double *x, error **err)
{
   double s = 0.0;
   for (size_t k = 0; k < m->ncomp; k++) {
      if (m->wght[k] == 0.0) continue;
      double lp = mvdens_log_pdf(m->comp[k], x, err);
      forwardError(*err, __LINE__, 0.0);
      s += m->wght[k] * exp(lp);
   }
   return log(s);
}

double mix_mvdens_log_pdf_void(void *m, const double *x, error **err)
{
   double r = mix_mvdens_log_pdf((mix_mvdens *)m, x, err);
   forwardError(*err, __LINE__, 0.0);
   return r;
}

/* ENC = 1 / sum alpha_d^2 (manual.tex:599-603) */
double effective_number_of_components(const mix_mvdens *m, error **err)
{
   double s = 0.0;
   for (size_t k = 0; k < m->ncomp; k++) s += m->wght[k] * m->wght[k];
   testErrorRet(!(s > 0.0), mv_negWeight, "All component weights are zero", *err, __LINE__, 0.0);
   return 1.0 / s;
}
