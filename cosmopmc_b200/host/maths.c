/* maths.c -- pmctools/maths.h: Numerical-Recipes style host helpers used for
 * set-up (initial proposal from the Fisher matrix, histograms, dust table). */
#include "pmctools/maths.h"
#include <math.h>
#include <string.h>

interTable *init_interTable(int n, double a, double b, double dx, double lower, double upper, error **err)
{
   interTable *t = (interTable *)malloc_err(sizeof(interTable), err);
   forwardError(*err, __LINE__, NULL);
   t->table = (double *)calloc_err(n > 0 ? n : 1, sizeof(double), err);
   forwardError(*err, __LINE__, NULL);
   t->n = n; t->a = a; t->b = b; t->dx = dx; t->lower = lower; t->upper = upper;
   return t;
}

void del_interTable(interTable **t)
{
   if (!t || !*t) return;
   free((*t)->table); free(*t); *t = NULL;
}

/* linear interpolation on the equidistant grid a + i dx */
double interpol_wr(interTable *t, double x, error **err)
{
   testErrorRetVA(x < t->a - 1e-12 || x > t->b + 1e-12, math_interpoloutofrange, "x = %g outside [%g, %g]", *err,
                  __LINE__, 0.0, x, t->a, t->b);
   double u = (x - t->a) / t->dx;
   int i = (int)floor(u);
   if (i < 0) i = 0;
   if (i > t->n - 2) i = t->n - 2;
   if (t->n < 2) return t->table[0];
   double f = u - i;
   return (1.0 - f) * t->table[i] + f * t->table[i + 1];
}

double *sm2_vector(long nl, long nh, error **err)
{
   double *v = (double *)calloc_err((size_t)(nh - nl + 2), sizeof(double), err);
   forwardError(*err, __LINE__, NULL);
   return v - nl + 1;
}
void sm2_free_vector(double *v, long nl, long nh) { (void)nh; if (v) free(v + nl - 1); }

double **sm2_matrix(long nrl, long nrh, long ncl, long nch, error **err)
{
   long nrow = nrh - nrl + 1, ncol = nch - ncl + 1;
   double **m = (double **)calloc_err((size_t)(nrow + 1), sizeof(double *), err);
   forwardError(*err, __LINE__, NULL);
   m += 1; m -= nrl;
   m[nrl] = (double *)calloc_err((size_t)(nrow * ncol + 1), sizeof(double), err);
   forwardError(*err, __LINE__, NULL);
   m[nrl] += 1; m[nrl] -= ncl;
   for (long i = nrl + 1; i <= nrh; i++) m[i] = m[i - 1] + ncol;
   return m;
}
void sm2_free_matrix(double **m, long nrl, long nrh, long ncl, long nch)
{
   (void)nrh; (void)nch;
   if (!m) return;
   free(m[nrl] + ncl - 1);
   free(m + nrl - 1);
}

/* Gauss-Jordan with partial pivoting, in place; returns the determinant */
double sm2_inverse(double *C, int N, error **err)
{
   double det = 1.0;
   int *piv = (int *)malloc_err(sizeof(int) * (size_t)N, err);
   forwardError(*err, __LINE__, 0.0);
   double *A = (double *)malloc_err(sizeof(double) * (size_t)N * 2 * N, err);
   forwardError(*err, __LINE__, 0.0);
   for (int i = 0; i < N; i++)
      for (int j = 0; j < N; j++) { A[i * 2 * N + j] = C[i * N + j]; A[i * 2 * N + N + j] = (i == j) ? 1.0 : 0.0; }
   for (int c = 0; c < N; c++) {
      int p = c;
      for (int r = c + 1; r < N; r++) if (fabs(A[r * 2 * N + c]) > fabs(A[p * 2 * N + c])) p = r;
      if (A[p * 2 * N + c] == 0.0 || !isfinite(A[p * 2 * N + c])) {
         free(A); free(piv);
         *err = addError(math_singularValue, "Singular matrix", *err, __LINE__);
         return 0.0;
      }
      if (p != c) { for (int j = 0; j < 2 * N; j++) { double t = A[c * 2 * N + j]; A[c * 2 * N + j] = A[p * 2 * N + j]; A[p * 2 * N + j] = t; } det = -det; }
      double d = A[c * 2 * N + c];
      det *= d;
      for (int j = 0; j < 2 * N; j++) A[c * 2 * N + j] /= d;
      for (int r = 0; r < N; r++) {
         if (r == c) continue;
         double f = A[r * 2 * N + c];
         if (f != 0.0) for (int j = 0; j < 2 * N; j++) A[r * 2 * N + j] -= f * A[c * 2 * N + j];
      }
   }
   for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) C[i * N + j] = A[i * 2 * N + N + j];
   free(A); free(piv);
   return det;
}

/* cyclic Jacobi for a symmetric matrix; eigenvector k in row v[k][1..n] */
void jacobi_transform(double *a, int n, double *d, double **v, int *nrot, error **err)
{
   #define A(i, j) a[((i) - 1) * n + ((j) - 1)]
   double *b = sm2_vector(1, n, err), *z = sm2_vector(1, n, err);
   forwardError(*err, __LINE__, );
   for (int ip = 1; ip <= n; ip++) { for (int iq = 1; iq <= n; iq++) v[ip][iq] = 0.0; v[ip][ip] = 1.0; }
   for (int ip = 1; ip <= n; ip++) { b[ip] = d[ip] = A(ip, ip); z[ip] = 0.0; }
   *nrot = 0;
   for (int sweep = 1; sweep <= 100; sweep++) {
      double sm = 0.0;
      for (int ip = 1; ip <= n - 1; ip++) for (int iq = ip + 1; iq <= n; iq++) sm += fabs(A(ip, iq));
      if (sm == 0.0) { sm2_free_vector(z, 1, n); sm2_free_vector(b, 1, n); return; }
      double tresh = sweep < 4 ? 0.2 * sm / (n * n) : 0.0;
      for (int ip = 1; ip <= n - 1; ip++) {
         for (int iq = ip + 1; iq <= n; iq++) {
            double g = 100.0 * fabs(A(ip, iq));
            if (sweep > 4 && fabs(d[ip]) + g == fabs(d[ip]) && fabs(d[iq]) + g == fabs(d[iq])) A(ip, iq) = 0.0;
            else if (fabs(A(ip, iq)) > tresh) {
               double h = d[iq] - d[ip], t;
               if (fabs(h) + g == fabs(h)) t = A(ip, iq) / h;
               else { double theta = 0.5 * h / A(ip, iq); t = 1.0 / (fabs(theta) + sqrt(1.0 + theta * theta)); if (theta < 0.0) t = -t; }
               double c = 1.0 / sqrt(1 + t * t), s = t * c, tau = s / (1.0 + c);
               h = t * A(ip, iq);
               z[ip] -= h; z[iq] += h; d[ip] -= h; d[iq] += h;
               A(ip, iq) = 0.0;
               #define ROT(i, j, k, l) { double g_ = A(i, j), h_ = A(k, l); A(i, j) = g_ - s * (h_ + g_ * tau); A(k, l) = h_ + s * (g_ - h_ * tau); }
               for (int j = 1; j <= ip - 1; j++) ROT(j, ip, j, iq)
               for (int j = ip + 1; j <= iq - 1; j++) ROT(ip, j, j, iq)
               for (int j = iq + 1; j <= n; j++) ROT(ip, j, iq, j)
               #undef ROT
               for (int j = 1; j <= n; j++) {     /* rows of v are the eigenvectors */
                  double g_ = v[ip][j], h_ = v[iq][j];
                  v[ip][j] = g_ - s * (h_ + g_ * tau); v[iq][j] = h_ + s * (g_ - h_ * tau);
               }
               ++(*nrot);
            }
         }
      }
      for (int ip = 1; ip <= n; ip++) { b[ip] += z[ip]; d[ip] = b[ip]; z[ip] = 0.0; }
   }
   sm2_free_vector(z, 1, n); sm2_free_vector(b, 1, n);
   *err = addError(math_tooManySteps, "Too many iterations in jacobi_transform", *err, __LINE__);
   #undef A
}

/* index table: arr[indx[1]] <= arr[indx[2]] <= ... (1-based) */
void indexx(unsigned long n, double arr[], unsigned long indx[], error **err)
{
   (void)err;
   for (unsigned long j = 1; j <= n; j++) indx[j] = j;
   for (unsigned long i = 2; i <= n; i++) {          /* insertion sort: n is the number of parameters */
      unsigned long t = indx[i]; double a = arr[t]; unsigned long j = i;
      while (j > 1 && arr[indx[j - 1]] > a) { indx[j] = indx[j - 1]; j--; }
      indx[j] = t;
   }
}

/* Ridders' method (NR dfridr) on the symmetric 4-point estimate of d^2 f/dx_a dx_b */
static double mixed_stencil(double (*func)(void *, const double *, error **), int a, int b, double *x, double ha,
                            double hb, void *extra, error **err)
{
   double xa = x[a], xb = x[b], f[4];
   const int sg[4][2] = {{+1, +1}, {+1, -1}, {-1, +1}, {-1, -1}};
   for (int j = 0; j < 4; j++) {
      x[a] = xa; x[b] = xb;
      x[a] += sg[j][0] * ha;
      x[b] += sg[j][1] * hb;      /* a == b: x_a +- 2h or x_a, the second-difference stencil */
      f[j] = func(extra, x, err);
      if (isError(*err)) { x[a] = xa; x[b] = xb; return 0.0; }
   }
   x[a] = xa; x[b] = xb;
   return (f[0] - f[1] - f[2] + f[3]) / (4.0 * ha * hb);
}

double nd_dfridr2(double (*func)(void *, const double *, error **), int a, int b, double *x, double ha, double hb,
                  void *extra, double *errn, error **err)
{
   enum { NTAB = 8 };
   const double CON = 1.4, CON2 = CON * CON, SAFE = 2.0;
   double A[NTAB][NTAB], ans = 0.0;
   testErrorRet(ha == 0.0 || hb == 0.0, math_wrongValue, "Step size must be non-zero", *err, __LINE__, 0.0);
   A[0][0] = mixed_stencil(func, a, b, x, ha, hb, extra, err);
   forwardError(*err, __LINE__, 0.0);
   *errn = 1.0e30;
   for (int i = 1; i < NTAB; i++) {
      ha /= CON; hb /= CON;
      A[0][i] = mixed_stencil(func, a, b, x, ha, hb, extra, err);
      forwardError(*err, __LINE__, 0.0);
      double fac = CON2;
      for (int j = 1; j <= i; j++) {
         A[j][i] = (A[j - 1][i] * fac - A[j - 1][i - 1]) / (fac - 1.0);
         fac = CON2 * fac;
         double errt = fmax(fabs(A[j][i] - A[j - 1][i]), fabs(A[j][i] - A[j - 1][i - 1]));
         if (errt <= *errn) { *errn = errt; ans = A[j][i]; }
      }
      if (fabs(A[i][i] - A[i - 1][i - 1]) >= SAFE * (*errn)) break;
   }
   return ans;
}
