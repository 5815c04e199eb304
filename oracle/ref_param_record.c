/* ============================================================================
 * ref_param_record.c -- environment for running the REFERENCE'S OWN, UNCHANGED
 * wrappers/src/{param,sn,bao,wmap}.c as a bit-level checker (oracle/_ref/libref_param.so,
 * recipe oracle/build_ref_param.py).
 *
 * TEST INFRASTRUCTURE ONLY (see pmc_oracle.h).  Nothing here is a likelihood.
 *
 * The in-tree half of the posterior -- the `switch (like->par[i])` blocks of
 * likeli_SNIa / likeli_BAO / likeli_CMBDistPrior (sn.c:138-281, bao.c:80-184,
 * wmap.c:945-1049), set_base_parameters (param.c:1544-1661), the order of SetDl /
 * test_range_de_conservative / chi2_*, posterior_log_pdf_common (param.c:958-1041),
 * prior_log_pdf_special (param.c:1055-1101) and read_config_base's logpr_default
 * (param.c:124-129) -- is reference source that compiles here.  What it calls in the
 * absent nicaea library is replaced by RECORDING stubs: every stub stores the `cosmo`
 * / `cosmo_SN` it is handed, and the chi2_* stubs return values the test chose, so
 * that (1) the parameter mapping can be compared bit for bit with the oracle's and
 * the device's apply_params, and (2) the assembly of the posterior can be compared
 * bit for bit when the test feeds the oracle's own per-probe log-likelihoods back in.
 * ========================================================================== */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#include "nicaea/cosmo.h"
#include "nicaea/sn1a.h"
#include "nicaea/cmb_bao.h"
#include "param.h"          /* the reference's own header: config_base, posterior_log_pdf_common */

/* ---- the record ------------------------------------------------------------ */
enum { FN_SetDl = 1, FN_chi2_SN, FN_chi2_bao_A, FN_chi2_bao_d_z, FN_chi2_bao_D_V_ratio, FN_chi2_cmbDP,
       FN_test_range };
#define REC_LEN 24
#define REC_MAX 32
static double g_rec[REC_MAX][REC_LEN];
static int g_nrec = 0;
static double g_ret_sn = 0.0, g_ret_bao = 0.0, g_ret_cmb = 0.0;
static int g_de_prior = 0;          /* what test_range_de_conservative answers */
static int g_fail_setdl = 0;        /* SetDl raises an error (distance failure) */
static cosmo g_default;             /* what the *_to_default* constructors return */
static double g_default_Theta2[NLCP] = {0, 0, 0, 0};

static void record(int fn, const cosmo *c, const cosmo_SN *s, int a, int b)
{
   if (g_nrec >= REC_MAX) return;
   double *r = g_rec[g_nrec++];
   memset(r, 0, sizeof(double) * REC_LEN);
   r[0] = fn;
   r[1] = c->Omega_m; r[2] = c->Omega_de; r[3] = c->w0_de; r[4] = c->w1_de; r[5] = c->h_100;
   r[6] = c->Omega_b; r[7] = c->Omega_nu_mass; r[8] = c->Neff_nu_mass; r[9] = (double)c->de_param;
   if (s) {
      for (int i = 0; i < 4; i++) r[10 + i] = s->Theta2[i];
      r[14] = s->stretch; r[15] = s->color; r[16] = s->beta_d; r[17] = (double)s->chi2mode;
      r[18] = s->Theta2_denom[1]; r[19] = s->Theta2_denom[2];
   }
   r[20] = a; r[21] = b;
}

/* ---- nicaea-named constructors: plain struct management ----------------------- */
static cosmo *dup_cosmo(const cosmo *s, error **err)
{
   cosmo *c = (cosmo *)malloc_err(sizeof(cosmo), err);
   forwardError(*err, __LINE__, NULL);
   *c = *s;
   c->w_poly_de = NULL; c->N_poly_de = 0; c->tables = NULL;
   return c;
}
cosmo *copy_parameters_only(cosmo *source, error **err) { return dup_cosmo(source, err); }
cosmo *copy_parameters(cosmo *source, error **err) { return dup_cosmo(source, err); }
cosmo *set_cosmological_parameters_to_default2(error **err) { return dup_cosmo(&g_default, err); }
cosmo *set_cosmological_parameters_to_default(error **err) { return dup_cosmo(&g_default, err); }
void free_parameters(cosmo **self) { if (self && *self) { free(*self); *self = NULL; } }
void updateFrom(cosmo *avant, cosmo *apres, error **err) { (void)avant; (void)apres; (void)err; }
void dump_param(cosmo *self, FILE *F) { (void)self; (void)F; }
void read_cosmological_parameters(cosmo **self, FILE *F, error **err) { (void)F; *self = dup_cosmo(&g_default, err); }
double getH0fromCMB(double a, double b, double c, int d) { (void)a; (void)b; (void)c; (void)d; return 0.7; }

cosmo_SN *set_cosmological_parameters_to_default_SN(error **err)
{
   cosmo_SN *s = (cosmo_SN *)malloc_err(sizeof(cosmo_SN), err);
   forwardError(*err, __LINE__, NULL);
   memset(s, 0, sizeof(*s));
   s->cosmo = dup_cosmo(&g_default, err);
   forwardError(*err, __LINE__, NULL);
   for (int i = 0; i < NLCP; i++) s->Theta2[i] = g_default_Theta2[i];
   s->stretch = 1.0;
   return s;
}
void read_cosmological_parameters_SN(cosmo_SN **self, FILE *F, error **err)
{
   (void)F;
   *self = set_cosmological_parameters_to_default_SN(err);
}
cosmo_SN *copy_parameters_SN_only(cosmo_SN *src, error **err)
{
   cosmo_SN *s = (cosmo_SN *)malloc_err(sizeof(cosmo_SN), err);
   forwardError(*err, __LINE__, NULL);
   *s = *src;
   s->cosmo = dup_cosmo(src->cosmo, err);
   forwardError(*err, __LINE__, NULL);
   return s;
}
void free_parameters_SN(cosmo_SN **s) { if (s && *s) { free_parameters(&(*s)->cosmo); free(*s); *s = NULL; } }
SnSample *SnSample_read(const char *FileName, sndatformat_t f, error **err)
{
   (void)FileName; (void)f;
   SnSample *sn = (SnSample *)malloc_err(sizeof(SnSample), err);
   forwardError(*err, __LINE__, NULL);
   memset(sn, 0, sizeof(*sn));
   return sn;
}

/* ---- the recording stubs ------------------------------------------------------ */
void SetDl(cosmo_SN *self, SnSample *sn, error **err)
{
   (void)sn;
   record(FN_SetDl, self->cosmo, self, 0, 0);
   if (g_fail_setdl) *err = addError(ce_negative, "recorder: SetDl asked to fail", *err, __LINE__);
}
double chi2_SN(const cosmo_SN *m, const SnSample *sn, mvdens *data_beta_d, int wTheta1, int add_logdetCov, error **err)
{
   (void)sn; (void)data_beta_d; (void)err;
   record(FN_chi2_SN, m->cosmo, m, wTheta1, add_logdetCov);
   return g_ret_sn;
}
double chi2_bao_A(cosmo *model, mvdens *g, const double *z, error **err)
{ (void)err; record(FN_chi2_bao_A, model, NULL, g->ndim, (int)lround(1e6 * z[0])); return g_ret_bao; }
double chi2_bao_d_z(cosmo *model, mvdens *g, const double *z, error **err)
{ (void)err; record(FN_chi2_bao_d_z, model, NULL, g->ndim, (int)lround(1e6 * z[0])); return g_ret_bao; }
double chi2_bao_D_V_ratio(cosmo *model, mvdens *g, const double *z, error **err)
{ (void)err; record(FN_chi2_bao_D_V_ratio, model, NULL, g->ndim, (int)lround(1e6 * z[0])); return g_ret_bao; }
double chi2_cmbDP(cosmo *model, mvdens *g, error **err)
{ (void)err; record(FN_chi2_cmbDP, model, NULL, g->ndim, 0); return g_ret_cmb; }
int test_range_de_conservative(cosmo *model, error **err)
{ (void)err; record(FN_test_range, model, NULL, g_de_prior, 0); return g_de_prior; }

/* ---- driver API (ctypes) ------------------------------------------------------- */
#define REFP_MAXH 16
static config_base *g_cfg[REFP_MAXH];

static int take_error(error **err, char *msg, int msglen)
{
   if (!isError(*err)) return 0;
   int code = getErrorValue(*err);          /* the originating error, not the forward markers */
   if (msg && msglen > 0) stringError(msg, *err);
   purgeError(err);
   return code ? code : -1;
}

/* model[9]: Omega_m Omega_de w0 w1 h_100 Omega_b Omega_nu_mass Neff_nu_mass de_param */
void refp_set_default_model(const double *model, const double *Theta2)
{
   memset(&g_default, 0, sizeof(g_default));
   g_default.Omega_m = model[0]; g_default.Omega_de = model[1]; g_default.w0_de = model[2]; g_default.w1_de = model[3];
   g_default.h_100 = model[4]; g_default.Omega_b = model[5]; g_default.Omega_nu_mass = model[6];
   g_default.Neff_nu_mass = model[7]; g_default.de_param = (de_param_t)(int)model[8];
   g_default.nonlinear = smith03; g_default.transfer = eisenhu; g_default.growth = growth_de;
   for (int i = 0; i < NLCP; i++) g_default_Theta2[i] = Theta2 ? Theta2[i] : 0.0;
}
void refp_set_returns(double sn, double bao, double cmb, int de_prior, int fail_setdl)
{
   g_ret_sn = sn; g_ret_bao = bao; g_ret_cmb = cmb; g_de_prior = de_prior; g_fail_setdl = fail_setdl;
}

/* read_config_base (param.c:73-200) on a config file's base part, no_init = 0 */
int refp_open(const char *path, char *msg, int msglen)
{
   int h = 0;
   while (h < REFP_MAXH && g_cfg[h]) h++;
   if (h == REFP_MAXH) return -1;
   error *myerr = NULL, **err = &myerr;
   FILE *F = fopen(path, "r");
   if (!F) { if (msg) snprintf(msg, (size_t)msglen, "cannot open %s", path); return -2; }
   config_base *c = (config_base *)calloc(1, sizeof(config_base));
   read_config_base(c, F, 0, err);
   fclose(F);
   int code = take_error(err, msg, msglen);
   if (code) { free(c); return code < 0 ? code : -code; }
   g_cfg[h] = c;
   return h;
}
void refp_close(int h) { if (h >= 0 && h < REFP_MAXH) g_cfg[h] = NULL; /* the reference has no destructor for config_base */ }
double refp_logpr_default(int h) { return g_cfg[h]->logpr_default; }

static int copy_records(double *rec, int maxrec)
{
   int n = g_nrec < maxrec ? g_nrec : maxrec;
   if (rec) memcpy(rec, g_rec, sizeof(double) * REC_LEN * (size_t)n);
   return g_nrec;
}
/* posterior_log_pdf_common (param.c:958-1041); *errcode = 0 or the reference's error value */
double refp_posterior(int h, const double *x, int *errcode, double *rec, int maxrec, int *nrec)
{
   error *myerr = NULL, **err = &myerr;
   g_nrec = 0;
   double r = posterior_log_pdf_common(g_cfg[h], x, err);
   *errcode = take_error(err, NULL, 0);
   if (nrec) *nrec = copy_records(rec, maxrec);
   return r;
}
/* one data set's func_likeli through likelihood_log_pdf_single (param.c:930-946) */
double refp_likeli(int h, int idata, const double *x, int *errcode, double *rec, int maxrec, int *nrec)
{
   error *myerr = NULL, **err = &myerr;
   g_nrec = 0;
   double r = likelihood_log_pdf_single(g_cfg[h]->data_extra[idata], x, err);
   *errcode = take_error(err, NULL, 0);
   if (nrec) *nrec = copy_records(rec, maxrec);
   return r;
}
/* prior_log_pdf_special (param.c:1055-1101) */
double refp_prior_special(int special, const int *par, const double *min, const double *max, int npar, int *errcode)
{
   error *myerr = NULL, **err = &myerr;
   par_t p[64];
   for (int i = 0; i < npar && i < 64; i++) p[i] = (par_t)par[i];
   double r = prior_log_pdf_special((special_t)special, p, min, max, npar, err);
   *errcode = take_error(err, NULL, 0);
   return r;
}
int refp_record_len(void) { return REC_LEN; }
