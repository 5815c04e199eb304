/* nicaea.c -- the nicaea-named host functions that the reference's SNIa / BAO /
 * CMBDistPrior wrappers call (include/nicaea/{cosmo,sn1a,cmb_bao}.h): parameter
 * file readers, model copies, the SN data reader, and the scalar log-likelihoods,
 * which evaluate ONE model by launching the batched CUDA kernels with N = 1
 * (pmc_b200_single_loglike) -- there is no CPU implementation of the distance
 * integrals in the product. */
#include "nicaea/cosmo.h"
#include "nicaea/sn1a.h"
#include "nicaea/cmb_bao.h"
#include "pmclib/pmc.h"
#include <ctype.h>
#include <math.h>
#include <string.h>

/* ---- "key value(s) # comment" parameter files (par_files/cosmo.par) ---------------- */
typedef struct { char key[64]; char val[8][128]; int nval; } kv_t;

static int read_kv(FILE *F, kv_t *out, int maxn)
{
   char line[4096];
   int n = 0;
   while (n < maxn && fgets(line, sizeof(line), F)) {
      char *h = strchr(line, '#');
      if (h) *h = 0;
      char *tok = strtok(line, " \t\r\n");
      if (!tok) continue;
      snprintf(out[n].key, sizeof(out[n].key), "%s", tok);
      out[n].nval = 0;
      while ((tok = strtok(NULL, " \t\r\n")) && out[n].nval < 8) snprintf(out[n].val[out[n].nval++], 128, "%s", tok);
      n++;
   }
   return n;
}
static const kv_t *find_kv(const kv_t *kv, int n, const char *key)
{
   for (int i = 0; i < n; i++) if (!strcmp(kv[i].key, key)) return &kv[i];
   return NULL;
}
static double kv_double(const kv_t *kv, int n, const char *key, double dflt, int required, error **err)
{
   const kv_t *e = find_kv(kv, n, key);
   if (!e || e->nval < 1) {
      if (required) *err = addErrorVA(ce_file, "Key '%s' not found in parameter file", *err, __LINE__, key);
      return dflt;
   }
   return atof(e->val[0]);
}

cosmo *init_parameters(double OMEGAM, double OMEGADE, double W0_DE, double W1_DE, double *W_POLY_DE, int N_POLY_DE,
                       double H100, double OMEGAB, double OMEGANUMASS, double NEFFNUMASS, double NORM, double NSPEC,
                       nonlinear_t NONLINEAR, transfer_t TRANSFER, growth_t GROWTH, de_param_t DEPARAM,
                       norm_t normmode, double AMIN, error **err)
{
   cosmo *c = (cosmo *)calloc_err(1, sizeof(cosmo), err);
   forwardError(*err, __LINE__, NULL);
   c->Omega_m = OMEGAM; c->Omega_de = OMEGADE; c->w0_de = W0_DE; c->w1_de = W1_DE;
   c->N_poly_de = N_POLY_DE;
   if (N_POLY_DE > 0 && W_POLY_DE) {
      c->w_poly_de = (double *)malloc_err(sizeof(double) * N_POLY_DE, err);
      forwardError(*err, __LINE__, NULL);
      memcpy(c->w_poly_de, W_POLY_DE, sizeof(double) * N_POLY_DE);
   }
   c->h_100 = H100; c->Omega_b = OMEGAB; c->Omega_nu_mass = OMEGANUMASS; c->Neff_nu_mass = NEFFNUMASS;
   c->normalization = NORM; c->n_spec = NSPEC; c->normmode = normmode;
   if (normmode == norm_s8) c->sigma_8 = NORM; else c->As = NORM;
   c->nonlinear = NONLINEAR; c->transfer = TRANSFER; c->growth = GROWTH; c->de_param = DEPARAM; c->a_min = AMIN;
   return c;
}

cosmo *copy_parameters_only(cosmo *s, error **err)
{
   cosmo *c = init_parameters(s->Omega_m, s->Omega_de, s->w0_de, s->w1_de, s->w_poly_de, s->N_poly_de, s->h_100,
                              s->Omega_b, s->Omega_nu_mass, s->Neff_nu_mass, s->normalization, s->n_spec,
                              s->nonlinear, s->transfer, s->growth, s->de_param, (norm_t)s->normmode, s->a_min, err);
   forwardError(*err, __LINE__, NULL);
   c->sigma_8 = s->sigma_8; c->As = s->As;
   return c;
}
cosmo *copy_parameters(cosmo *s, error **err) { return copy_parameters_only(s, err); }

cosmo *set_cosmological_parameters_to_default2(error **err)
{
   return init_parameters(0.25, 0.75, -1.0, 0.0, NULL, 0, 0.70, 0.044, 0.0, 0.0, 0.80, 0.96, smith03, eisenhu,
                          growth_de, linder, norm_s8, 0.0, err);
}
cosmo *set_cosmological_parameters_to_default(error **err) { return set_cosmological_parameters_to_default2(err); }

void free_parameters(cosmo **c)
{
   if (!c || !*c) return;
   free((*c)->w_poly_de); free(*c); *c = NULL;
}

/* nicaea deletes pre-computed tables when parameters change; the device path keeps none */
void updateFrom(cosmo *avant, cosmo *apres, error **err) { (void)avant; (void)apres; (void)err; }

#define ENUM_FROM_STRING(dst, str, type, sfun, N, what)                                              \
   { int j_, ok_ = 0; for (j_ = 0; j_ < (N); j_++) if (!strcmp((str), sfun(j_))) { (dst) = (type)j_; ok_ = 1; } \
     testErrorRetVA(!ok_, ce_unknown, "Unknown %s '%s'", *err, __LINE__, , what, (str)); }

void read_cosmological_parameters(cosmo **self, FILE *F, error **err)
{
   static kv_t kv[128];
   int n = read_kv(F, kv, 128);
   cosmo *c = set_cosmological_parameters_to_default2(err);
   forwardError(*err, __LINE__, );
   c->Omega_m = kv_double(kv, n, "Omega_m", c->Omega_m, 1, err);            forwardError(*err, __LINE__, );
   c->Omega_de = kv_double(kv, n, "Omega_de", c->Omega_de, 1, err);         forwardError(*err, __LINE__, );
   c->w0_de = kv_double(kv, n, "w0_de", c->w0_de, 0, err);
   c->w1_de = kv_double(kv, n, "w1_de", c->w1_de, 0, err);
   c->h_100 = kv_double(kv, n, "h_100", c->h_100, 1, err);                  forwardError(*err, __LINE__, );
   c->Omega_b = kv_double(kv, n, "Omega_b", c->Omega_b, 1, err);            forwardError(*err, __LINE__, );
   c->Omega_nu_mass = kv_double(kv, n, "Omega_nu_mass", 0.0, 0, err);
   c->Neff_nu_mass = kv_double(kv, n, "Neff_nu_mass", 0.0, 0, err);
   c->normalization = kv_double(kv, n, "normalization", c->normalization, 0, err);
   c->n_spec = kv_double(kv, n, "n_spec", c->n_spec, 0, err);
   c->normmode = (int)kv_double(kv, n, "normmode", 0, 0, err);
   if (c->normmode == norm_s8) c->sigma_8 = c->normalization; else c->As = c->normalization;
   c->a_min = kv_double(kv, n, "a_min", 0.0, 0, err);
   const kv_t *e;
   if ((e = find_kv(kv, n, "snonlinear")) && e->nval) ENUM_FROM_STRING(c->nonlinear, e->val[0], nonlinear_t, snonlinear_t, Nnonlinear_t, "snonlinear");
   if ((e = find_kv(kv, n, "stransfer")) && e->nval) ENUM_FROM_STRING(c->transfer, e->val[0], transfer_t, stransfer_t, Ntransfer_t, "stransfer");
   if ((e = find_kv(kv, n, "sgrowth")) && e->nval) ENUM_FROM_STRING(c->growth, e->val[0], growth_t, sgrowth_t, Ngrowth_t, "sgrowth");
   if ((e = find_kv(kv, n, "sde_param")) && e->nval) ENUM_FROM_STRING(c->de_param, e->val[0], de_param_t, sde_param_t, Nde_param_t, "sde_param");
   if (c->de_param == poly_DE) {
      c->N_poly_de = (int)kv_double(kv, n, "N_poly_de", 0, 1, err);         forwardError(*err, __LINE__, );
      e = find_kv(kv, n, "w_poly_de");
      testErrorRet(!e || e->nval < c->N_poly_de, ce_file, "w_poly_de entries missing", *err, __LINE__, );
      c->w_poly_de = (double *)malloc_err(sizeof(double) * c->N_poly_de, err);
      forwardError(*err, __LINE__, );
      for (int i = 0; i < c->N_poly_de; i++) c->w_poly_de[i] = atof(e->val[i]);
   }
   *self = c;
}

void dump_param(cosmo *c, FILE *F)
{
   if (!F) F = stderr;
   fprintf(F, "#  O_m    O_de   w0_de  w1_de  h_100  O_b    O_nu   Neffnu norm   n_s  nonlin transf growth de_param normmode a_min\n");
   fprintf(F, "# %6.4f % 6.4f % 6.3f % 6.3f %6.4f %6.4f %6.4f %6.3f %6.4f %5.3f %s %s %s %s %d %g\n", c->Omega_m, c->Omega_de,
           c->w0_de, c->w1_de, c->h_100, c->Omega_b, c->Omega_nu_mass, c->Neff_nu_mass, c->normalization, c->n_spec,
           snonlinear_t(c->nonlinear), stransfer_t(c->transfer), sgrowth_t(c->growth), sde_param_t(c->de_param),
           c->normmode, c->a_min);
}

/* conservative dark-energy prior: -1 <= w(a) <= -1/3 at a = 1 and a = a_acc
 * (volume terms at wrappers/src/param.c:1079-1094).  Returns 1 if violated. */
int test_range_de_conservative(cosmo *m, error **err)
{
   (void)err;
   double w_now = m->w0_de, w_acc;
   if (m->de_param == linder) w_acc = m->w0_de + m->w1_de * (1.0 - a_acc);
   else if (m->de_param == jassal) w_acc = m->w0_de + m->w1_de * a_acc * (1.0 - a_acc);
   else w_acc = w_now;
   return (w_now < -1.0 || w_now > -1.0 / 3.0 || w_acc < -1.0 || w_acc > -1.0 / 3.0) ? 1 : 0;
}

/* only used with the coyote10 emulator (param.c:1563-1566), which has no device path */
double getH0fromCMB(double omega_m, double omega_b, double w0_de, int flag)
{
   (void)omega_m; (void)omega_b; (void)w0_de; (void)flag;
   return -1.0;
}

/* ---- SN Ia ------------------------------------------------------------------------------ */
SnSample *SnSample_read(const char *FileName, sndatformat_t fmt, error **err)
{
   (void)fmt;     /* both formats share the row layout name z m dm s ds c dc cov_ms cov_mc cov_sc */
   unsigned int ncomment;
   unsigned int nl = numberoflines_comments(FileName, &ncomment, err);
   forwardError(*err, __LINE__, NULL);
   FILE *F = fopen_err(FileName, "r", err);
   forwardError(*err, __LINE__, NULL);
   SnSample *sn = (SnSample *)calloc_err(1, sizeof(SnSample), err);       forwardError(*err, __LINE__, NULL);
   sn->data = (SnData *)calloc_err(nl + 1, sizeof(SnData), err);            forwardError(*err, __LINE__, NULL);
   char line[4096];
   int n = 0;
   while (fgets(line, sizeof(line), F)) {
      char *s = line;
      while (isspace((unsigned char)*s)) s++;
      if (*s == '#' || *s == 0) continue;
      if (*s == '@') {
         char key[128]; double v;
         if (sscanf(s, "%127s %lg", key, &v) == 2) {
            if (!strcmp(key, "@INTRINSIC_DISPERSION")) sn->int_disp = v;
            else if (!strcmp(key, "@PECULIAR_VELOCITY")) sn->sig_mu_pec_vel = v;
         }
         continue;
      }
      SnData *d = &sn->data[n];
      double cms, cmc, csc;
      int k = sscanf(s, "%31s %lg %lg %lg %lg %lg %lg %lg %lg %lg %lg", d->name, &d->z, &d->musb, &d->dmusb, &d->s,
                     &d->ds, &d->c, &d->dc, &cms, &cmc, &csc);
      if (k != 11) {
         fclose(F);
         *err = addErrorVA(ce_file, "Cannot parse SN data row %d of '%s' (%d of 11 columns)", *err, __LINE__, n, FileName, k);
         return NULL;
      }
      /* columns 4/6/8 are standard deviations: the reader squares them */
      d->cov[0][0] = d->dmusb * d->dmusb; d->cov[1][1] = d->ds * d->ds; d->cov[2][2] = d->dc * d->dc;
      d->cov[0][1] = d->cov[1][0] = cms; d->cov[0][2] = d->cov[2][0] = cmc; d->cov[1][2] = d->cov[2][1] = csc;
      n++;
   }
   fclose(F);
   testErrorRetVA(n == 0, ce_file, "No supernova found in '%s'", *err, __LINE__, NULL, FileName);
   sn->Nsample = n;
   sn->z = (double *)malloc_err(sizeof(double) * 10 * n, err);              forwardError(*err, __LINE__, NULL);
   sn->m = sn->z + n; sn->s = sn->m + n; sn->c = sn->s + n; sn->cov6 = sn->c + n;
   for (int i = 0; i < n; i++) {
      const SnData *d = &sn->data[i];
      sn->z[i] = d->z; sn->m[i] = d->musb; sn->s[i] = d->s; sn->c[i] = d->c;
      double *q = sn->cov6 + 6 * i;
      q[0] = d->cov[0][0]; q[1] = d->cov[1][1]; q[2] = d->cov[2][2]; q[3] = d->cov[0][1]; q[4] = d->cov[0][2]; q[5] = d->cov[1][2];
   }
   return sn;
}

void SnSample_free(SnSample **sn)
{
   if (!sn || !*sn) return;
   free((*sn)->z); free((*sn)->data); free(*sn); *sn = NULL;
}

cosmo_SN *set_cosmological_parameters_to_default_SN(error **err)
{
   cosmo_SN *s = (cosmo_SN *)calloc_err(1, sizeof(cosmo_SN), err);
   forwardError(*err, __LINE__, NULL);
   s->cosmo = set_cosmological_parameters_to_default2(err);
   forwardError(*err, __LINE__, NULL);
   s->Theta2[0] = 19.31; s->Theta2[1] = 1.6; s->Theta2[2] = -1.8; s->Theta2[3] = 0.0;
   s->stretch = 1.0; s->color = 0.0; s->beta_d = 0.0; s->chi2mode = chi2_simple;
   return s;
}

/* par_files/cosmo_SN.par: `cosmo_file NAME` (relative to the working directory) + `Theta2 -M alpha -beta beta_z` */
void read_cosmological_parameters_SN(cosmo_SN **self, FILE *F, error **err)
{
   static kv_t kv[64];
   int n = read_kv(F, kv, 64);
   cosmo_SN *s = (cosmo_SN *)calloc_err(1, sizeof(cosmo_SN), err);
   forwardError(*err, __LINE__, );
   const kv_t *e = find_kv(kv, n, "cosmo_file");
   testErrorRet(!e || e->nval < 1, ce_file, "Key 'cosmo_file' not found in SN parameter file", *err, __LINE__, );
   if (!strcmp(e->val[0], "-")) { s->cosmo = set_cosmological_parameters_to_default2(err); forwardError(*err, __LINE__, ); }
   else {
      FILE *FC = fopen_err(e->val[0], "r", err);                            forwardError(*err, __LINE__, );
      read_cosmological_parameters(&s->cosmo, FC, err);
      fclose(FC);
      forwardError(*err, __LINE__, );
   }
   e = find_kv(kv, n, "Theta2");
   testErrorRet(!e || e->nval < 3, ce_file, "Key 'Theta2' (-M alpha -beta [beta_z]) not found", *err, __LINE__, );
   for (int i = 0; i < NLCP; i++) s->Theta2[i] = i < e->nval ? atof(e->val[i]) : 0.0;
   if ((e = find_kv(kv, n, "Theta1"))) for (int i = 0; i < e->nval && i < NTHETA1; i++) s->Theta1[i] = atof(e->val[i]);
   s->stretch = 1.0; s->color = 0.0; s->chi2mode = chi2_simple;
   *self = s;
}

cosmo_SN *copy_parameters_SN_only(cosmo_SN *src, error **err)
{
   cosmo_SN *s = (cosmo_SN *)malloc_err(sizeof(cosmo_SN), err);
   forwardError(*err, __LINE__, NULL);
   *s = *src;
   s->cosmo = copy_parameters_only(src->cosmo, err);
   forwardError(*err, __LINE__, NULL);
   return s;
}

void free_parameters_SN(cosmo_SN **s)
{
   if (!s || !*s) return;
   free_parameters(&(*s)->cosmo); free(*s); *s = NULL;
}

void updateFrom_SN(cosmo_SN *a, cosmo_SN *b, error **err) { (void)a; (void)b; (void)err; }

void dump_param_SN(cosmo_SN *s, FILE *F)
{
   if (!F) F = stderr;
   dump_param(s->cosmo, F);
   fprintf(F, "# Theta2 = (%g %g %g %g) stretch = %g color = %g chi2mode = %s\n", s->Theta2[0], s->Theta2[1], s->Theta2[2],
           s->Theta2[3], s->stretch, s->color, schi2mode_t(s->chi2mode));
}

double distance_module(cosmo *self, double dlum, error **err)
{
   (void)self;
   testErrorRetVA(!(dlum > 0.0), ce_negative, "Luminosity distance %g not positive", *err, __LINE__, 0.0, dlum);
   return 5.0 * log10(dlum / 0.7) + 25.0;       /* h absorbed in M (manual.tex:1321) */
}

static void cosmo_to_b200(const cosmo *c, pmcb200_cosmo_t *o)
{
   o->Omega_m = c->Omega_m; o->Omega_de = c->Omega_de; o->w0_de = c->w0_de; o->w1_de = c->w1_de;
   o->h_100 = c->h_100; o->Omega_b = c->Omega_b; o->Omega_nu_mass = c->Omega_nu_mass;
   o->Neff_nu_mass = c->Neff_nu_mass; o->de_param = (int)c->de_param; o->_pad = 0;
}

/* The luminosity distances are computed together with the chi^2 on the device
 * (one kernel); SetDl only validates the model. */
void SetDl(cosmo_SN *self, SnSample *sn, error **err)
{
   (void)sn;
   testErrorRet(self->cosmo->de_param != linder && self->cosmo->de_param != jassal, ce_de,
                "Only the jassal / linder dark-energy parametrisations have a device path", *err, __LINE__, );
}

/* log L = -chi^2/2 [- sum log sigma^2 / 2] for ONE model: the batched SN kernel with N = 1.
 * The two "parameters" handed over are the model's stretch and colour zero points. */
double chi2_SN(const cosmo_SN *m, const SnSample *sn, mvdens *data_beta_d, int wTheta1, int add_logdetCov, error **err)
{
   (void)data_beta_d;
   testErrorRet(wTheta1 != 0 || m->chi2mode > chi2_betaz, ce_unknown,
                "chi2_Theta1 / chi2_dust / chi2_residual have no device path", *err, __LINE__, 0.0);
   static pmcb200_like_t L;
   memset(&L, 0, sizeof(L));
   L.kind = PMCB200_LIKE_SNIa; L.npar = 2; L.par[0] = PMCB200_P_stretch; L.par[1] = PMCB200_P_color;
   L.special = PMCB200_SPECIAL_none;
   cosmo_to_b200(m->cosmo, &L.model);
   L.sn_chi2mode = (int)m->chi2mode; L.sn_add_logdetCov = add_logdetCov;
   for (int i = 0; i < 4; i++) L.sn_Theta2[i] = m->Theta2[i];
   for (int i = 0; i < 3; i++) L.sn_Theta2_denom[i] = m->Theta2_denom[i];
   L.sn_sig_int = sn->int_disp; L.sn_v_pec = sn->sig_mu_pec_vel;
   L.sn_n = sn->Nsample; L.sn_z = sn->z; L.sn_m = sn->m; L.sn_s = sn->s; L.sn_c = sn->c; L.sn_cov = sn->cov6;
   double x[2] = {m->stretch, m->color};
   double r = pmc_b200_single_loglike(&L, x, err);
   forwardError(*err, __LINE__, 0.0);
   return r;
}

/* ---- BAO / CMB distance priors ----------------------------------------------------------------- */
static double gauss_like(int kind, int method, cosmo *model, mvdens *g, const double *z, error **err)
{
   static pmcb200_like_t L;
   memset(&L, 0, sizeof(L));
   mvdens_cholesky_decomp(g, err);              /* as mvdens_log_pdf does on first use */
   forwardError(*err, __LINE__, 0.0);
   testErrorRet(model->de_param != linder && model->de_param != jassal, ce_de,
                "Only the jassal / linder dark-energy parametrisations have a device path", *err, __LINE__, 0.0);
   L.kind = kind; L.npar = 1; L.par[0] = 111 /* p_dummy */; L.special = PMCB200_SPECIAL_none;
   cosmo_to_b200(model, &L.model);
   L.bao_method = method; L.g_ndim = (int)g->ndim; L.g_z = z; L.g_mean = g->mean; L.g_chol = g->std;
   double x[1] = {0.0};
   double r = pmc_b200_single_loglike(&L, x, err);
   forwardError(*err, __LINE__, 0.0);
   return r;
}
double chi2_bao_A(cosmo *model, mvdens *g, const double *z, error **err)
{ return gauss_like(PMCB200_LIKE_BAO, PMCB200_BAO_distance_A, model, g, z, err); }
double chi2_bao_d_z(cosmo *model, mvdens *g, const double *z, error **err)
{ return gauss_like(PMCB200_LIKE_BAO, PMCB200_BAO_distance_d_z, model, g, z, err); }
double chi2_bao_D_V_ratio(cosmo *model, mvdens *g, const double *z, error **err)
{ return gauss_like(PMCB200_LIKE_BAO, PMCB200_BAO_distance_D_V_ratio, model, g, z, err); }
double chi2_cmbDP(cosmo *model, mvdens *g, error **err)
{ return gauss_like(PMCB200_LIKE_CMBDistPrior, 0, model, g, NULL, err); }
