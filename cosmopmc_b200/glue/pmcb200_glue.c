/* pmcb200_glue.c -- CosmoPMC-side binding.
 *
 * Compiled WITH THE REFERENCE'S OWN HEADERS (param.h, sn.h, bao.h, wmap.h) next
 * to the unchanged exec/ and wrappers/ sources.  It overrides the library's weak
 * pmc_b200_autobind hook: when generic_get_importance_weight_and_deduced_verb is
 * called with (posterior_log_pdf_common_void, &config->base)
 * (exec/cosmo_pmc.c:343-345) it flattens `config_base` and the plug-in states
 * into the plain-C device target, so the unchanged driver runs on the GPU. */
#include <string.h>
#include "param.h"
#include "sn.h"
#include "bao.h"
#include "wmap.h"
#include "pmclib/pmc.h"

static void cosmo_to_b200(const cosmo *c, pmcb200_cosmo_t *o)
{
   o->Omega_m = c->Omega_m; o->Omega_de = c->Omega_de; o->w0_de = c->w0_de; o->w1_de = c->w1_de;
   o->h_100 = c->h_100; o->Omega_b = c->Omega_b; o->Omega_nu_mass = c->Omega_nu_mass;
   o->Neff_nu_mass = c->Neff_nu_mass; o->de_param = (int)c->de_param;
}

/* mvdens holding a covariance (after mvdens_inverse at init) -> mean + lower Cholesky factor */
static void gauss_to_b200(mvdens *g, pmcb200_like_t *L, error **err)
{
   mvdens_cholesky_decomp(g, err);
   forwardError(*err, __LINE__, );
   L->g_ndim = (int)g->ndim; L->g_mean = g->mean; L->g_chol = g->std;
}

static void mix_to_b200(void *state, int is_mix, pmcb200_like_t *L, error **err)
{
   /* likeli_Mvdens / likeli_MixMvdens states are Cholesky-decomposed at read time
    * (param.c:1476, 1524); the component buffers are contiguous per component, so
    * gather means and factors into two arrays that live as long as the config */
   size_t K, d;
   mvdens **comp, *one;
   double *wght;
   if (is_mix) { mix_mvdens *m = (mix_mvdens *)state; K = m->ncomp; d = m->ndim; comp = m->comp; wght = m->wght; }
   else { one = (mvdens *)state; K = 1; d = one->ndim; comp = &one; wght = NULL; }
   double *mean = (double *)malloc_err(sizeof(double) * K * d * (d + 1), err);
   forwardError(*err, __LINE__, );
   double *chol = mean + K * d;
   for (size_t k = 0; k < K; k++) {
      mvdens_cholesky_decomp(comp[k], err);
      forwardError(*err, __LINE__, );
      memcpy(mean + k * d, comp[k]->mean, d * sizeof(double));
      memcpy(chol + k * d * d, comp[k]->std, d * d * sizeof(double));
   }
   L->mix_ncomp = (int)K; L->mix_ndim = (int)d; L->mix_df = comp[0]->df;
   L->mix_wght = wght; L->mix_mean = mean; L->mix_chol = chol;
}

int pmc_b200_autobind(posterior_log_pdf_func *f, void *target_data, pmcb200_target_t *t, error **err)
{
   if (f != posterior_log_pdf_common_void) return 0;
   config_base *cfg = (config_base *)target_data;
   int i, j;
   memset(t, 0, sizeof(*t));
   testErrorRetVA(cfg->npar > PMCB200_MAX_DIM || cfg->ndata > PMCB200_MAX_DATA, mk_npar,
                  "npar = %d / ndata = %d exceed the device limits", *err, __LINE__, 0, cfg->npar, cfg->ndata);
   testErrorRet(cfg->n_ded > 0, mk_undef, "Deduced parameters have no device path", *err, __LINE__, 0);
   t->npar = cfg->npar; t->ndata = cfg->ndata;
   for (j = 0; j < cfg->npar; j++) { t->min[j] = cfg->min[j]; t->max[j] = cfg->max[j]; }
   for (i = 0; i < cfg->ndata; i++) {
      common_like *like = (common_like *)cfg->data_extra[i];
      pmcb200_like_t *L = &t->like[i];
      L->kind = (int)cfg->data[i];                                      /* data_t == PMCB200_LIKE_* */
      L->npar = like->npar;
      for (j = 0; j < like->npar; j++) L->par[j] = (int)like->par[j];   /* par_t == PMCB200_P_*     */
      switch (cfg->data[i]) {
         case SNIa: {
            Sn_state *s = (Sn_state *)like->state;
            SnSample *sn = s->sample;
            L->special = (int)s->special;
            cosmo_to_b200(s->model->cosmo, &L->model);
            L->sn_chi2mode = (int)s->chi2mode; L->sn_add_logdetCov = s->add_logdetCov;
            for (j = 0; j < 4; j++) L->sn_Theta2[j] = s->model->Theta2[j];
            for (j = 0; j < 3; j++) L->sn_Theta2_denom[j] = s->Theta2_denom[j];
            L->sn_sig_int = sn->int_disp; L->sn_v_pec = sn->sig_mu_pec_vel;
            L->sn_n = sn->Nsample;
            L->sn_z = sn->z; L->sn_m = sn->m; L->sn_s = sn->s; L->sn_c = sn->c; L->sn_cov = sn->cov6;
            break; }
         case BAO: {
            bao_state *b = (bao_state *)like->state;
            L->special = (int)b->special;
            cosmo_to_b200(b->model, &L->model);
            L->bao_method = (int)b->method; L->g_z = b->z;
            gauss_to_b200(b->data, L, err);
            forwardError(*err, __LINE__, 0);
            break; }
         case CMBDistPrior: {
            cmbDP_state *c = (cmbDP_state *)like->state;
            L->special = (int)c->special;
            cosmo_to_b200(c->model, &L->model);
            gauss_to_b200(c->data, L, err);
            forwardError(*err, __LINE__, 0);
            break; }
         case Mvdens: case MixMvdens:
            mix_to_b200(like->state, cfg->data[i] == MixMvdens, L, err);
            forwardError(*err, __LINE__, 0);
            break;
         default:
            *err = addErrorVA(mk_undef, "Data type %s has no device likelihood", *err, __LINE__, sdata_t(cfg->data[i]));
            return 0;
      }
   }
   if (cfg->prior != NULL) {                                            /* param.c:1009-1026 */
      mvdens_cholesky_decomp(cfg->prior, err);
      forwardError(*err, __LINE__, 0);
      t->prior_ndim = (int)cfg->prior->ndim; t->prior_mean = cfg->prior->mean; t->prior_chol = cfg->prior->std;
      t->nprior = cfg->nprior;
      for (j = 0; j < cfg->npar; j++) t->indprior[j] = cfg->nprior ? cfg->indprior[j] : 0;
   }
   return 1;
}
