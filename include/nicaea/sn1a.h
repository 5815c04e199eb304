/* nicaea/sn1a.h -- SN Ia interface of nicaea as called by wrappers/src/sn.c
 * (:69 SnSample_read, :128-132 model readers, :153 copy, :260 SetDl, :270 chi2_SN,
 * :278 free).  The chi^2 of ONE model is evaluated by the batched GPU kernel with
 * N = 1. */
#ifndef NICAEA_SN1A_H
#define NICAEA_SN1A_H
#include "nicaea/cosmo.h"
#include "pmctools/mvdens.h"
#ifdef __cplusplus
extern "C" {
#endif
#define sn_cosmo_base (-1500)
#define NDER 4
#define NLCP 4
#define NTHETA1 8

/* values 0..3 are PMCB200_CHI2_* */
typedef enum {chi2_simple, chi2_Theta2_denom_fixed, chi2_no_sc, chi2_betaz, chi2_dust, chi2_Theta1, chi2_residual} chi2mode_t;
#define schi2mode_t(i) ( \
  i==chi2_simple ? "chi2_simple" : i==chi2_Theta2_denom_fixed ? "chi2_Theta2_denom_fixed" : \
  i==chi2_no_sc ? "chi2_no_sc" : i==chi2_betaz ? "chi2_betaz" : i==chi2_dust ? "chi2_dust" : \
  i==chi2_Theta1 ? "chi2_Theta1" : i==chi2_residual ? "chi2_residual" : "")
#define Nchi2mode_t 7
typedef enum {SNLS_firstyear, SN_SALT} sndatformat_t;
#define ssndatformat_t(i) (i==SNLS_firstyear ? "SNLS_firstyear" : i==SN_SALT ? "SN_SALT" : "")
#define Nsndatformat_t 2

typedef struct {
  double z, musb, s, c;      /* redshift, rest-frame B magnitude, stretch, colour */
  double dmusb, ds, dc;      /* their errors */
  double cov[3][3];          /* covariance of (m, s, c) */
  double dl, mu_c, dust;
  char name[32];
} SnData;

typedef struct {
  SnData *data;
  int Nsample;
  double int_disp, sig_mu_pec_vel;     /* @INTRINSIC_DISPERSION, @PECULIAR_VELOCITY */
  double logdetW1;
  /* contiguous views for the device path */
  double *z, *m, *s, *c, *cov6;
} SnSample;

typedef struct {
  cosmo *cosmo;
  double Theta1[NTHETA1], Theta2[NLCP], Theta2_denom[NLCP];
  double beta_d, stretch, color;
  chi2mode_t chi2mode;
} cosmo_SN;

SnSample *SnSample_read(const char *FileName, sndatformat_t sndatformat, error **err);
void      SnSample_free(SnSample **sn);
void      read_cosmological_parameters_SN(cosmo_SN **self, FILE *F, error **err);
cosmo_SN *set_cosmological_parameters_to_default_SN(error **err);
cosmo_SN *copy_parameters_SN_only(cosmo_SN *source, error **err);
void      free_parameters_SN(cosmo_SN **self);
void      updateFrom_SN(cosmo_SN *avant, cosmo_SN *apres, error **err);
void      dump_param_SN(cosmo_SN *self, FILE *F);
void      SetDl(cosmo_SN *self, SnSample *sn, error **err);
double    chi2_SN(const cosmo_SN *cosmo, const SnSample *sn, mvdens *data_beta_d, int wTheta1, int add_logdetCov,
                  error **err);
double    distance_module(cosmo *self, double dlum, error **err);
#ifdef __cplusplus
}
#endif
#endif
