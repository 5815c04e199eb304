/* nicaea/cosmo.h -- the subset of nicaea's cosmology interface that the
 * reference's wrappers call (wrappers/src/sn.c:69,128-132,153;
 * bao.c:70-74,92,151; wmap.c:935-939,957,1024; param.c:1544-1661).
 * nicaea is external, un-vendored and un-pinned (install_CosmoPMC.sh:261); the
 * struct carries the fields of the .par files (par_files/cosmo.par) that the
 * wrappers read or write.  Distances are computed on the GPU (pmcb200.h); the
 * scalar functions here evaluate ONE model by launching the same kernels with
 * N = 1 (no CPU implementation of the integrals). */
#ifndef NICAEA_COSMO_H
#define NICAEA_COSMO_H
#include <stdio.h>
#include "pmctools/errorlist.h"
#include "pmctools/maths.h"
#include "pmctools/io.h"
#ifdef __cplusplus
extern "C" {
#endif

#define ce_base     (-1400)
#define ce_alloc    (-1 + ce_base)
#define ce_file     (-2 + ce_base)
#define ce_unknown  (-3 + ce_base)
#define ce_negative (-4 + ce_base)
#define ce_infnan   (-5 + ce_base)
#define ce_de       (-6 + ce_base)
#define ce_range    (-7 + ce_base)
#define ce_noknown  (-8 + ce_base)

#define R_HUBBLE 2997.92458         /* c/(100 km/s/Mpc) [Mpc/h] */

typedef enum {linear, pd96, smith03, smith03_de, coyote10, coyote13, smith03_revised} nonlinear_t;
#define snonlinear_t(i) ( \
  i==linear ? "linear" : i==pd96 ? "pd96" : i==smith03 ? "smith03" : i==smith03_de ? "smith03_de" : \
  i==coyote10 ? "coyote10" : i==coyote13 ? "coyote13" : i==smith03_revised ? "smith03_revised" : "")
#define Nnonlinear_t 7
typedef enum {bbks, eisenhu, eisenhu_osc, be84} transfer_t;
#define stransfer_t(i) (i==bbks ? "bbks" : i==eisenhu ? "eisenhu" : i==eisenhu_osc ? "eisenhu_osc" : i==be84 ? "be84" : "")
#define Ntransfer_t 4
typedef enum {heath, growth_de, camb_vinschter_gr} growth_t;
#define sgrowth_t(i) (i==heath ? "heath" : i==growth_de ? "growth_de" : i==camb_vinschter_gr ? "camb_vinschter_gr" : "")
#define Ngrowth_t 3
/* values 0,1 are PMCB200_DE_jassal / PMCB200_DE_linder */
typedef enum {jassal, linder, earlyDE, poly_DE} de_param_t;
#define sde_param_t(i) (i==jassal ? "jassal" : i==linder ? "linder" : i==earlyDE ? "earlyDE" : i==poly_DE ? "poly_DE" : "")
#define Nde_param_t 4
typedef enum {norm_s8, norm_as} norm_t;

typedef struct {
  double Omega_m, Omega_de, w0_de, w1_de;
  double *w_poly_de;
  int N_poly_de;
  double h_100, Omega_b, Omega_nu_mass, Neff_nu_mass;
  double normalization, sigma_8, As, n_spec;
  nonlinear_t nonlinear;
  transfer_t transfer;
  growth_t growth;
  de_param_t de_param;
  int normmode;
  double a_min;
  /* nicaea keeps interpolation tables here; the device path has none */
  void *tables;
} cosmo;

cosmo *init_parameters(double OMEGAM, double OMEGADE, double W0_DE, double W1_DE, double *W_POLY_DE, int N_POLY_DE,
                       double H100, double OMEGAB, double OMEGANUMASS, double NEFFNUMASS, double NORM, double NSPEC,
                       nonlinear_t NONLINEAR, transfer_t TRANSFER, growth_t GROWTH, de_param_t DEPARAM,
                       norm_t normmode, double AMIN, error **err);
cosmo *copy_parameters_only(cosmo *source, error **err);
cosmo *copy_parameters(cosmo *source, error **err);
void   read_cosmological_parameters(cosmo **self, FILE *F, error **err);
cosmo *set_cosmological_parameters_to_default(error **err);
cosmo *set_cosmological_parameters_to_default2(error **err);
void   free_parameters(cosmo **self);
void   updateFrom(cosmo *avant, cosmo *apres, error **err);
void   dump_param(cosmo *self, FILE *F);
int    test_range_de_conservative(cosmo *model, error **err);
double getH0fromCMB(double omega_m, double omega_b, double w0_de, int flag);

#ifdef __cplusplus
}
#endif
#endif
