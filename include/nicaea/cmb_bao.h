/* nicaea/cmb_bao.h -- BAO / CMB distance-prior interface as called by
 * wrappers/src/bao.c:163-171 and wrappers/src/wmap.c:1034 (device, N = 1). */
#ifndef NICAEA_CMB_BAO_H
#define NICAEA_CMB_BAO_H
#include "nicaea/cosmo.h"
#include "pmctools/mvdens.h"
#ifdef __cplusplus
extern "C" {
#endif
double chi2_bao_A(cosmo *model, mvdens *g, const double *z_BAO, error **err);
double chi2_bao_d_z(cosmo *model, mvdens *g, const double *z_BAO, error **err);
double chi2_bao_D_V_ratio(cosmo *model, mvdens *g, const double *z_BAO, error **err);
double chi2_cmbDP(cosmo *model, mvdens *g, error **err);
#ifdef __cplusplus
}
#endif
#endif
