/* nicaea/lensing.h -- OPAQUE type stubs only.  The weak-lensing probe is out of
 * scope; these declarations exist so that the reference's
 * all_wrappers.h / lens.h / param.h parse unchanged. */
#ifndef NICAEA_LENSING_H
#define NICAEA_LENSING_H
#include "nicaea/cosmo.h"
#include "nicaea/nofz.h"
typedef struct cosmo_lens_stub { cosmo *cosmo; redshift_t *redshift; } cosmo_lens;
typedef struct datcov_stub { int Ntheta, Nzbin, Nzcorr, n; double *data, *theta, *cov[3]; } datcov;
typedef int lensdata_t;
typedef int decomp_eb_filter_t;
typedef int lensformat_t;
typedef int cov_scaling_t;
typedef int order_t;
typedef struct cosebi_info_stub { int n; } cosebi_info_t;
#endif
