// cosmo_types.cuh -- device-side likelihood descriptor shared by the host API
// (which fills it) and the likelihood kernels.
#pragma once
#include "common.cuh"

#define SN_NODES 64      // tabulated Romberg nodes per redshift: stages 1..7
#define SN_ROW 12        // doubles per supernova row in the device table

// Device-side likelihood descriptor (pointers are device pointers).
struct DevLike {
  int kind, npar, special, pad0;
  int par[PMCB200_MAX_DIM];
  pmcb200_cosmo_t model;
  // SN Ia
  int sn_chi2mode, sn_add_logdetCov, sn_n, sn_nz;
  double Theta2[4], Theta2_denom[3], sig_int2, pv_fac;
  const double2 *nodes;   // [sn_nz][SN_NODES] {ln a, a^-1/2}; entry 0 = a(z), entry i in
                          // [2^(j-2), 2^(j-1)) = nodes of trapezoid stage j >= 2
  const double *nodes_a;  // [sn_nz][SN_NODES] a itself
  const double *nodes4;   // [sn_nz][16][4] {ln a, a^-1/2, a, -}: stages 1..5, one 32-byte load per node (curved / w1)
  const int *first;       // [sn_nz+1] ranges of supernovae sharing a redshift
  const double *sn;       // [sn_n][SN_ROW]: m s | c z | Vmm+pv2+int2 Vss | Vcc Cms | Cmc Csc | - -
  int sn_hasq, sn_flat;   // launch-uniform specialisation flags (set by the host)
  // spectral form (sn_spectral.cuh): Chebyshev points of [a(z_max), 1], the Romberg functional on T_m, its error bound
  const double *cheb_nodes4;  // [SNS_M][4] {ln a_j, 1, a_j, -}
  const double *cheb_W;       // [sn_nz][SNS_M]
  const double *cheb_dmax;    // [SNS_M] max_z |D[z][m]| / h_z
  const double *cheb_Wf;      // [sn_ntile][SNS_M/4 + 3][32] B fragments of the tensor-core kernel (8 supernovae per tile)
  const int *sn_tile_sec;     // [sn_ntile] 0 = primary tile (8 distinct redshifts), 1 = secondary (further supernovae at the
                              // redshifts of the primary tile before it, same columns)
  const uint32_t *cheb_Wt;    // [sn_ntile][32][8] TF32 hi / lo B fragments of the coefficient tail m >= 16 (m16n8k8)
  int sn_ntile, pad1;
  // Gaussian data (BAO, CMB distance priors): packed like a mixture component
  int bao_method, g_ndim;
  const double *g_z;
  const double *g_comp;   // [2 + 2n + tri(n)] = wght, lognorm, mean, L packed, 1/diag
  // analytic targets
  MixHdr mixh;
  const double *mix;
  double banana_b, banana_sigma1sq;
};

