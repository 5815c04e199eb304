"""ctypes mirror of include/pmcb200.h (POD structs + the library loader).

Python here is plumbing only: it builds the plain-C descriptors the C-ABI
takes and loads ``libpmc_b200.so``.  There is no Python/torch compute path and
no CPU fallback: if the CUDA library is missing, loading raises.
"""
import ctypes as C
import os

MAX_DIM, MAX_COMP, MAX_DATA = 32, 64, 4

# par_t values, tools/include/par.h:11-34 of the reference (see pmcb200.h)
P = dict(Omegam=0, Omegab=1, Omegade=2, h100=3, Omeganumass=4, Omegac=5, OmegaK=6,
         omegam=7, omegab=8, omegab100=9, omegade=10, omeganumass=11, omegac=12,
         omegaK=13, w0de=14, w1de=15, Neffnumass=20, M=37, alpha=38, beta=39,
         beta_z=40, logbeta=41, stretch=42, color=43, dummy=111)
# spar_t strings of the reference config files -> par_t (tools/include/par.h:36-149)
SPAR = {"Omega_m": P["Omegam"], "Omega_b": P["Omegab"], "Omega_de": P["Omegade"],
        "h_100": P["h100"], "Omega_nu_mass": P["Omeganumass"], "Omega_c": P["Omegac"],
        "Omega_K": P["OmegaK"], "omega_m": P["omegam"], "omega_b": P["omegab"],
        "100_omega_b": P["omegab100"], "omega_de": P["omegade"], "omega_c": P["omegac"],
        "omega_K": P["omegaK"], "w_0_de": P["w0de"], "w_1_de": P["w1de"],
        "M": P["M"], "alpha": P["alpha"], "beta": P["beta"], "logbeta": P["logbeta"],
        "beta_z": P["beta_z"],
        # the reference's own spellings (tools/include/par.h:36-149)
        "log_beta": P["logbeta"], "omega_nu_mass": P["omeganumass"], "N_eff_nu_mass": P["Neffnumass"],
        "stretch": P["stretch"], "color": P["color"]}

LIKE = dict(Mvdens=0, MixMvdens=1, SNIa=3, CMBDistPrior=6, BAO=7, BANANA=100)
SPECIAL = dict(none=0, unity=1, de_conservative=2)
CHI2 = dict(chi2_simple=0, chi2_Theta2_denom_fixed=1, chi2_no_sc=2, chi2_betaz=3)
BAO_METHOD = dict(distance_A=0, distance_d_z=1, distance_D_V_ratio=2)
DE = dict(jassal=0, linder=1)

ERR = dict(CUDA=-9001, ARG=-9002, DIM=-9003, CHOLESKY=-9004, NOSAMPLE=-9005,
           UNSUP=-9006, STATE=-9007)

dptr = C.POINTER(C.c_double)


class Cosmo(C.Structure):
    _fields_ = [("Omega_m", C.c_double), ("Omega_de", C.c_double), ("w0_de", C.c_double),
                ("w1_de", C.c_double), ("h_100", C.c_double), ("Omega_b", C.c_double),
                ("Omega_nu_mass", C.c_double), ("Neff_nu_mass", C.c_double),
                ("de_param", C.c_int), ("_pad", C.c_int)]


class Like(C.Structure):
    _fields_ = [("kind", C.c_int), ("npar", C.c_int), ("par", C.c_int * MAX_DIM),
                ("special", C.c_int), ("model", Cosmo),
                ("sn_chi2mode", C.c_int), ("sn_add_logdetCov", C.c_int),
                ("sn_Theta2", C.c_double * 4), ("sn_Theta2_denom", C.c_double * 3),
                ("sn_sig_int", C.c_double), ("sn_v_pec", C.c_double),
                ("sn_n", C.c_int),
                ("sn_z", dptr), ("sn_m", dptr), ("sn_s", dptr), ("sn_c", dptr),
                ("sn_cov", dptr),
                ("bao_method", C.c_int), ("g_ndim", C.c_int),
                ("g_z", dptr), ("g_mean", dptr), ("g_chol", dptr),
                ("mix_ncomp", C.c_int), ("mix_ndim", C.c_int), ("mix_df", C.c_int),
                ("mix_wght", dptr), ("mix_mean", dptr), ("mix_chol", dptr),
                ("banana_b", C.c_double), ("banana_sigma1sq", C.c_double)]


class Target(C.Structure):
    _fields_ = [("npar", C.c_int), ("ndata", C.c_int),
                ("min", C.c_double * MAX_DIM), ("max", C.c_double * MAX_DIM),
                ("like", Like * MAX_DATA),
                ("nprior", C.c_int), ("indprior", C.c_int * MAX_DIM),
                ("prior_ndim", C.c_int),
                ("prior_mean", dptr), ("prior_chol", dptr)]


class Stats(C.Structure):
    _fields_ = [("nsamples", C.c_int64), ("nok_box", C.c_int64), ("nok", C.c_int64),
                ("maxW", C.c_double), ("logSum", C.c_double), ("sum_shift", C.c_double),
                ("perplexity", C.c_double), ("ess", C.c_double),
                ("ln_evidence", C.c_double), ("enc", C.c_double),
                ("ndead", C.c_int32), ("_pad", C.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "_pad"}


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# PMCB200_LIB selects another build of the same library (A/B kernel experiments)
LIB_PATH = os.environ.get("PMCB200_LIB") or os.path.join(ROOT, "cosmopmc_b200", "libpmc_b200.so")

# every symbol include/pmcb200.h declares: name -> (restype, argtypes)
_vp, _i, _i64, _u64, _u32, _d = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_uint32, C.c_double
SYMBOLS = {
    "pmcb200_create": (_i, [_i, _vp, C.POINTER(_vp)]),
    "pmcb200_destroy": (None, [_vp]),
    "pmcb200_last_error": (C.c_char_p, [_vp]),
    "pmcb200_version": (_i, []),
    "pmcb200_device_count": (_i, []),
    "pmcb200_sync": (_i, [_vp]),
    "pmcb200_stream": (_vp, [_vp]),
    "pmcb200_set_proposal": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp]),
    "pmcb200_set_proposal_cov": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp]),
    "pmcb200_get_proposal": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "pmcb200_set_target": (_i, [_vp, C.POINTER(Target)]),
    "pmcb200_simulate": (_i, [_vp, _i64, _u64, _u32, _i64, _vp, _vp, _vp]),
    "pmcb200_simulate_from_draws": (_i, [_vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    "pmcb200_proposal_log_pdf": (_i, [_vp, _i64, _vp, _vp]),
    "pmcb200_posterior_log_pdf": (_i, [_vp, _i64, _vp, _vp, _vp]),
    "pmcb200_map_params": (_i, [_vp, _i, _i64, _vp, _vp, _vp]),
    "pmcb200_importance_weights": (_i, [_vp, _i64, _vp, _d, _vp, _vp]),
    "pmcb200_normalize_weights": (_i, [_vp, _i64, _vp, _vp]),
    "pmcb200_stat_block_len": (_i64, [_vp]),
    "pmcb200_em_local": (_i, [_vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    "pmcb200_em_finish": (_i, [_vp, _i, _vp, _i64, C.POINTER(Stats)]),
    "pmcb200_iteration_local": (_i, [_vp, _i64, _u64, _u32, _i64, _d, _vp, _vp, _vp, _vp, _vp]),
    "pmcb200_iteration_shard_host": (_i, [_vp, _i64, _u64, _u32, _i64, _d, _vp, _vp, _vp, _vp]),
    "pmcb200_shard_weights_host": (_i, [_vp, _i64, _vp]),
    "pmcb200_iteration_host": (_i, [_vp, _i64, _u64, _u32, _d, _vp, _vp, _vp, _vp, C.POINTER(Stats)]),
    "pmcb200_iteration_host_begin": (_i, [_vp, _i64, _u64, _u32, _d, _vp, _vp, _vp, _vp, C.POINTER(Stats)]),
    "pmcb200_shard_weights_host_begin": (_i, [_vp, _i64, _vp]),
    "pmcb200_host_wait": (_i, [_vp, _i]),
    "pmcb200_samples_host_begin": (_i, [_vp, _i64, _vp, _vp]),
    "pmcb200_launch_count": (_i64, [_vp]),
    "pmcb200_sn_tile_plan": (_i, [_i, _vp, _i, _vp, _vp]),
    "pmcb200_cmb_spectral_tables": (_i, [_vp, _vp, _vp]),
    "pmcb200_set_box": (_i, [_vp, _i, _vp, _vp]),
    "pmcb200_read_counts": (_i, [_vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_double)]),
    "pmcb200_weight_stats": (_i, [_vp, _i64, _vp, _vp, _i, C.POINTER(C.c_double * 8)]),
    "pmcb200_normalize_log_weights": (_i, [_vp, _i64, _vp, _vp, C.POINTER(_d), C.POINTER(_d), C.POINTER(_d)]),
    "pmcb200_em_local_linear": (_i, [_vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    "pmcb200_counters": (_i, [_vp, C.POINTER(C.c_int64 * 4)]),
    "pmcb200_counters_ex": (_i, [_vp, C.POINTER(C.c_int64), _i]),
    "pmcb200_fp64_peak": (_i, [_vp, C.POINTER(C.c_double)]),
    "pmcb200_dev_alloc": (_i, [_vp, C.c_size_t, C.POINTER(_vp)]),
    "pmcb200_dev_free": (_i, [_vp, _vp]),
    "pmcb200_h2d": (_i, [_vp, _vp, _vp, C.c_size_t]),
    "pmcb200_d2h": (_i, [_vp, _vp, _vp, C.c_size_t]),
    "pmcb200_h2d_async": (_i, [_vp, _vp, _vp, C.c_size_t]),
    "pmcb200_d2h_async": (_i, [_vp, _vp, _vp, C.c_size_t]),
    "pmcb200_host_alloc": (_i, [C.c_size_t, C.POINTER(_vp)]),
    "pmcb200_host_free": (_i, [_vp]),
    "pmcb200_allgather_blocks": (_i, [C.POINTER(_vp), _i, C.POINTER(_vp), C.POINTER(_vp), _i64]),
    "pmcb200_iteration_host_multi": (_i, [C.POINTER(_vp), _i, _i64, _u64, _u32, _d, _vp, _vp, _vp, _vp,
                                          C.POINTER(Stats)]),
    "pmcb200_normalize_with": (_i, [_vp, _i64, _vp, _vp, _d, _d]),
    "pmcb200_post_moments": (_i, [_vp, _i64, _i, _vp, _vp, _vp, _vp, _vp]),
    "pmcb200_post_sigma": (_i, [_vp, _i64, _i, _vp, _vp, _vp, _i, _d, _vp, _vp, C.POINTER(_d), C.POINTER(_i64)]),
    "pmcb200_post_histogram": (_i, [_vp, _i64, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pmcb200_fisher_host": (_i, [_vp, _vp, _vp, _i, _vp, C.POINTER(_i)]),
}

_lib = None


def load_library(path=None):
    """Load libpmc_b200.so and bind every declared symbol.  Raises if absent."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(
            "cosmopmc_b200: CUDA library %s not built (run `python -c 'import "
            "__graft_entry__ as g; g.build()'`). There is no CPU fallback." % p)
    lib = C.CDLL(p)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)      # AttributeError if the symbol is missing
        fn.restype, fn.argtypes = res, args
    if path is None:
        _lib = lib
    return lib
