"""ctypes loader for oracle/_ref/libref_param.so: the reference's own wrappers/src/{param,sn,bao,wmap}.c
compiled unchanged with a recording stand-in for nicaea (oracle/build_ref_param.py, ref_param_record.c).

TEST INFRASTRUCTURE ONLY.  The library is built in the container (needs /root/reference) and travels to the
GPU box as a prebuilt file; `available()` says whether it is there.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libref_param.so")

FN = {1: "SetDl", 2: "chi2_SN", 3: "chi2_bao_A", 4: "chi2_bao_d_z", 5: "chi2_bao_D_V_ratio", 6: "chi2_cmbDP",
      7: "test_range_de_conservative"}
_lib = None


def available():
    return os.path.exists(LIB)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB)
        vp, i, d = C.c_void_p, C.c_int, C.c_double
        pi = C.POINTER(C.c_int)
        L.refp_set_default_model.argtypes = [vp, vp]
        L.refp_set_returns.argtypes = [d, d, d, i, i]
        L.refp_open.restype, L.refp_open.argtypes = i, [C.c_char_p, C.c_char_p, i]
        L.refp_close.argtypes = [i]
        L.refp_logpr_default.restype, L.refp_logpr_default.argtypes = d, [i]
        L.refp_posterior.restype, L.refp_posterior.argtypes = d, [i, vp, pi, vp, i, pi]
        L.refp_likeli.restype, L.refp_likeli.argtypes = d, [i, i, vp, pi, vp, i, pi]
        L.refp_prior_special.restype, L.refp_prior_special.argtypes = d, [i, vp, vp, vp, i, pi]
        L.refp_record_len.restype = i
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def set_default_model(cosmo, Theta2=(0.0, 0.0, 0.0, 0.0)):
    """cosmo: dict with the pmcb200_cosmo_t field names (targets.COSMO_SN / COSMO_DP)"""
    m = np.array([cosmo["Omega_m"], cosmo["Omega_de"], cosmo["w0_de"], cosmo["w1_de"], cosmo["h_100"], cosmo["Omega_b"],
                  cosmo["Omega_nu_mass"], cosmo["Neff_nu_mass"], float(cosmo["de_param"])])
    t = np.array(Theta2, dtype=np.float64)
    lib().refp_set_default_model(_p(m), _p(t))


def set_returns(sn=0.0, bao=0.0, cmb=0.0, de_prior=0, fail_setdl=0):
    lib().refp_set_returns(sn, bao, cmb, de_prior, fail_setdl)


def write_config(path, spar, pmin, pmax, data, sprior="-", nprior=None, indprior=None):
    """The base part of a CosmoPMC config file (Manual/manual.tex, read by read_config_base param.c:73-200).
    data: list of (sdata, [(key, value), ...]) in the order the plug-in's func_read expects."""
    f = ["version\t1.3", "npar\t%d" % len(spar), "n_ded\t0", "spar\t" + " ".join(spar),
         "min\t" + " ".join(repr(float(v)) for v in pmin), "max\t" + " ".join(repr(float(v)) for v in pmax),
         "ndata\t%d" % len(data)]
    f += ["sdata\t%s" % sd for sd, _ in data]
    for _, kv in data:
        f += ["%s\t%s" % (k, v) for k, v in kv]
    f.append("sprior\t%s" % sprior)
    if sprior != "-":
        f += ["nprior\t%d" % nprior, "indprior\t" + " ".join(str(int(v)) for v in indprior)]
    with open(path, "w") as fo:
        fo.write("\n".join(f) + "\n")


def sn_section(chi2mode="chi2_simple", add_logdetCov=0, special="none", Theta2_denom=None):
    kv = [("datname", "unused.list"), ("sdatformat", "SN_SALT"), ("schi2mode", chi2mode)]
    if chi2mode == "chi2_Theta2_denom_fixed":
        kv.append(("Theta2_denom", "%r %r" % (float(Theta2_denom[0]), float(Theta2_denom[1]))))
    kv += [("add_logdetCov", str(add_logdetCov)), ("model_file", "-"), ("sspecial", special)]
    return ("SNIa", kv)


def gauss_file(path, mean, covinv, z=None):
    """mvdens text format (manual.tex:3204-3255) + the `z` key of the BAO files (data/BAO/*)"""
    n = len(mean)
    with open(path, "w") as f:
        f.write("%d -1 %d 0\n" % (n, n) + " ".join(repr(float(v)) for v in mean) + "\n")
        for r in covinv:
            f.write(" ".join(repr(float(v)) for v in r) + "\n")
        if z is not None:
            f.write("z\t" + " ".join(repr(float(v)) for v in z) + "\n")


def bao_section(datname, method="distance_d_z", special="none"):
    return ("BAO", [("smethod", method), ("datname", datname), ("model_file", "-"), ("sspecial", special)])


def cmbdp_section(datname, special="none"):
    return ("CMBDistPrior", [("datname", datname), ("model_file", "-"), ("sspecial", special)])


class RefConfig:
    def __init__(self, path):
        msg = C.create_string_buffer(4096)
        self.h = lib().refp_open(path.encode(), msg, 4096)
        if self.h < 0:
            raise RuntimeError("read_config_base failed (%d): %s" % (self.h, msg.value.decode(errors="replace")))
        self.rl = lib().refp_record_len()

    def close(self):
        lib().refp_close(self.h)

    @property
    def logpr_default(self):
        return lib().refp_logpr_default(self.h)

    def _call(self, fn, *lead):
        rec = np.zeros((32, self.rl))
        e, n = C.c_int(0), C.c_int(0)
        r = fn(*lead, C.byref(e), _p(rec), 32, C.byref(n))
        return r, e.value, rec[:n.value]

    def posterior(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        return self._call(lambda *a: lib().refp_posterior(self.h, _p(x), *a))

    def likeli(self, idata, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        return self._call(lambda *a: lib().refp_likeli(self.h, idata, _p(x), *a))


def prior_special(special, par, pmin, pmax):
    par = np.ascontiguousarray(par, dtype=np.int32)
    pmin, pmax = np.ascontiguousarray(pmin, dtype=np.float64), np.ascontiguousarray(pmax, dtype=np.float64)
    e = C.c_int(0)
    r = lib().refp_prior_special(special, _p(par), _p(pmin), _p(pmax), len(par), C.byref(e))
    return r, e.value


def record_model(rec):
    """the 15 numbers orc_map_params / pmcb200_map_params return, from a record of a nicaea entry point"""
    return np.concatenate([rec[1:10], rec[10:16]])
