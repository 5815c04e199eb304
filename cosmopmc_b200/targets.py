"""Builders for the plain-C posterior descriptor (pmcb200_target_t).

These mirror what the reference builds from a `config_pmc` file:
`read_config_base` + each plug-in's `func_read`/`func_init`
(wrappers/src/param.c:73-200, sn.c:20-135, bao.c:23-77, wmap.c:896-942).
Host-side set-up only (numpy for the tiny data matrices); no sample-path
compute happens here.
"""
import ctypes as C
import os

import numpy as np

from . import _abi as A

ROOT = A.ROOT
SN_FIXTURE = os.path.join(ROOT, "tests", "golden", "sn_union307.txt")

# par_files/cosmo.par:3-13,44 (h=0.73) and par_files/cosmoDP.par (h=0.71, Ob=0.045)
COSMO_SN = dict(Omega_m=0.27, Omega_de=0.73, w0_de=-1.0, w1_de=0.0, h_100=0.73,
                Omega_b=0.049, Omega_nu_mass=0.0, Neff_nu_mass=0.0, de_param=A.DE["linder"])
COSMO_DP = dict(Omega_m=0.27, Omega_de=0.73, w0_de=-1.0, w1_de=0.0, h_100=0.71,
                Omega_b=0.045, Omega_nu_mass=0.0, Neff_nu_mass=0.0, de_param=A.DE["linder"])

# data/BAO/bao_BOSS12_d_z_0.57, data/BAO/bao_Reid10_A_0.35 (mean, inverse variance, z)
BAO_BOSS12_DZ = dict(method="distance_d_z", mean=[0.07315], covinv=[[721850.0]], z=[0.57])
BAO_REID10_A = dict(method="distance_A", mean=[0.493], covinv=[[3460.21]], z=[0.35])
# data/WMAP_Distance_Priors/wmap7DistPrior_ML_covinv (l_A, R, z_star; inverse covariance)
WMAP7_DP = dict(mean=[302.09, 1.725, 1091.3],
                covinv=[[2.305, 29.698, -1.333],
                        [29.698, 6825.270, -113.180],
                        [-1.333, -113.180, 3.414]])


def _arr(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64))


def _ptr(a):
    return a.ctypes.data_as(A.dptr)


class TargetSpec:
    """Owns a pmcb200_target_t plus the numpy arrays its pointers reference."""

    def __init__(self, spar, pmin, pmax):
        self.t = A.Target()
        self.keep = []
        self.spar = list(spar)
        d = len(spar)
        assert d <= A.MAX_DIM
        self.t.npar = d
        self.t.ndata = 0
        for j in range(d):
            self.t.min[j] = pmin[j]
            self.t.max[j] = pmax[j]
        self.t.nprior = 0
        self.t.prior_ndim = 0

    @property
    def ndim(self):
        return self.t.npar

    @property
    def box(self):
        d = self.t.npar
        return (np.array([self.t.min[j] for j in range(d)]),
                np.array([self.t.max[j] for j in range(d)]))

    def _new_like(self, kind, special="none", cosmo=None):
        i = self.t.ndata
        assert i < A.MAX_DATA
        L = self.t.like[i]
        L.kind = A.LIKE[kind]
        L.npar = self.t.npar
        for j, s in enumerate(self.spar):
            L.par[j] = A.SPAR.get(s, A.P["dummy"])
        L.special = A.SPECIAL[special]
        if cosmo:
            for k, v in cosmo.items():
                setattr(L.model, k, v)
        self.t.ndata = i + 1
        return L

    def _hold(self, x):
        a = _arr(x)
        self.keep.append(a)
        return _ptr(a)

    # -- SN Ia (sn.c:20-135) -------------------------------------------------
    def add_snia(self, table=None, chi2mode="chi2_simple", add_logdetCov=0,
                 Theta2=(19.31, 1.6, -1.8, 0.0), Theta2_denom=(0.0, 0.0, 0.0),
                 cosmo=None, special="none"):
        tab, sig_int, v_pec = load_sn_table(table or SN_FIXTURE)
        L = self._new_like("SNIa", special, cosmo or COSMO_SN)
        L.sn_chi2mode = A.CHI2[chi2mode]
        L.sn_add_logdetCov = add_logdetCov
        for i in range(4):
            L.sn_Theta2[i] = Theta2[i]
        for i in range(3):
            L.sn_Theta2_denom[i] = Theta2_denom[i]
        L.sn_sig_int, L.sn_v_pec = sig_int, v_pec
        L.sn_n = tab.shape[0]
        L.sn_z = self._hold(tab[:, 0])
        L.sn_m = self._hold(tab[:, 1])
        L.sn_s = self._hold(tab[:, 2])
        L.sn_c = self._hold(tab[:, 3])
        L.sn_cov = self._hold(tab[:, 4:10])
        return self

    # -- Gaussian-data likelihoods: BAO (bao.c:40-77), CMBDistPrior (wmap.c:916-942)
    def _gauss(self, L, mean, covinv):
        cov = np.linalg.inv(_arr(covinv))          # mvdens_inverse at init
        L.g_ndim = len(mean)
        L.g_mean = self._hold(mean)
        L.g_chol = self._hold(np.linalg.cholesky(cov))

    def add_bao(self, data=BAO_BOSS12_DZ, cosmo=None, special="none"):
        L = self._new_like("BAO", special, cosmo or COSMO_DP)
        L.bao_method = A.BAO_METHOD[data["method"]]
        self._gauss(L, data["mean"], data["covinv"])
        L.g_z = self._hold(data["z"])
        return self

    def add_cmbdp(self, data=WMAP7_DP, cosmo=None, special="none"):
        L = self._new_like("CMBDistPrior", special, cosmo or COSMO_DP)
        self._gauss(L, data["mean"], data["covinv"])
        return self

    # -- analytic targets (param.c:1485-1537) ---------------------------------
    def add_mix(self, wght, mean, cov, df=-1, special="none"):
        wght, mean, cov = _arr(wght), _arr(mean), _arr(cov)
        K, d = mean.shape
        L = self._new_like("MixMvdens" if K > 1 else "Mvdens", special)
        L.mix_ncomp, L.mix_ndim, L.mix_df = K, d, df
        L.mix_wght = self._hold(wght)
        L.mix_mean = self._hold(mean)
        L.mix_chol = self._hold(np.stack([np.linalg.cholesky(c) for c in cov]))
        return self

    def add_banana(self, b=0.03, sigma1sq=100.0):
        L = self._new_like("BANANA")
        L.banana_b, L.banana_sigma1sq = b, sigma1sq
        return self

    def set_prior(self, mean, cov, indprior=None):
        """Gaussian prior, param.c:42-68,1009-1026."""
        mean, cov = _arr(mean), _arr(cov)
        self.t.prior_ndim = len(mean)
        self.t.prior_mean = self._hold(mean)
        self.t.prior_chol = self._hold(np.linalg.cholesky(cov))
        if indprior is None:
            self.t.nprior = 0
        else:
            self.t.nprior = int(sum(indprior))
            for j, v in enumerate(indprior):
                self.t.indprior[j] = int(v)
        return self


def load_sn_table(path):
    """Parsed SN sample: rows z m s c Vmm Vss Vcc Cms Cmc Csc (see
    tests/golden/make_sn_fixture.py)."""
    sig_int, v_pec, rows = 0.0, 0.0, []
    for line in open(path):
        line = line.strip()
        if not line or line.startswith("#"):
            continue
        if line.startswith("@"):
            k, v = line.split()
            if k == "@sig_int":
                sig_int = float(v)
            elif k == "@v_pec":
                v_pec = float(v)
            continue
        rows.append([float(t) for t in line.split()])
    return np.array(rows, dtype=np.float64), sig_int, v_pec


# ---- the BASELINE.json configurations -----------------
def target_sn_demo():
    """C1/C2: Demo/MC_Demo/SN/config_pmc:6-24 (flat wCDM, 5 parameters)."""
    return TargetSpec(["Omega_m", "w_0_de", "M", "alpha", "beta"],
                      [0.0, -3.5, 19.1, 0.5, -3.5],
                      [1.2, 0.5, 19.8, 2.6, -0.8]).add_snia()


def target_sn_curved():
    """C2 variant named in BASELINE.json: Omega_m Omega_de M alpha beta."""
    return TargetSpec(["Omega_m", "Omega_de", "M", "alpha", "beta"],
                      [0.0, 0.0, 19.1, 0.5, -3.5],
                      [1.2, 1.6, 19.8, 2.6, -0.8]).add_snia()


def target_sn_bao_w0wa():
    """C4: SN + BAO d_z, w0-wa dark energy, 7 parameters."""
    return (TargetSpec(["Omega_m", "Omega_de", "w_0_de", "w_1_de", "M", "alpha", "beta"],
                       [0.05, 0.0, -3.0, -3.0, 19.1, 0.5, -3.5],
                       [1.0, 1.5, 0.0, 2.0, 19.8, 2.6, -0.8])
            .add_snia(cosmo=COSMO_DP).add_bao(BAO_BOSS12_DZ))


def target_cmb_bao_sn():
    """C5: WMAP7 distance priors + BAO + SN, 8 parameters."""
    return (TargetSpec(["Omega_b", "Omega_m", "Omega_de", "h_100", "w_0_de", "M", "alpha", "beta"],
                       [0.02, 0.1, 0.3, 0.5, -2.5, 19.1, 0.5, -3.5],
                       [0.08, 0.6, 1.1, 0.9, -0.3, 19.8, 2.6, -0.8])
            .add_cmbdp().add_bao(BAO_BOSS12_DZ).add_snia(cosmo=COSMO_DP))


def target_banana(d=20, b=0.03, sigma1sq=100.0):
    """C3: Wraith et al. 2009 twisted Gaussian."""
    lo = [-40.0, -70.0] + [-10.0] * (d - 2)
    hi = [40.0, 30.0] + [10.0] * (d - 2)
    return TargetSpec(["dummy%d" % j for j in range(d)], lo, hi).add_banana(b, sigma1sq)


def target_gauss2d():
    """Demo/tempering/1_mvnorm_2D_temp_none/config_pmc:14-21."""
    return TargetSpec(["dummy0", "dummy1"], [0.0, 0.0], [1.0, 1.0]).add_mix(
        [1.0], [[0.5, 0.5]], [[[0.01, 0.0], [0.0, 0.02]]])


# ---- synthetic proposals ---------------------------------------
# Manual/manual.tex:3226-3234: a plausible SN posterior (mean, covariance)
SN_POST_MEAN = np.array([0.38559, -1.5238, 19.338, 1.3692, -2.4358])
SN_POST_COV = np.array([
    [0.0053677, -0.025608, 0.00066748, -0.0011893, 0.00087517],
    [-0.025608, 0.16837, -0.0079163, 0.0027364, -0.0035709],
    [0.00066748, -0.0079163, 0.0011077, 0.0010986, -0.00067815],
    [-0.0011893, 0.0027364, 0.0010986, 0.016716, 0.0026266],
    [0.00087517, -0.0035709, -0.00067815, 0.0026266, 0.014881]])


def proposal_sn(K=10, seed=20090903, fshift=0.02, fvar=1.8):
    """K Gaussians around the manual's SN posterior, mimicking `fisher_rshift`
    (param.c:472-489,536-551): mean shifted by U(-1,1)*fshift*(max-min),
    covariance fvar * posterior covariance, equal weights."""
    rng = np.random.default_rng(seed)
    lo, hi = target_sn_demo().box
    mean = SN_POST_MEAN[None, :] + rng.uniform(-1, 1, (K, 5)) * fshift * (hi - lo)[None, :]
    cov = np.repeat((fvar * SN_POST_COV)[None], K, axis=0)
    return np.full(K, 1.0 / K), mean, cov


def proposal_banana(K=10, d=20, seed=20090307):
    rng = np.random.default_rng(seed)
    sd = np.array([5.0, 5.0] + [1.0] * (d - 2))
    mean = rng.normal(size=(K, d)) * sd[None, :]
    cov = np.repeat(np.diag([50.0, 25.0] + [2.0] * (d - 2))[None], K, axis=0)
    return np.full(K, 1.0 / K), mean, cov


def proposal_generic(spec, K, seed, center, sigma, fshift=0.5, fvar=1.5):
    """K Gaussians N(center + shift_k, fvar*diag(sigma^2))."""
    rng = np.random.default_rng(seed)
    center, sigma = _arr(center), _arr(sigma)
    d = len(center)
    mean = center[None, :] + rng.uniform(-1, 1, (K, d)) * fshift * sigma[None, :]
    cov = np.repeat(np.diag(fvar * sigma ** 2)[None], K, axis=0)
    return np.full(K, 1.0 / K), mean, cov
