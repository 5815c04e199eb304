"""Host-side mirror of the pmclib call sequence of `run_pmc_iteration_MPI`
(reference exec/cosmo_pmc.c:293-402) on top of the C-ABI.

Method names follow the pmclib functions they stand for so the parity tests
read like the reference's iteration body.  torch is used only to own device
memory / streams and for torch.distributed (NCCL); every computation is a
kernel inside libpmc_b200.so.  No fallback: construction raises without the
library or without a CUDA device.
"""
import ctypes as C

import numpy as np
import torch

from . import _abi as A


class PMCError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("pmcb200 error %d: %s" % (code, msg))
        self.code = code


def _dp(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class PMC:
    """One GPU's share of a PMC run (one process per GPU)."""

    def __init__(self, device=0, use_torch_stream=True):
        self.lib = A.load_library()
        if not torch.cuda.is_available():
            raise RuntimeError("cosmopmc_b200 needs a CUDA device (no CPU fallback)")
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream if use_torch_stream else None
        h = C.c_void_p()
        rc = self.lib.pmcb200_create(device, C.c_void_p(stream) if stream else None, C.byref(h))  # 0 = default stream
        if rc:
            raise PMCError(rc, "pmcb200_create failed (no usable CUDA device?)")
        self.h = h
        self.spec = None
        self.K = self.d = 0
        self.df = -1

    def close(self):
        if getattr(self, "h", None):
            self.lib.pmcb200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise PMCError(rc, self.lib.pmcb200_last_error(self.h).decode())

    # -- set-up ---------------------------------------------------------------
    def set_target(self, spec):
        self.spec = spec              # keeps the host arrays alive
        self._ck(self.lib.pmcb200_set_target(self.h, C.byref(spec.t)))

    def set_proposal(self, wght, mean, chol=None, cov=None, df=-1):
        wght = np.ascontiguousarray(wght, dtype=np.float64)
        mean = np.ascontiguousarray(mean, dtype=np.float64)
        K, d = mean.shape
        self.K, self.d, self.df = K, d, df
        if chol is not None:
            m = np.ascontiguousarray(chol, dtype=np.float64)
            fn = self.lib.pmcb200_set_proposal
        else:
            m = np.ascontiguousarray(cov, dtype=np.float64)
            fn = self.lib.pmcb200_set_proposal_cov
        assert m.shape == (K, d, d) and wght.shape == (K,)
        self._ck(fn(self.h, K, d, df, wght.ctypes.data, mean.ctypes.data, m.ctypes.data))

    def get_proposal(self):
        K, d = self.K, self.d
        w, m = np.empty(K), np.empty((K, d))
        ch, cov = np.empty((K, d, d)), np.empty((K, d, d))
        self._ck(self.lib.pmcb200_get_proposal(self.h, w.ctypes.data, m.ctypes.data,
                                               ch.ctypes.data, cov.ctypes.data))
        return w, m, ch, cov

    # -- device buffers -----------------------------------------------------------
    def alloc(self, N):
        dev, n1 = self.device, max(N, 1)
        return dict(X=torch.empty((n1, self.d), dtype=torch.float64, device=dev),
                    idx=torch.empty(n1, dtype=torch.int32, device=dev),
                    flg=torch.empty(n1, dtype=torch.int16, device=dev),
                    logw=torch.empty(n1, dtype=torch.float64, device=dev))

    def stat_block_len(self):
        return int(self.lib.pmcb200_stat_block_len(self.h))

    # -- pmclib-named stages (device tensors) ---------------------------------------
    def simulate_mix_mvdens(self, N, seed, it, offset, X, idx, flg):
        self._ck(self.lib.pmcb200_simulate(self.h, N, seed, it, offset, _dp(X), _dp(idx), _dp(flg)))

    def simulate_from_draws(self, u, z, X, idx, flg):
        self._ck(self.lib.pmcb200_simulate_from_draws(self.h, u.numel(), _dp(u), _dp(z), _dp(X),
                                                      _dp(idx), _dp(flg)))

    def mix_mvdens_log_pdf(self, X):
        N = X.shape[0]
        out = torch.empty(max(N, 1), dtype=torch.float64, device=self.device)
        self._ck(self.lib.pmcb200_proposal_log_pdf(self.h, N, _dp(X), _dp(out)))
        return out[:N]

    def posterior_log_pdf(self, X):
        N = X.shape[0]
        out = torch.empty(max(N, 1), dtype=torch.float64, device=self.device)
        err = torch.zeros(max(N, 1), dtype=torch.int32, device=self.device)
        self._ck(self.lib.pmcb200_posterior_log_pdf(self.h, N, _dp(X), _dp(out), _dp(err)))
        return out[:N], err[:N]

    def map_params(self, idata, X):
        """device apply_params of data set idata (parity probe): (N, 16) models + error flags"""
        N = X.shape[0]
        out = torch.zeros((max(N, 1), 16), dtype=torch.float64, device=self.device)
        err = torch.zeros(max(N, 1), dtype=torch.int32, device=self.device)
        self._ck(self.lib.pmcb200_map_params(self.h, idata, N, _dp(X), _dp(out), _dp(err)))
        return out[:N], err[:N]

    def get_importance_weight(self, X, flg, logw, beta=1.0):
        self._ck(self.lib.pmcb200_importance_weights(self.h, X.shape[0], _dp(X), beta, _dp(flg), _dp(logw)))

    def em_local(self, X, idx, flg, logw, block, N=None):
        N = X.shape[0] if N is None else N
        self._ck(self.lib.pmcb200_em_local(self.h, N, _dp(X), _dp(idx), _dp(flg), _dp(logw), _dp(block)))

    def update_prop_rb(self, nranks, all_blocks, N_global):
        """em_finish: combine the rank blocks, M-step, install the new proposal."""
        st = A.Stats()
        rc = self.lib.pmcb200_em_finish(self.h, nranks, _dp(all_blocks), N_global, C.byref(st))
        self.last_stats = st.as_dict()
        self._ck(rc)
        return self.last_stats

    def normalize_importance_weight(self, flg, w, N=None):
        N = w.shape[0] if N is None else N
        self._ck(self.lib.pmcb200_normalize_weights(self.h, N, _dp(flg), _dp(w)))

    # -- whole iteration ---------------------------------------------------------------
    def iteration_local(self, N, seed, it, offset, beta, block, bufs=None):
        b = bufs or {}
        self._ck(self.lib.pmcb200_iteration_local(
            self.h, N, seed, it, offset, beta, _dp(b.get("X")), _dp(b.get("idx")),
            _dp(b.get("flg")), _dp(b.get("logw")), _dp(block)))

    def iteration_host(self, N, seed, it, beta=1.0, hX=None, hidx=None, hflg=None, hw=None):
        """Single GPU, HOST (numpy or pinned torch) buffers."""
        def hp(a):
            if a is None:
                return None
            return C.c_void_p(a.data_ptr() if isinstance(a, torch.Tensor) else a.ctypes.data)
        st = A.Stats()
        rc = self.lib.pmcb200_iteration_host(self.h, N, seed, it, beta, hp(hX), hp(hidx), hp(hflg),
                                             hp(hw), C.byref(st))
        self.last_stats = st.as_dict()
        self._ck(rc)
        return self.last_stats

    def iteration_host_begin(self, N, seed, it, beta=1.0, hX=None, hidx=None, hflg=None, hw=None):
        """as iteration_host, but returns with the host copies still draining (host_wait)"""
        def hp(a):
            return None if a is None else C.c_void_p(a.data_ptr() if isinstance(a, torch.Tensor) else a.ctypes.data)
        st = A.Stats()
        rc = self.lib.pmcb200_iteration_host_begin(self.h, N, seed, it, beta, hp(hX), hp(hidx), hp(hflg), hp(hw), C.byref(st))
        self.last_stats = st.as_dict()
        self._ck(rc)
        return self.last_stats

    def shard_weights_host_begin(self, N, hw):
        p = C.c_void_p(hw.data_ptr() if isinstance(hw, torch.Tensor) else hw.ctypes.data)
        self._ck(self.lib.pmcb200_shard_weights_host_begin(self.h, N, p))

    def samples_host_begin(self, N, hX=None, hidx=None):
        """sample array / component indices of the most recent iteration (left on the device by hX=None, hidx=None)
        to the host, on request"""
        def hp(a):
            return None if a is None else C.c_void_p(a.data_ptr() if isinstance(a, torch.Tensor) else a.ctypes.data)
        self._ck(self.lib.pmcb200_samples_host_begin(self.h, N, hp(hX), hp(hidx)))

    def host_wait(self, lag=0):
        self._ck(self.lib.pmcb200_host_wait(self.h, lag))

    def iteration_shard_host(self, N, seed, it, offset, beta, block, hX=None, hidx=None, hflg=None):
        def hp(a):
            return None if a is None else C.c_void_p(a.data_ptr() if isinstance(a, torch.Tensor) else a.ctypes.data)
        self._ck(self.lib.pmcb200_iteration_shard_host(self.h, N, seed, it, offset, beta, hp(hX), hp(hidx),
                                                        hp(hflg), _dp(block)))

    def shard_weights_host(self, N, hw):
        p = C.c_void_p(hw.data_ptr() if isinstance(hw, torch.Tensor) else hw.ctypes.data)
        self._ck(self.lib.pmcb200_shard_weights_host(self.h, N, p))

    # -- weighted post-processing of a stored sample (device tensors in, host results out) --
    def post_moments(self, X, flg, w, with_cov=True):
        N, d = X.shape
        mean, cov = np.empty(d), np.empty((d, d))
        self._ck(self.lib.pmcb200_post_moments(self.h, N, d, _dp(X), _dp(flg), _dp(w), mean.ctypes.data,
                                               cov.ctypes.data if with_cov else None))
        return (mean, cov) if with_cov else mean

    def post_sigma(self, X, flg, w, a, center, conf):
        N, d = X.shape
        cf = np.ascontiguousarray(conf, dtype=np.float64)
        sig = np.empty(6)
        med, nf = C.c_double(), C.c_int64()
        self._ck(self.lib.pmcb200_post_sigma(self.h, N, d, _dp(X), _dp(flg), _dp(w), a, center, cf.ctypes.data,
                                             sig.ctypes.data, C.byref(med), C.byref(nf)))
        return sig, med.value, nf.value

    def post_histogram(self, X, flg, w, pidx, nbins, limits):
        N, d = X.shape
        pi = np.ascontiguousarray(pidx, dtype=np.int32)
        nb = np.ascontiguousarray(nbins, dtype=np.int32)
        lim = np.ascontiguousarray(limits, dtype=np.float64)
        tdim = int(np.prod(nb))
        cnt, sw, sw2 = np.empty(tdim), np.empty(tdim), np.empty(tdim)
        self._ck(self.lib.pmcb200_post_histogram(self.h, N, d, _dp(X), _dp(flg), _dp(w), len(pi), pi.ctypes.data,
                                                 nb.ctypes.data, lim.ctypes.data, cnt.ctypes.data, sw.ctypes.data,
                                                 sw2.ctypes.data))
        return cnt, sw, sw2

    def fisher_matrix(self, pos, h, diag_only=False):
        """go_fishing.c fisher_element over all (a, b): one batched posterior launch."""
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        h = np.ascontiguousarray(h, dtype=np.float64)
        d = pos.size
        F = np.empty((d, d))
        nbad = C.c_int(0)
        self._ck(self.lib.pmcb200_fisher_host(self.h, pos.ctypes.data, h.ctypes.data, int(bool(diag_only)),
                                              F.ctypes.data, C.byref(nbad)))
        return F

    def launch_count(self):
        return int(self.lib.pmcb200_launch_count(self.h))

    def counters(self):
        out = (C.c_int64 * 7)()
        self._ck(self.lib.pmcb200_counters_ex(self.h, out, 7))
        return dict(sn_evals=int(out[0]), sn_zsteps=int(out[1]), gen_evals=int(out[2]), gen_integrals=int(out[3]),
                    sn_spec=int(out[4]), sn_exact=int(out[5]), cmb_spec=int(out[6]))

    def fp64_peak_tflops(self):
        v = C.c_double()
        self._ck(self.lib.pmcb200_fp64_peak(self.h, C.byref(v)))
        return v.value

    def sync(self):
        self._ck(self.lib.pmcb200_sync(self.h))


def iteration_host_multi(pmcs, N, seed, it, beta=1.0, hX=None, hidx=None, hflg=None, hw=None):
    """pmcb200_iteration_host_multi: one iteration sharded over several contexts of THIS
    process (one per GPU, or several on one GPU); statistics blocks travel by peer copies."""
    def hp(a):
        if a is None:
            return None
        return C.c_void_p(a.data_ptr() if isinstance(a, torch.Tensor) else a.ctypes.data)
    lib = pmcs[0].lib
    arr = (C.c_void_p * len(pmcs))(*[p.h for p in pmcs])
    st = A.Stats()
    rc = lib.pmcb200_iteration_host_multi(arr, len(pmcs), N, seed, it, beta, hp(hX), hp(hidx), hp(hflg),
                                          hp(hw), C.byref(st))
    if rc:
        raise PMCError(rc, lib.pmcb200_last_error(pmcs[0].h).decode())
    return st.as_dict()


def run_iteration_distributed(pmc, N_global, seed, it, beta, block, all_blocks, bufs=None,
                              rank=0, world=1):
    """One PMC iteration sharded over `world` ranks (one GPU each): contiguous
    sample-index shards, ONE collective (all-gather of the EM stat blocks over
    NCCL/NVLink), identical fixed-order combine + M-step on every rank."""
    import torch.distributed as dist
    per = (N_global + world - 1) // world
    off = rank * per
    n_loc = max(0, min(per, N_global - off))
    pmc.iteration_local(n_loc, seed, it, off, beta, block, bufs)
    if world > 1:
        dist.all_gather_into_tensor(all_blocks, block)
        src = all_blocks
    else:
        src = block
    return pmc.update_prop_rb(world, src, N_global)
