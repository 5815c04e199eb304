"""CPU tests of the host side: the C-ABI library loads and exports every symbol
of include/pmcb200.h, descriptors have the C layout, the product path fails
loudly without a GPU, and the package never touches the oracle."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from cosmopmc_b200 import _abi as A
from cosmopmc_b200 import targets as T

ROOT = A.ROOT
REF = "/root/reference"


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "pmcb200.h")).read()
    declared = set(re.findall(r"\b(pmcb200_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"pmcb200_ctx"}
    assert len(declared) >= 25
    lib = A.load_library()                      # binds every entry of SYMBOLS
    assert declared == set(A.SYMBOLS), declared ^ set(A.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name)
    assert lib.pmcb200_version() == 100


def test_struct_layouts_match_c_header(tmp_path):
    """sizeof/offsetof of the ctypes mirrors against the C compiler's view of pmcb200.h"""
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "pmcb200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(pmcb200_cosmo_t),sizeof(pmcb200_like_t),sizeof(pmcb200_target_t),sizeof(pmcb200_stats_t),'
                   'offsetof(pmcb200_like_t,sn_z),offsetof(pmcb200_like_t,g_z),offsetof(pmcb200_target_t,like),'
                   'offsetof(pmcb200_target_t,prior_mean));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc", "-I",
                           os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(t) for t in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(A.Cosmo), C.sizeof(A.Like), C.sizeof(A.Target), C.sizeof(A.Stats),
            A.Like.sn_z.offset, A.Like.g_z.offset, A.Target.like.offset, A.Target.prior_mean.offset]
    assert got == want


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = A.load_library()
    assert lib.pmcb200_device_count() == 0
    h = C.c_void_p()
    assert lib.pmcb200_create(0, None, C.byref(h)) == A.ERR["CUDA"]
    assert not h.value
    from cosmopmc_b200.pmc import PMC
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        PMC(0)


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        A.load_library(str(tmp_path / "nope.so"))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "cosmopmc_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.path.basename(dirpath) == "build":
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower() or f == "common.cuh" and "oracle orc_select_component" in txt, f
    hdr = open(os.path.join(ROOT, "include", "pmcb200.h")).read()
    assert "oracle" not in hdr.lower()


def test_target_builders_describe_baseline_configs():
    c1 = T.target_sn_demo()
    assert c1.t.npar == 5 and c1.t.ndata == 1 and c1.t.like[0].kind == A.LIKE["SNIa"]
    assert c1.t.like[0].sn_n == 307
    assert [c1.t.like[0].par[j] for j in range(5)] == [0, 14, 37, 38, 39]       # par_t values
    assert abs(c1.t.like[0].sn_sig_int - 0.15) < 1e-15 and c1.t.like[0].sn_v_pec == 300.0
    c5 = T.target_cmb_bao_sn()
    assert c5.t.npar == 8 and [c5.t.like[i].kind for i in range(3)] == [6, 7, 3]
    c3 = T.target_banana(20)
    assert c3.t.npar == 20 and c3.t.like[0].kind == 100
    w, m, cov = T.proposal_sn(10)
    assert w.shape == (10,) and abs(w.sum() - 1) < 1e-15 and m.shape == (10, 5)
    lo, hi = c1.box
    assert np.all(m > lo) and np.all(m < hi)
    for c in cov:
        np.linalg.cholesky(c)


def test_sn_fixture_matches_reference_file():
    """the committed parsed SN table is the reference's data file (only checked where the tree exists)"""
    if not os.path.isdir(REF):
        pytest.skip("reference tree not present")
    tab, sig_int, v_pec = T.load_sn_table(T.SN_FIXTURE)
    rows = []
    for line in open(os.path.join(REF, "data/Sn/Union/sne_union_marek.list")):
        line = line.strip()
        if not line or line[0] in "#@":
            continue
        f = [float(t) for t in line.split()[1:11]]
        rows.append([f[0], f[1], f[3], f[5], f[2] ** 2, f[4] ** 2, f[6] ** 2, f[7], f[8], f[9]])
    assert np.array_equal(tab, np.array(rows))
    assert (sig_int, v_pec) == (0.15, 300.0)


def test_par_t_values_match_reference():
    """PMCB200_P_* / PMCB200_LIKE_* are the reference's par_t / data_t enum values."""
    if not os.path.isdir(REF):
        pytest.skip("reference tree not present")
    def enum_names(path, typedef):
        txt = open(path).read()
        body = re.search(r"typedef enum\s*\{(.*?)\}\s*" + typedef, txt, re.S).group(1)
        body = re.sub(r"//.*", "", body)
        return [t.strip() for t in body.split(",") if t.strip()]
    par = enum_names(os.path.join(REF, "tools/include/par.h"), "par_t")
    hdr = open(os.path.join(ROOT, "include", "pmcb200.h")).read()
    ours = dict((n, int(v)) for n, v in re.findall(r"PMCB200_P_([A-Za-z0-9_]+)\s*=\s*(\d+)", hdr))
    assert len(ours) >= 20
    for name, val in ours.items():
        assert par[val] == "p_" + name, (name, val, par[val])
    data = enum_names(os.path.join(REF, "wrappers/include/types.h"), "data_t")
    kinds = dict((n, int(v)) for n, v in re.findall(r"PMCB200_LIKE_([A-Za-z]+)\s*=\s*(\d+)", hdr))
    for name, val in kinds.items():
        if name != "BANANA":
            assert data[val] == name
    for name, val in A.P.items():
        if name == "omegab100":
            assert par[val] == "p_100_omegab"
        else:
            assert par[val] == "p_" + name


def test_pmclib_named_host_api_cpu(tmp_path):
    """tests/c/test_host_cpu.c: error stack, mvdens/mix_mvdens formats and helpers, parabox,
    gsl shim, pmc_simu container, loud failure of the batched calls without a GPU."""
    exe = tmp_path / "test_host_cpu"
    libdir = os.path.join(ROOT, "cosmopmc_b200")
    A.load_library()
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.check_call([gcc, "-std=gnu99", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c", "test_host_cpu.c"), "-o", str(exe),
                           "-L", libdir, "-lpmc_b200", "-Wl,-rpath," + libdir, "-lm"])
    out = subprocess.run([str(exe), str(tmp_path)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.startswith("ok ") and int(out.stdout.split()[1]) > 100


def test_device_math_helpers_on_host(tmp_path):
    """tests/c/test_device_math.cu (nvcc, runs on the CPU): the column-oriented multi-sample whitening of
    k_weights_multi / k_em_stats_mma is bit-identical to the row-oriented forward substitution and agrees with a
    long-double reference; the tensor-core kernel's feature enumeration covers the statistics block exactly once."""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    exe = str(tmp_path / "test_device_math")
    subprocess.check_call([nvcc, "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a",
                           "--extended-lambda", "-o", exe, os.path.join(ROOT, "tests", "c", "test_device_math.cu")],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "device math host check ok" in out.stdout, out.stdout + out.stderr


def test_sn_tile_plan_of_the_tensor_core_kernel():
    """Column plan of the tensor-core SN kernel (pmcb200_sn_tile_plan: host only).  Every supernova sits in exactly one
    column; a primary tile holds distinct redshifts; the supernova in column j of a secondary tile has the redshift of
    column j of the primary tile before it; the Union sample (307 supernovae, 241 redshifts) takes 31 + 12 tiles."""
    import ctypes as C
    lib = A.load_library()

    def plan(z):
        z = np.ascontiguousarray(z, dtype=np.float64)
        need = -lib.pmcb200_sn_tile_plan(len(z), z.ctypes.data, 0, None, None)
        assert need > 0
        sec = np.zeros(need, np.int32); col = np.zeros(8 * need, np.int32)
        nt = lib.pmcb200_sn_tile_plan(len(z), z.ctypes.data, need, sec.ctypes.data, col.ctypes.data)
        assert nt == need
        return sec, col.reshape(nt, 8)

    def check(z):
        z = np.asarray(z, dtype=np.float64)
        sec, col = plan(z)
        live = col[col >= 0]
        assert sorted(live.tolist()) == list(range(len(z)))            # every supernova exactly once
        assert sec[0] == 0
        prim = None
        for t in range(len(sec)):
            zc = np.where(col[t] >= 0, z[np.maximum(col[t], 0)], z[-1 - np.minimum(col[t], -1)])     # redshift of every column
            if not sec[t]:
                prim = zc
                lz = zc[col[t] >= 0]
                assert len(set(lz.tolist())) == len(lz) and len(lz) >= 1   # distinct redshifts, first supernova of each
            else:
                assert np.array_equal(zc, prim)                            # same column = same redshift as in the primary tile
                assert (col[t] >= 0).any()                                 # no empty secondary tile
        return sec, col

    rng = np.random.default_rng(3)
    sec, col = check([0.5]); assert len(sec) == 1 and (col[0] >= 0).sum() == 1
    sec, col = check([0.3] * 9); assert sec.tolist() == [0] + [1] * 8
    check([0.02, 0.02, 0.4, 0.4, 0.4, 0.9, 1.4])
    check(np.round(rng.random(100) * 1.5 + 0.01, 2))        # many repeated redshifts
    check(rng.random(64) + 0.01)                            # none repeated: 8 primary tiles
    from cosmopmc_b200 import targets as T
    tab, _, _ = T.load_sn_table(T.SN_FIXTURE)
    sec, col = check(tab[:, 0])
    assert (sec == 0).sum() == 31 and (sec == 1).sum() == 12
