"""Build recipe for libpmc_b200.so (sm_100a only, in-tree, parallel nvcc)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.environ.get("PMCB200_OBJ_DIR") or os.path.join(HERE, "build")
LIB = os.environ.get("PMCB200_LIB_OUT") or os.path.join(HERE, "libpmc_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "--extended-lambda", "-Xcompiler", "-fPIC,-O2"] + os.environ.get("PMCB200_NVCC_FLAGS", "").split()

# (object name, source, extra flags)
UNITS = [("pmcb200.o", "pmcb200.cu", []),
         ("k_cosmo.o", "k_cosmo.cu", []),
         ("k_post.o", "k_post.cu", []),
         ("k_mix0.o", "k_mix.cu", ["-DMIX_GROUP=0"]),
         ("k_mix1.o", "k_mix.cu", ["-DMIX_GROUP=1"]),
         ("k_mix2.o", "k_mix.cu", ["-DMIX_GROUP=2"]),
         ("k_mix3.o", "k_mix.cu", ["-DMIX_GROUP=3"])]


HOST = os.path.join(HERE, "host")
INC = os.path.join(os.path.dirname(HERE), "include")
GCC = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
HOST_UNITS = ["errorlist.c", "gsl_shim.c", "mvdens.c", "pmc.c", "io.c", "maths.c", "pmc_mpi.c", "nicaea.c"]


def _compile_host(src):
    obj = os.path.join(OBJ, "host_" + src.replace(".c", ".o"))
    cmd = [GCC, "-std=gnu99", "-O2", "-g", "-fPIC", "-Wall", "-Wno-format-truncation", "-I", INC,
           "-c", os.path.join(HOST, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return obj, r.returncode, r.stdout + r.stderr


def _host_deps():
    out = []
    for root, _, files in os.walk(INC):
        out += [os.path.join(root, f) for f in files if f.endswith(".h")]
    return out


def _deps():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))] + \
           [os.path.join(os.path.dirname(HERE), "include", "pmcb200.h")]


def _stale(target, srcs):
    return not os.path.exists(target) or any(os.path.getmtime(target) < os.path.getmtime(s) for s in srcs)


def _compile(unit, verbose):
    obj, src, extra = unit
    cmd = [NVCC] + FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + \
          ["-c", os.path.join(CSRC, src), "-o", os.path.join(OBJ, obj)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return obj, r.returncode, r.stdout + r.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    deps = _deps()
    todo = [u for u in UNITS
            if force or _stale(os.path.join(OBJ, u[0]), deps + [os.path.join(CSRC, u[1])])]
    if todo:
        with ThreadPoolExecutor(max_workers=min(len(todo), os.cpu_count() or 1)) as ex:
            for obj, rc, out in ex.map(lambda u: _compile(u, verbose), todo):
                if verbose or rc:
                    sys.stderr.write("== %s\n%s" % (obj, out))
                if rc:
                    raise RuntimeError("nvcc failed for %s" % obj)
    hdeps = _host_deps()
    htodo = [h for h in HOST_UNITS
             if force or _stale(os.path.join(OBJ, "host_" + h.replace(".c", ".o")), hdeps + [os.path.join(HOST, h)])]
    for h in htodo:
        obj, rc, out = _compile_host(h)
        if verbose or rc:
            sys.stderr.write("== %s\n%s" % (obj, out))
        if rc:
            raise RuntimeError("gcc failed for %s" % h)
    objs = [os.path.join(OBJ, u[0]) for u in UNITS] + \
           [os.path.join(OBJ, "host_" + h.replace(".c", ".o")) for h in HOST_UNITS]
    if todo or htodo or _stale(LIB, objs):
        subprocess.check_call([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a",
                               "-o", LIB] + objs)
    return LIB


if __name__ == "__main__":
    build(force="-f" in sys.argv, verbose="-v" in sys.argv)
