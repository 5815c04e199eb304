#!/usr/bin/env python
"""Build the reference's OWN, UNCHANGED PMC driver against this repository's library.

Compiles, where they lie under /root/reference (never copied into the repo):
  exec/cosmo_pmc.c exec/exec_helper.c
  wrappers/src/{param,sn,bao,wmap,wrappers,init_wrappers,timexec}.c
  tools/src/{config,par,nhist}.c
plus this repo's glue (cosmopmc_b200/glue/*.c, compiled with the reference's headers)
and links them with cosmopmc_b200/libpmc_b200.so, which provides the pmclib /
nicaea / gsl / MPI-named symbols (link line of the reference: exec/Makefile:37-41).
Output: build_ref/cosmo_pmc (+ a copy of the SN demo inputs in build_ref/demo_SN/,
both git-ignored, shipped to the GPU box).  Container-only: needs /root/reference.
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("COSMOPMC_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "build_ref")
GCC = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
SRC = ["exec/cosmo_pmc.c", "exec/exec_helper.c", "wrappers/src/param.c", "wrappers/src/sn.c", "wrappers/src/bao.c",
       "wrappers/src/wmap.c", "wrappers/src/wrappers.c", "wrappers/src/init_wrappers.c", "wrappers/src/timexec.c",
       "tools/src/config.c", "tools/src/par.c", "tools/src/nhist.c"]
GLUE = ["cosmopmc_b200/glue/pmcb200_glue.c", "cosmopmc_b200/glue/out_of_scope_stubs.c"]
INC = ["-I", os.path.join(ROOT, "include"), "-I", os.path.join(REF, "exec"), "-I", os.path.join(REF, "wrappers/include"),
       "-I", os.path.join(REF, "tools/include")]
# the reference's own flags: -std=gnu9x, no optimisation, -DCOMM_DEBUG (Makefile.main:22-26,239)
CFLAGS = ["-std=gnu9x", "-g", "-w"]


def build():
    if not os.path.isdir(REF):
        raise SystemExit("reference tree %s not present (this recipe only runs in the build container)" % REF)
    sys.path.insert(0, ROOT)
    from cosmopmc_b200 import build as b
    b.build()
    os.makedirs(os.path.join(OUT, "obj"), exist_ok=True)
    objs = []
    for s, base in [(x, REF) for x in SRC] + [(x, ROOT) for x in GLUE]:
        o = os.path.join(OUT, "obj", os.path.basename(s).replace(".c", ".o"))
        subprocess.check_call([GCC] + CFLAGS + INC + ["-c", os.path.join(base, s), "-o", o])
        objs.append(o)
    libdir = os.path.join(ROOT, "cosmopmc_b200")
    exe = os.path.join(OUT, "cosmo_pmc")
    subprocess.check_call([GCC, "-o", exe] + objs + ["-L", libdir, "-lpmc_b200", "-Wl,-rpath,$ORIGIN/../cosmopmc_b200",
                                                      "-lm"])
    # the SN demo's inputs (Demo/MC_Demo/SN): config + data + parameter files, as bin/cosmo_pmc.pl stages them
    demo = os.path.join(OUT, "demo_SN")
    os.makedirs(demo, exist_ok=True)
    for f in ["Demo/MC_Demo/SN/config_pmc", "data/Sn/Union/sne_union_marek.list", "par_files/cosmo_SN.par",
              "par_files/cosmo.par"]:
        shutil.copy(os.path.join(REF, f), demo)
    # the tempering demos (Demo/tempering/README.md: evidence known answers)
    for sub in ["1_mvnorm_2D_temp_none", "2_mixmvnorm_2D_temp_none"]:
        dst = os.path.join(OUT, "demo_" + sub)
        os.makedirs(dst, exist_ok=True)
        shutil.copy(os.path.join(REF, "Demo/tempering", sub, "config_pmc"), dst)
    return exe


if __name__ == "__main__":
    print(build())
