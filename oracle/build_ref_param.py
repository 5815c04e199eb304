#!/usr/bin/env python
"""oracle/_ref/libref_param.so -- the REFERENCE'S OWN parameter mapping, prior and posterior assembly,
compiled where it lies, with a recording stand-in for the absent nicaea library.

TEST INFRASTRUCTURE ONLY (see pmc_oracle.h).  Container-only: needs /root/reference.

Compiled unchanged (nothing is copied into the repo):
    wrappers/src/param.c          read_config_base (:73-200, logpr_default :124-129), posterior_log_pdf_common
                                  (:958-1041), prior_log_pdf_special (:1055-1101), set_base_parameters (:1544-1661)
    wrappers/src/{sn,bao,wmap}.c  likeli_SNIa (:138-281), likeli_BAO (:80-184), likeli_CMBDistPrior (:945-1049)
    wrappers/src/{wrappers,init_wrappers}.c, tools/src/{config,par}.c   the plug-in registry and the config reader
Their environment:
  * nicaea (absent): oracle/ref_param_record.c -- struct management + stubs that RECORD the model each
    nicaea entry point receives and return values chosen by the test;
  * pmclib's plain-C utilities (error stack, mvdens text reader / inverse / log-pdf, sm2_*, gsl_ran_flat): this
    repo's cosmopmc_b200/host/{errorlist,io,mvdens,maths,gsl_shim}.c -- host utilities without any device code;
  * every other undefined symbol (lensing / halo-model / topology plug-ins, histogram writers, fork helpers:
    functions of param.c the tests never reach) becomes a generated abort() stub.
Output goes to oracle/_ref/ only (git-ignored, shipped to the GPU box)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("COSMOPMC_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
LIB = os.path.join(OUT, "libref_param.so")
GCC = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
REF_SRC = ["wrappers/src/param.c", "wrappers/src/sn.c", "wrappers/src/bao.c", "wrappers/src/wmap.c",
           "wrappers/src/wrappers.c", "wrappers/src/init_wrappers.c", "tools/src/config.c", "tools/src/par.c"]
OWN_SRC = ["oracle/ref_param_record.c", "cosmopmc_b200/host/errorlist.c", "cosmopmc_b200/host/io.c",
           "cosmopmc_b200/host/mvdens.c", "cosmopmc_b200/host/maths.c", "cosmopmc_b200/host/gsl_shim.c"]
INC = ["-I", os.path.join(ROOT, "include"), "-I", os.path.join(REF, "exec"), "-I", os.path.join(REF, "wrappers/include"),
       "-I", os.path.join(REF, "tools/include")]
# -O0 -ffp-contract=off: the reference's arithmetic exactly as written (its own build uses -g without -O,
# Makefile.main:22-26)
CFLAGS = ["-std=gnu9x", "-O0", "-g", "-w", "-fPIC", "-ffp-contract=off"]


def build(force=False):
    srcs = [os.path.join(REF, s) for s in REF_SRC] + [os.path.join(ROOT, s) for s in OWN_SRC]
    if not os.path.isdir(REF):
        return LIB if os.path.exists(LIB) else None
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) > os.path.getmtime(s) for s in srcs + [__file__]):
        return LIB
    os.makedirs(OUT, exist_ok=True)
    objs = []
    for s in srcs:
        o = os.path.join(OUT, "rp_" + os.path.basename(s).replace(".c", ".o"))
        subprocess.check_call([GCC] + CFLAGS + INC + ["-c", s, "-o", o])
        objs.append(o)
    defined, undefined = set(), set()
    for o in objs:
        for line in subprocess.check_output(["nm", o], text=True).split("\n"):
            t = line.split()
            if len(t) == 2 and t[0] == "U":
                undefined.add(t[1])
            elif len(t) == 3 and t[1] in "TDBRCVW":
                defined.add(t[2])
    probe = os.path.join(OUT, "rp_probe.c")
    stubs = []
    for sym in sorted(u for u in undefined - defined if not u.startswith("_GLOBAL_")):
        with open(probe, "w") as f:     # a symbol the C library provides links on its own
            f.write("extern char %s; void *p_(void) { return &%s; }\n" % (sym, sym))
        r = subprocess.run([GCC, "-w", "-shared", "-fPIC", "-Wl,--no-undefined", probe, "-o", os.path.join(OUT, "rp_probe.so"),
                            "-lm"], capture_output=True)
        if r.returncode != 0:
            stubs.append(sym)
    with open(os.path.join(OUT, "rp_stubs.c"), "w") as f:
        f.write("/* generated: functions of param.c's environment that the parameter-mapping tests never reach */\n"
                "#include <stdio.h>\n#include <stdlib.h>\n")
        for sym in stubs:
            f.write("void %s(void) { fprintf(stderr, \"oracle/_ref: stub %s called\\n\"); abort(); }\n" % (sym, sym))
    subprocess.check_call([GCC, "-shared", "-fPIC", "-w", "-o", LIB] + objs + [os.path.join(OUT, "rp_stubs.c"), "-lm",
                                                                                 "-Wl,--no-undefined"])
    for f in ("rp_probe.c", "rp_probe.so"):
        if os.path.exists(os.path.join(OUT, f)):
            os.remove(os.path.join(OUT, f))
    return LIB, stubs


if __name__ == "__main__":
    print(build(force="-f" in sys.argv))
