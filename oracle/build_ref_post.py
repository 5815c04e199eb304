#!/usr/bin/env python
"""oracle/_ref/libref_post.so -- the REFERENCE'S OWN post-processing code, compiled where it lies.

TEST INFRASTRUCTURE ONLY (see pmc_oracle.h).  Container-only: needs /root/reference.

Unlike the PMC iteration itself (pmclib / nicaea: absent, parity unpinned), the weighted
post-processing of a sample is in-tree reference code:
    exec/exec_helper.c   mean/median/sigma_from_psim, covariance_from_sample   (:63-349)
    tools/src/nhist.c    init_nd_histogram, acc_histogram, make_histogram      (:15-260)
This recipe compiles those two files unchanged (nothing is copied into the repo) into a shared
object that the tests call through ctypes, so the device kernels of cosmopmc_b200/csrc/k_post.cu
are pinned against the real reference.  What the two files need from absent libraries:
  * pmclib / gsl headers: this repo's API-compatible headers under include/ (declarations only);
  * the error stack and malloc_err (pmclib's errorlist): cosmopmc_b200/host/errorlist.c, a
    plain-C utility with no compute in it;
  * every other undefined symbol (file writers, mvdens, config readers -- functions of
    exec_helper.c that the tests never call) is satisfied by an abort() stub generated here, so
    the library loads without the product library.
Output goes to oracle/_ref/ only (git-ignored, shipped to the GPU box)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("COSMOPMC_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
LIB = os.path.join(OUT, "libref_post.so")
GCC = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
SRC = [os.path.join(REF, "exec/exec_helper.c"), os.path.join(REF, "tools/src/nhist.c"),
       os.path.join(ROOT, "cosmopmc_b200/host/errorlist.c")]
INC = ["-I", os.path.join(ROOT, "include"), "-I", os.path.join(REF, "exec"), "-I", os.path.join(REF, "wrappers/include"),
       "-I", os.path.join(REF, "tools/include")]


def build(force=False):
    if not os.path.isdir(REF):
        return LIB if os.path.exists(LIB) else None
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) > os.path.getmtime(s) for s in SRC):
        return LIB
    os.makedirs(OUT, exist_ok=True)
    objs = []
    for s in SRC:
        o = os.path.join(OUT, os.path.basename(s).replace(".c", ".o"))
        subprocess.check_call([GCC, "-std=gnu9x", "-O1", "-g", "-w", "-fPIC"] + INC + ["-c", s, "-o", o])
        objs.append(o)
    # undefined symbols of the three objects that neither they nor libc / libm define -> abort() stubs
    defined, undefined = set(), set()
    for o in objs:
        for line in subprocess.check_output(["nm", o], text=True).split("\n"):
            t = line.split()
            if len(t) == 2 and t[0] == "U":
                undefined.add(t[1])
            elif len(t) == 3 and t[1] in "TDBRCVW":
                defined.add(t[2])
    probe = os.path.join(OUT, "probe.c")
    need = sorted(u for u in undefined - defined if not u.startswith("_GLOBAL_"))
    stubs = []
    for sym in need:       # a symbol the C library provides links on its own
        with open(probe, "w") as f:
            f.write("extern char %s; void *p_(void) { return &%s; }\n" % (sym, sym))
        r = subprocess.run([GCC, "-w", "-shared", "-fPIC", "-Wl,--no-undefined", probe, "-o", os.path.join(OUT, "probe.so"), "-lm"],
                           capture_output=True)
        if r.returncode != 0:
            stubs.append(sym)
    with open(os.path.join(OUT, "stubs.c"), "w") as f:
        f.write("/* generated: functions of exec_helper.c's environment that the post-processing tests never reach */\n"
                "#include <stdio.h>\n#include <stdlib.h>\n")
        for sym in stubs:
            f.write("void %s(void) { fprintf(stderr, \"oracle/_ref: stub %s called\\n\"); abort(); }\n" % (sym, sym))
    subprocess.check_call([GCC, "-shared", "-fPIC", "-w", "-o", LIB] + objs + [os.path.join(OUT, "stubs.c"), "-lm",
                                                                                 "-Wl,--no-undefined"])
    for f in ("probe.c", "probe.so"):
        if os.path.exists(os.path.join(OUT, f)):
            os.remove(os.path.join(OUT, f))
    return LIB


if __name__ == "__main__":
    print(build(force="-f" in sys.argv))
