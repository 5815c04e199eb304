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
    link = ["-L", libdir, "-lpmc_b200", "-Wl,-rpath,$ORIGIN/../cosmopmc_b200", "-lm"]
    subprocess.check_call([GCC, "-o", exe] + objs + link)
    # the two executables that precede cosmo_pmc in bin/cosmo_pmc.pl:103-116 (maximum + Fisher matrix);
    # their scalar likelihood calls run the same CUDA kernels with N = 1
    common = [o for o in objs if not o.endswith("cosmo_pmc.o")]
    for main_src, extra in (("exec/max_post.c", ["exec/mkmax.c"]), ("exec/go_fishing.c", ["exec/mkmax.c"])):
        eo = []
        for s_ in [main_src] + extra:
            o = os.path.join(OUT, "obj", os.path.basename(s_).replace(".c", ".o"))
            subprocess.check_call([GCC] + CFLAGS + INC + ["-c", os.path.join(REF, s_), "-o", o])
            eo.append(o)
        subprocess.check_call([GCC, "-o", os.path.join(OUT, os.path.basename(main_src)[:-2])] + eo + common + link)
    # the post-processing executables bin/cosmo_pmc.pl:150-190 runs on the pmcsim / proposal files
    # (host-side file tools; importance_sample re-weights through the scalar posterior = N=1 launches)
    for name in ("importance_sample", "meanvar_sample", "histograms_sample", "add_pmc_proposal",
                 "meanvar_mixmvdens", "cl_one_sided"):
        src = os.path.join(REF, "exec", name + ".c")
        if not os.path.exists(src):
            continue
        o = os.path.join(OUT, "obj", name + ".o")
        subprocess.check_call([GCC] + CFLAGS + INC + ["-c", src, "-o", o])
        subprocess.check_call([GCC, "-o", os.path.join(OUT, name), o] + common + link)
    # this repo's batched counterpart of importance_sample (one device launch instead of one host callback per point)
    o = os.path.join(OUT, "obj", "importance_sample_b200.o")
    subprocess.check_call([GCC] + CFLAGS + INC + ["-c", os.path.join(ROOT, "cosmopmc_b200/exec/importance_sample_b200.c"), "-o", o])
    subprocess.check_call([GCC, "-o", os.path.join(OUT, "importance_sample_b200"), o] + common + link)
    # the SN demo's inputs (Demo/MC_Demo/SN): config + data + parameter files, as bin/cosmo_pmc.pl stages them
    demo = os.path.join(OUT, "demo_SN")
    os.makedirs(demo, exist_ok=True)
    for f in ["Demo/MC_Demo/SN/config_pmc", "data/Sn/Union/sne_union_marek.list", "par_files/cosmo_SN.par",
              "par_files/cosmo.par"]:
        shutil.copy(os.path.join(REF, f), demo)
    # the full pipeline of bin/cosmo_pmc.pl:103-137 (max_post -> go_fishing -> cosmo_pmc) needs the
    # reference's config converter at run time; stage it next to the binaries
    pipe = os.path.join(OUT, "demo_SN_pipeline")
    os.makedirs(pipe, exist_ok=True)
    for f in os.listdir(demo):
        if os.path.isfile(os.path.join(demo, f)) and f not in ("fisher",):
            shutil.copy(os.path.join(demo, f), pipe)
    shutil.copy(os.path.join(REF, "bin/config_pmc_to_max_and_fish.pl"), OUT)
    with open(os.path.join(pipe, "config_max"), "w") as fo:
        subprocess.check_call(["perl", os.path.join(REF, "bin/config_pmc_to_max_and_fish.pl"), "-M", "-f",
                               "0.3 -1.0 19.3 1.5 -2.0", "-c", os.path.join(pipe, "config_pmc")], stdout=fo)
    # two more demos through the whole pipeline: WMAP distance priors (CMBDistPrior, nclipw 5, revive)
    # and BAO d_z (Demo/MC_Demo/{WMAP_Distance_Priors,BAO/distance_d_z})
    for name, cfgdir, files, fid in (
            ("demo_WMAP_DP", "Demo/MC_Demo/WMAP_Distance_Priors",
             ["data/WMAP_Distance_Priors/wmap7DistPrior_ML_covinv", "par_files/cosmoDP.par"], "0.045 0.27 0.73 0.71"),
            ("demo_BAO_dz", "Demo/MC_Demo/BAO/distance_d_z",
             ["data/BAO/bao_BOSS12_d_z_0.57", "par_files/cosmoDP.par"], "0.27 0.73")):
        dst = os.path.join(OUT, name)
        os.makedirs(dst, exist_ok=True)
        shutil.copy(os.path.join(REF, cfgdir, "config_pmc"), dst)
        for f in files:
            shutil.copy(os.path.join(REF, f), dst)
        with open(os.path.join(dst, "config_max"), "w") as fo:
            subprocess.check_call(["perl", os.path.join(REF, "bin/config_pmc_to_max_and_fish.pl"), "-M", "-f", fid,
                                   "-c", os.path.join(dst, "config_pmc")], stdout=fo)
    # the reference's own regression recipe, bin/test_suite_cosmo_pmc.pl:51-58,321-367,431-452: `max_post -m n`
    # (no maximisation: the log-posterior at a fixed fiducial point) for the in-scope demo directories
    ts = os.path.join(OUT, "test_suite")
    SUITE = (("SN", "Demo/MC_Demo/SN", ["data/Sn/Union/sne_union_marek.list", "par_files/cosmo_SN.par", "par_files/cosmo.par"],
              "0.27 -1.0 19.31 1.6 -1.8"),
             ("BAO_distance_A", "Demo/MC_Demo/BAO/distance_A", ["data/BAO/bao_Reid10_A_0.35", "par_files/cosmoDP.par"], "0.27 0.73"),
             ("BAO_distance_d_z", "Demo/MC_Demo/BAO/distance_d_z", ["data/BAO/bao_BOSS12_d_z_0.57", "par_files/cosmoDP.par"], "0.27 0.73"),
             ("WMAP_Distance_Priors", "Demo/MC_Demo/WMAP_Distance_Priors",
              ["data/WMAP_Distance_Priors/wmap7DistPrior_ML_covinv", "par_files/cosmoDP.par"], "0.045 0.27 0.73 0.71"))
    for name, cfgdir, files, fid in SUITE:
        dst = os.path.join(ts, name)
        os.makedirs(dst, exist_ok=True)
        shutil.copy(os.path.join(REF, cfgdir, "config_pmc"), dst)
        for f in files:
            shutil.copy(os.path.join(REF, f), dst)
        with open(os.path.join(dst, "config_max_test_suite"), "w") as fo:
            subprocess.check_call(["perl", os.path.join(REF, "bin/config_pmc_to_max_and_fish.pl"), "-M", "-f", fid,
                                   "-c", os.path.join(dst, "config_pmc")], stdout=fo)
    # the joint set of the test suite (COSMOS-S10+SN+BAO, fid :56) without its out-of-scope lensing probe:
    # the SN and BAO sections of that config file, the lensing-only parameters (sigma_8, z_rescale) dropped
    dst = os.path.join(ts, "SN+BAO")
    os.makedirs(dst, exist_ok=True)
    for f in ["data/Sn/Union/sne_union_marek.list", "par_files/cosmo_SN.par", "par_files/cosmo.par",
              "data/BAO/bao_Reid10_A_0.35", "par_files/cosmoDP.par"]:
        shutil.copy(os.path.join(REF, f), dst)
    src = open(os.path.join(REF, "Demo/MC_Demo/COSMOS-S10+SN+BAO/config_pmc")).read().split("\n")
    keep, skip = [], False
    for line in src:
        t = line.split()
        if t[:1] == ["#"] and t[1:2] == ["Lensing"]:
            skip = True
        elif t[:1] == ["#"] and t[1:2] == ["BAO"]:
            skip = False
        if skip or t[:2] == ["sdata", "Lensing"]:
            continue
        if t[:1] == ["npar"]: line = "npar            6"
        if t[:1] == ["spar"]: line = "spar            Omega_m w_0_de h_100 M     alpha  beta"
        if t[:1] == ["min"]: line = "min             0.0   -3.5     0.4   19.1  0.5   -3.5"
        if t[:1] == ["max"]: line = "max             1.2    0.5     1.0   19.8  2.6   -0.8"
        if t[:1] == ["ndata"]: line = "ndata           2"
        keep.append(line)
    open(os.path.join(dst, "config_pmc"), "w").write("\n".join(keep))
    with open(os.path.join(dst, "config_max_test_suite"), "w") as fo:
        subprocess.check_call(["perl", os.path.join(REF, "bin/config_pmc_to_max_and_fish.pl"), "-M", "-f",
                               "0.27 -1.0 0.7 19.31 1.6 -1.8", "-c", os.path.join(dst, "config_pmc")], stdout=fo)
    # the tempering demos (Demo/tempering/README.md: evidence known answers)
    for sub in ["1_mvnorm_2D_temp_none", "2_mixmvnorm_2D_temp_none"]:
        dst = os.path.join(OUT, "demo_" + sub)
        os.makedirs(dst, exist_ok=True)
        shutil.copy(os.path.join(REF, "Demo/tempering", sub, "config_pmc"), dst)
    return exe


if __name__ == "__main__":
    print(build())
