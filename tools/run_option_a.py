#!/usr/bin/env python
"""Option A at scale: the reference's UNCHANGED cosmo_pmc binary (build_ref/cosmo_pmc, linked against
libpmc_b200.so) on Demo/MC_Demo/SN with nsamples raised to 10^6 / 10^7, wall clock per pmclib-named entry point
(PMCB200_TIMING=1) beside the total.  usage: python tools/run_option_a.py NSAMPLES [NITER] [--always-upload]"""
import os, re, shutil, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from cosmopmc_b200 import targets as T

n = int(float(sys.argv[1])); niter = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 3
run = tempfile.mkdtemp(prefix="optA_")
demo = os.path.join(ROOT, "build_ref", "demo_SN")
for f in os.listdir(demo):
    if os.path.isfile(os.path.join(demo, f)):
        shutil.copy(os.path.join(demo, f), run)
cfg = open(os.path.join(run, "config_pmc")).read()
cfg = re.sub(r"nsamples\s+\d+", "nsamples        %d" % n, cfg)
cfg = re.sub(r"niter\s+\d+", "niter           %d" % niter, cfg)
cfg = re.sub(r"fsfinal\s+\S+", "fsfinal         1", cfg)
open(os.path.join(run, "config_pmc"), "w").write(cfg)
F = np.linalg.inv(T.SN_POST_COV)
with open(os.path.join(run, "fisher"), "w") as f:
    f.write("5 -1 5 0\n" + " ".join("%.10g" % v for v in T.SN_POST_MEAN) + "\n")
    for r in F:
        f.write(" ".join("%.10g" % v for v in r) + "\n")
env = dict(os.environ, PMCB200_TIMING="1")
if "--always-upload" in sys.argv:
    env["PMCB200_ALWAYS_UPLOAD"] = "1"
t = time.time()
o = subprocess.run([os.path.join(ROOT, "build_ref", "cosmo_pmc"), "-c", "config_pmc", "-s", "1", "-q"], cwd=run,
                   capture_output=True, text=True, env=env)
dt = time.time() - t
print("nsamples %d niter %d always_upload %d: total wall %.2f s, rc %d" % (n, niter, "--always-upload" in sys.argv, dt, o.returncode))
print("\n".join(l for l in o.stderr.split("\n") if "pmcb200 timing" in l))
try:
    print("perplexity:", open(os.path.join(run, "perplexity")).read().strip().replace("\n", " | "))
    sz = os.path.getsize(os.path.join(run, "iter_%d" % (niter - 1), "pmcsim"))
    print("text pmcsim of the last iteration: %.1f MB" % (sz / 1e6))
except Exception as e:
    print("outputs:", e, o.stderr[-500:])
shutil.rmtree(run, ignore_errors=True)
