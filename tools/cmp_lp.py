#!/usr/bin/env python
"""Compare two log-likelihood dumps written by tools/time_sn.py --save (A/B library builds)."""
import sys, torch
a, b = torch.load(sys.argv[1]), torch.load(sys.argv[2])
assert torch.equal(a["err"], b["err"]), "error flags differ"
ok = a["err"] == 0
d = (a["lp"][ok] - b["lp"][ok]).abs() / a["lp"][ok].abs().clamp_min(1.0)
print("n = %d  max rel diff = %.3e  mean = %.3e  n(>1e-12) = %d  n(>1e-10) = %d" %
      (ok.sum().item(), d.max().item(), d.mean().item(), (d > 1e-12).sum().item(), (d > 1e-10).sum().item()))
