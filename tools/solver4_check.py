#!/usr/bin/env python3
"""Diagnostic: batched Householder + divide-and-conquer solver (mode 4) against the block Jacobi (mode 0) on a small ensemble."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import xtb_b200
from bench import load_kat, make_ensemble
nmol = int(sys.argv[1]) if len(sys.argv) > 1 else 64
s = load_kat(sys.argv[2] if len(sys.argv) > 2 else "rivaroxaban")
Z, pos = make_ensemble(s, nmol)
out = {}
for mode in (0, 4):
    b = xtb_b200.BatchCalculator([Z] * nmol)
    b.set_solver(mode)
    b.set_profile("all")
    r = b.singlepoint(pos)
    t0 = time.time(); r = b.singlepoint(pos); wall = time.time() - t0
    t = b.timings()
    prof, passes = b.kernel_times()
    out[mode] = r
    print("mode", mode, "E0 %.10f" % r["energies"][0], "niter", r["niter"][:6], "conv", r["all_converged"], "device ms %.1f" % (t["setup_ms"] + t["scc_ms"] + t["grad_ms"]),
          "eig ms %.1f" % (prof["k_jacobi"]["ms"] / 2), "launches", t["launches"])
print("max |dE|", np.abs(out[0]["energies"] - out[4]["energies"]).max(), "max |dG|", np.abs(out[0]["gradients"] - out[4]["gradients"]).max(),
      "niter equal", np.array_equal(out[0]["niter"], out[4]["niter"]))
