#!/usr/bin/env python3
"""One ensemble single point + gradient (the bench workload, smaller batch) for ncu captures.
usage: profile_step.py [nmol] [molecule] [ncalls]; SOLVENT=<name> in the environment: ALPB implicit solvation"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import xtb_b200  # noqa: E402
from bench import load_kat, make_ensemble  # noqa: E402

nmol = int(sys.argv[1]) if len(sys.argv) > 1 else 592
s = load_kat(sys.argv[2] if len(sys.argv) > 2 else "rivaroxaban")
ncalls = int(sys.argv[3]) if len(sys.argv) > 3 else 1
Z, pos = make_ensemble(s, nmol)
b = xtb_b200.BatchCalculator([Z] * nmol)
if os.environ.get("SOLVENT"):
    b.set_solvent(os.environ["SOLVENT"], model="alpb")
for _ in range(ncalls):
    r = b.singlepoint(pos)
print("E0 %.10f niter %d launches %d" % (r["energies"][0], r["niter"][0], b.timings()["launches"]))
