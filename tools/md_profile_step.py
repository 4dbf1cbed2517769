#!/usr/bin/env python3
"""A few MD steps of the bench's peptide cluster in ALPB water for ncu launch lists.  usage: md_profile_step.py [nsteps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import xtb_b200  # noqa: E402
from bench import load_kat, make_peptide_cluster  # noqa: E402

nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
Z, xyz = make_peptide_cluster(load_kat("pdb_4qxx"), 3, gap=8.0)
b = xtb_b200.BatchCalculator([Z])
b.set_solvent("water", model="alpb")
ncons = b.md_init(xyz, seed=1)
r = b.md_run(nsteps)
print("constraints %d steps %d epot %.8f T %.1f" % (ncons, r["steps"], r["epot"][-1, 0], r["temperature"][-1, 0]))
