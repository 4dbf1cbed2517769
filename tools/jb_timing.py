#!/usr/bin/env python3
"""Phase timing of the block Jacobi kernel (diagnostic): run with XTBB200_LIB pointing at the `make timing` build; the kernel
prints the clock64 counters of two warps of molecule 0 per launch.  python tools/jb_timing.py [nmol]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench  # noqa: E402
import xtb_b200  # noqa: E402

nmol = int(sys.argv[1]) if len(sys.argv) > 1 else 148
Z, pos = bench.make_ensemble(bench.load_kat("rivaroxaban"), nmol)
b = xtb_b200.BatchCalculator([Z] * nmol)
r = b.singlepoint(pos)
print("converged", r["all_converged"], "E0", r["energies"][0], "niter", r["niter"][:4])
t = b.timings()
print({k: t[k] for k in t if "ms" in k or k == "launches"})
