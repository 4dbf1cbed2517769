#!/usr/bin/env python3
"""Per-kernel-class device times of one ensemble step (diagnostic; run on the GPU box).
usage: kernel_times.py [nmol] [molecule]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import xtb_b200  # noqa: E402
from bench import load_kat, make_ensemble  # noqa: E402

nmol = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
s = load_kat(sys.argv[2] if len(sys.argv) > 2 else "rivaroxaban")
Z, pos = make_ensemble(s, nmol)
b = xtb_b200.BatchCalculator([Z] * nmol)
b.singlepoint(pos)
b.set_profile("all")
r = b.singlepoint(pos)
prof, passes = b.kernel_times()
t = b.timings()
tot = t["setup_ms"] + t["scc_ms"] + t["grad_ms"]
sw = np.mean([b.counts(i)["sweeps"] for i in range(0, nmol, max(1, nmol // 64))])
print("nmol %d  total %.1f ms (%.0f mol/s)  passes %d  mean Jacobi sweeps/molecule %.1f  launches %d  E0 %.10f" % (
    nmol, tot, nmol / tot * 1e3, passes, sw, t["launches"], r["energies"][0]))
for k, v in sorted(prof.items(), key=lambda x: -x[1]["ms"]):
    if v["launches"]:
        print("  %-42s %9.2f ms %5.1f%%  %4d launches  %8.3f ms/launch" % (k, v["ms"], 100 * v["ms"] / tot, v["launches"],
                                                                            v["ms"] / v["launches"]))
