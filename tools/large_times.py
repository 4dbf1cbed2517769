#!/usr/bin/env python3
"""Per-kernel-class device times of one large single system (diagnostic; run on the GPU box).
usage: large_times.py water <nwat> | peptide <ncopy>   [oracle]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import xtb_b200  # noqa: E402
from bench import load_kat, make_peptide_cluster, make_water_cluster  # noqa: E402

kind, k = sys.argv[1], int(sys.argv[2])
Z, xyz = make_water_cluster(k) if kind == "water" else make_peptide_cluster(load_kat("pdb_4qxx"), k)
b = xtb_b200.BatchCalculator([Z])
b.set_profile("all")
t0 = time.time()
r = b.singlepoint(xyz)
wall = time.time() - t0
prof, passes = b.kernel_times()
t = b.timings()
tot = t["setup_ms"] + t["scc_ms"] + t["grad_ms"]
sz = b.sizes(0)
print("%s %d: natoms %d nao %d  device %.1f ms wall %.2f s  passes %d  status %s niter %d  sweeps %d launches %d  E %.10f  s/SCC-iter %.4f" % (
    kind, k, sz["natoms"], sz["nao"], tot, wall, passes, r["status"][0], r["niter"][0], b.counts(0)["sweeps"], t["launches"],
    r["energies"][0], t["scc_ms"] / 1e3 / max(passes, 1)))
for kk, v in sorted(prof.items(), key=lambda x: -x[1]["ms"]):
    if v["launches"] and v["ms"] / tot > 0.002:
        print("  %-42s %10.2f ms %5.1f%%  %6d launches" % (kk, v["ms"], 100 * v["ms"] / tot, v["launches"]))
if len(sys.argv) > 3:
    from oracle.oracle import Oracle
    t0 = time.time()
    o = Oracle(Z, xyz, threads=os.cpu_count())
    st = o.singlepoint()
    print("oracle: %.1f s status %d niter %d  dE %.2e  dG %.2e dq %.2e" % (
        time.time() - t0, st, o.info["niter"], r["energies"][0] - o.energies["e_total"],
        np.max(np.abs(r["gradients"] - o.get("gradient").reshape(-1, 3))), np.max(np.abs(r["charges"] - o.get("q")))))
