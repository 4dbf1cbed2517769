#!/usr/bin/env python3
"""Diagnostic of tests/test_gpu_hessian.py: the 6 N displaced geometries of C3H4 as one batch on the GPU against the oracle point by
point (iteration count, energy, dipole), and the finite-difference dipole gradient of both."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import xtb_b200  # noqa: E402
from oracle.oracle import Oracle, build  # noqa: E402

build()
kat = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_kat.json")))
s = kat["c3h4_api"]
Z = np.asarray(s["numbers"], dtype=np.int32)
x0 = np.asarray(s["xyz"], dtype=float).reshape(-1, 3)
n3 = x0.size
step = 0.005
geos = []
for ii in range(n3):
    for sgn in (+1, -1):
        x = x0.copy()
        x.reshape(-1)[ii] += sgn * step
        geos.append(x)
b = xtb_b200.BatchCalculator([Z] * len(geos))
r = b.singlepoint(np.concatenate(geos))
worst = 0
for k, x in enumerate(geos):
    o = Oracle(Z, x)
    assert o.singlepoint() == 0
    res = b.result(k)
    mu = np.array([res["dipole_x"], res["dipole_y"], res["dipole_z"]])
    dmu = np.max(np.abs(mu - o.get("dipole")))
    de = r["energies"][k] - o.energies["total"] if "total" in o.energies else float("nan")
    nit = o.info["niter"]
    flag = "" if int(r["niter"][k]) == int(nit) else "  <-- iteration count differs"
    if dmu > 2e-9 or flag:
        print("point %2d niter gpu %d oracle %d dE %.2e dmu %.2e%s" % (k, r["niter"][k], nit, de, dmu, flag))
    worst = max(worst, dmu)
print("worst dipole difference %.3e -> dipole gradient error up to %.2e" % (worst, worst / step))
