#!/usr/bin/env python3
"""Stage-by-stage comparison of the CUDA engine against the CPU oracle (diagnostic, run on the GPU box)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import xtb_b200  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402

d = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_kat.json")))
names = sys.argv[1:] or ["h2o_scc", "c3h4_api", "caffeine", "mindless01", "mindless02", "feco5", "fe_cnx6", "rivaroxaban"]


def mx(a, b):
    a = np.asarray(a); b = np.asarray(b)
    return float(np.max(np.abs(a - b))) if a.size else 0.0


for name in names:
    s = d[name]
    Z = np.array(s["numbers"], dtype=np.int32)
    xyz = np.array(s["xyz"]).reshape(-1, 3)
    guess = s.get("guess", "eeq")
    o = Oracle(Z, xyz, charge=s["charge"], uhf=s["uhf"], guess=guess, maxiter=s.get("maxiter", 250), threads=1)
    t0 = time.time(); st = o.singlepoint(); tc = time.time() - t0
    b = xtb_b200.BatchCalculator([Z], charges=[s["charge"]], uhf=[s["uhf"]], guess=guess, maxiter=s.get("maxiter", 250))
    t0 = time.time(); r = b.singlepoint(xyz); tg = time.time() - t0
    res = b.result(0)
    cnt = b.counts(0)
    info = o.info; e = o.energies
    n = len(Z); nao = info["nao"]
    print("==== %s  n=%d nao=%d  oracle st=%d it=%d (%.3fs) | gpu st=%s it=%d sweeps=%d (%.3fs) %s" % (
        name, n, nao, st, info["niter"], tc, r["status"], res["niter"], cnt["sweeps"], tg, b.timings()))
    print("  cn %.2e  radcn %.2e selfE %.2e gw %.2e" % (mx(o.get("cn"), b.array("cn")), mx(o.get("radcn"), b.array("radcn")),
          mx(o.get("selfenergy"), b.array("selfenergy")), mx(o.get("gw"), b.array("gw"))))
    S = o.get("S").reshape(nao, nao); Dp = o.get("dpint").reshape(nao, nao, 3); Qp = o.get("qpint").reshape(nao, nao, 6)
    gD = b.matrix("dpint", 0, 3); gQ = b.matrix("qpint", 0, 6)
    print("  S %.2e  D %.2e  Q %.2e" % (mx(S, b.matrix("S", 0)), mx(Dp.transpose(2, 0, 1), gD.transpose(0, 2, 1)),
                                        mx(Qp.transpose(2, 0, 1), gQ.transpose(0, 2, 1))))
    H0p = o.get("H0"); H0 = np.zeros((nao, nao)); iu = np.tril_indices(nao); H0[iu] = H0p; H0 = H0 + H0.T - np.diag(np.diag(H0))
    print("  H0 %.2e  jmat %.2e gab3 %.2e gab5 %.2e dispmat %.2e" % (mx(H0, b.matrix("H0", 0)), mx(o.get("jmat"), b.array("jmat")),
          mx(o.get("gab3"), b.array("gab3")), mx(o.get("gab5"), b.array("gab5")), mx(o.get("dispmat"), b.array("dispmat"))))
    print("  ndp %d/%d nqp %d/%d" % (info["ndp"], cnt["ndp"], info["nqp"], cnt["nqp"]))
    print("  E %.3e (%.10f)  eel-parts: es %.2e aes %.2e axc %.2e disp %.2e rep %.2e gap %.2e atm %.2e" % (
        res["energy"] - e["e_total"], res["energy"], res["e_es"] - e["e_es"], res["e_aes"] - e["e_aes"], res["e_axc"] - e["e_axc"],
        res["e_disp"] - e["e_disp"], res["e_rep"] - e["e_rep"], res["hl_gap"] - e["hl_gap"], res["e_atm"] - e["e_atm"]))
    print("  q %.2e qsh %.2e dipm %.2e qp %.2e  P %.2e emo %.2e" % (mx(o.get("q"), r["charges"]), mx(o.get("qsh"), b.array("qsh")),
          mx(o.get("dipm"), b.array("dipm")), mx(o.get("qp"), b.array("qp")), mx(o.get("P").reshape(nao, nao), b.matrix("P", 0)),
          mx(o.get("emo"), b.array("emo"))))
    g = o.get("gradient").reshape(-1, 3)
    print("  grad %.2e (gnorm %.3e)  wbo %.2e  dipole %.2e" % (mx(g, r["gradients"]), res["gnorm"] - e["gnorm"],
          mx(o.get("wbo"), b.array("wbo")), mx(o.get("dipole"), [res["dipole_x"], res["dipole_y"], res["dipole_z"]])))
    if "kat" in s:
        print("  vs reference KAT: dE %.2e" % (res["energy"] - s["kat"]["energy"]))
    del b
