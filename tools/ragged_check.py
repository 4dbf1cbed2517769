#!/usr/bin/env python3
"""Diagnostic: ragged batch of molecules beyond the shared-memory solvers, every solver mode against single-molecule runs."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import xtb_b200
kat = json.load(open(os.path.join(ROOT, "tests/golden/reference_kat.json")))
names = sys.argv[1:] or ["taxol", "caffeine", "remdesivir"]
mols = [(np.array(kat[n]["numbers"], dtype=np.int32), np.array(kat[n]["xyz"]).reshape(-1, 3)) for n in names]
single = []
for Z, x in mols:
    b = xtb_b200.BatchCalculator([Z])
    r = b.singlepoint(x)
    single.append((r["energies"][0], int(r["niter"][0]), b.sizes(0)["nao"]))
print("single:", single)
for mode in (0, 2, 3):
    b = xtb_b200.BatchCalculator([Z for Z, _ in mols])
    b.set_solver(mode)
    r = b.singlepoint(np.concatenate([x for _, x in mols]))
    print("batch solver", mode, [float(e) for e in r["energies"]], [int(i) for i in r["niter"]], [int(s) for s in r["status"]])
    if mode == 0:
        X = b.matrix("X", 0); S = b.matrix("S", 0)
        print("   |X^T S X - 1| mol0:", np.abs(X.T @ S @ X - np.eye(len(S))).max(), " mol1:", np.abs(b.matrix("X", 1).T @ b.matrix("S", 1) @ b.matrix("X", 1) - np.eye(len(b.matrix("S", 1)))).max(),
              " mol2:", np.abs(b.matrix("X", 2).T @ b.matrix("S", 2) @ b.matrix("X", 2) - np.eye(len(b.matrix("S", 2)))).max())
