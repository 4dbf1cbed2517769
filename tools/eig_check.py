#!/usr/bin/env python3
"""Accuracy and stage timings of the large-matrix eigensolver (xtbb200_symmetricEigen) against LAPACK; run on the GPU box.
usage: eig_check.py [n ...]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import xtb_b200  # noqa: E402


def matrices(n, rng):
    a = rng.standard_normal((n, n))
    yield "random", (a + a.T) / 2
    if n >= 14:
        b = rng.standard_normal((7, 7))
        c = np.kron(np.eye(n // 7 + 1), b + b.T)[:n, :n] + 1e-3 * rng.standard_normal((n, n))
        yield "clustered", (c + c.T) / 2
    yield "diag+eps", np.diag(np.arange(n, dtype=float)) + 1e-10 * (a + a.T)


def main():
    sizes = [int(x) for x in sys.argv[1:]] or [1, 2, 3, 5, 17, 64, 65, 129, 200, 333, 1000, 1472]
    rng = np.random.default_rng(5)
    worst = 0.0
    for n in sizes:
        for name, a in matrices(n, rng):
            if os.environ.get("EIG_ONLY") and name != os.environ["EIG_ONLY"]:
                continue
            t0 = time.time()
            w, z, ms = xtb_b200.symmetric_eigen(a)
            wall = time.time() - t0
            if n <= 3000:
                wr = np.linalg.eigvalsh(a)
                dw = np.max(np.abs(w - wr))
            else:
                dw = float("nan")
            sc = max(np.abs(a).max(), 1e-300) * n * np.finfo(float).eps
            orth = np.max(np.abs(z.T @ z - np.eye(n))) / (n * np.finfo(float).eps)
            res = np.max(np.abs(a @ z - z * w)) / sc
            worst = max(worst, orth, res)
            print("n %5d %-10s dw %.2e  orth/(n eps) %.2f  resid/(n eps |A|) %.2f  sytrd %.2f dc %.2f bt %.2f ms  wall %.2f s" % (
                n, name, dw, orth, res, ms["sytrd_ms"], ms["dc_ms"], ms["backtransform_ms"], wall), flush=True)
    print("worst scaled error %.2f" % worst)
    return 0 if worst < 50 else 1


if __name__ == "__main__":
    sys.exit(main())
