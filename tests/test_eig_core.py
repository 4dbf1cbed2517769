"""CPU checks of the scalar routines of the divide-and-conquer tridiagonal eigensolver (xtb_b200/csrc/eig_dc_core.h)
through the host harness tests/eig_core_host.cpp, which walks the tearing tree exactly like the CUDA kernels of
xtb_b200/csrc/kernels_eigh.cuh do.  Reference: LAPACK dstevd through scipy."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
from scipy.linalg import eigh_tridiagonal

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("eigcore") / "libeig_core_host.so")
    subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-o", out,
                           os.path.join(ROOT, "tests", "eig_core_host.cpp")])
    lib = ctypes.CDLL(out)
    lib.dc_tridiag_host.restype = ctypes.c_int
    return lib


def run(lib, d, e):
    n = len(d)
    lam = np.zeros(n)
    Q = np.zeros((n, n), order="F")
    stats = np.zeros(4, dtype=np.int32)
    dp = ctypes.POINTER(ctypes.c_double)
    e_ = np.ascontiguousarray(np.append(e, 0.0))
    rc = lib.dc_tridiag_host(n, np.ascontiguousarray(d).ctypes.data_as(dp), e_.ctypes.data_as(dp), lam.ctypes.data_as(dp),
                             Q.ctypes.data_as(dp), stats.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
    return rc, lam, Q, stats


def cases():
    rng = np.random.default_rng(7)
    out = []
    for n in (1, 2, 3, 5, 8, 17, 64, 100, 257):
        out.append(("random%d" % n, rng.standard_normal(n), rng.standard_normal(max(n - 1, 0))))
    n = 101
    out.append(("wilkinson", np.abs(np.arange(n) - n // 2).astype(float), np.ones(n - 1)))
    out.append(("toeplitz", 2 * np.ones(200), -np.ones(199)))
    out.append(("zero_offdiag", rng.standard_normal(50), np.zeros(49)))
    e = rng.standard_normal(149)
    e[::7] = 1e-14
    out.append(("nearly_split", rng.standard_normal(150), e))
    out.append(("clustered", 1.0 + 1e-9 * rng.standard_normal(120), 1e-7 * rng.standard_normal(119)))
    out.append(("graded", 10.0 ** (-np.arange(60) / 4.0), 10.0 ** (-np.arange(59) / 4.0 - 1)))
    out.append(("identical_blocks", np.tile([1.0, -2.0, 0.5, 3.0], 40), np.tile([0.3, 0.7, -0.2, 1e-6], 40)[:-1]))
    out.append(("negative_e", rng.standard_normal(90), -np.abs(rng.standard_normal(89))))
    return out


@pytest.mark.parametrize("name,d,e", cases(), ids=[c[0] for c in cases()])
def test_dc_tridiagonal_host(harness, name, d, e):
    rc, lam, Q, stats = run(harness, d, e)
    n = len(d)
    assert rc == 0, "secular equation did not converge"
    T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1) if n > 1 else np.diag(d)
    ref = eigh_tridiagonal(d, e, eigvals_only=True) if n > 1 else np.array(d)
    scale = max(np.abs(T).max(), 1e-300)
    assert np.max(np.abs(lam - ref)) <= 50 * n * np.finfo(float).eps * scale
    assert np.max(np.abs(Q.T @ Q - np.eye(n))) <= 50 * n * np.finfo(float).eps
    assert np.max(np.abs(T @ Q - Q * lam)) <= 50 * n * np.finfo(float).eps * scale
