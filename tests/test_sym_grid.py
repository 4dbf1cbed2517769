"""CPU check of the tile grid of the symmetric trailing-matrix product of the tridiagonalisation kernel
(xtb_b200/csrc/sytrd_sym_grid.h, used by k_sytrd_panel in kernels_eigh.cuh): tests/sym_grid_host.cpp walks the tiles with the
kernel's own index arithmetic; the gathered y must equal A v of the trailing matrix, every slot must be written once and
stay inside the buffers the launch code allocates."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("symgrid") / "libsym_grid_host.so")
    subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-o", out,
                           os.path.join(ROOT, "tests", "sym_grid_host.cpp")])
    lib = ctypes.CDLL(out)
    lib.sym_product_host.restype = ctypes.c_int
    return lib


CASES = [(130, 1, 148), (130, 129, 148), (258, 33, 7), (640, 1, 148), (640, 127, 148), (640, 128, 148), (640, 129, 148),
         (1000, 500, 148), (1922, 1, 148), (1922, 63, 148), (2000, 999, 3), (2000, 1999, 148), (2500, 641, 148), (3000, 1, 132),
         (3000, 1, 1), (96, 5, 148), (34, 33, 148)]


@pytest.mark.parametrize("n,jj,G", CASES)
def test_symmetric_product_from_tile_sums(harness, n, jj, G):
    rng = np.random.default_rng(n + jj)
    a = rng.standard_normal((n, n))
    a = np.asfortranarray((a + a.T) / 2)
    v = rng.standard_normal(n)
    ref = a[jj:, jj:] @ v[jj:]
    low = np.asfortranarray(np.tril(a) + np.triu(np.full((n, n), 1e30), 1))   # the upper triangle must never be used
    y = np.zeros(n)
    stats = np.zeros(6, dtype=np.int64)
    dp = ctypes.POINTER(ctypes.c_double)
    rc = harness.sym_product_host(n, jj, G, low.ctypes.data_as(dp), v.ctypes.data_as(dp), y.ctypes.data_as(dp),
                                  stats.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)))
    nS, nR = (n + 31) // 32, (n + 127) // 128
    if rc == -7:   # more terms per row than a slice of threads holds: only with fewer tiles than CTAs, where the launch code
        assert n < 1900 and G > 100   # (EighWork::solve, sym_gather_max_terms) keeps the two-triangle form
        return
    assert rc == 0
    assert np.max(np.abs(y[jj:] - ref)) < 1e-11 * max(1.0, np.abs(ref).max())
    assert stats[1] <= nR + G + 1 and stats[2] <= nS * nR      # buffer sizes of EighWork::init
    assert stats[4] <= (stats[0] + G - 1) // G                  # balanced ranges


def test_terms_per_row_fit_the_slices_at_production_orders(harness):
    # every order the panel kernel can meet on 148 CTAs with the symmetric form: 16 slices x 6 terms
    dp = ctypes.POINTER(ctypes.c_double)
    for n in (1924, 3000, 6000, 9000):
        stats = np.zeros(6, dtype=np.int64)
        a = np.zeros((1, 1))
        # jj = n - 1: a one-element trailing matrix keeps the walk cheap; stats[5] covers all trailing orders of n
        low = np.asfortranarray(np.zeros((n, n)))
        v = np.zeros(n)
        y = np.zeros(n)
        rc = harness.sym_product_host(n, n - 1, 148, low.ctypes.data_as(dp), v.ctypes.data_as(dp), y.ctypes.data_as(dp),
                                      stats.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)))
        assert rc in (0, -7)
        assert stats[5] <= 16 * 6, (n, stats[5])
