"""GPU checks of the large-matrix symmetric eigensolver (Householder tridiagonalisation + divide and conquer +
back-transformation, xtb_b200/csrc/kernels_eigh.cuh) through the C ABI (xtbb200_symmetricEigen), against LAPACK's
dsyevd -- the routine the reference calls at src/mctc/lapack/eigensolve.F90:240.  Tolerances in units of n*eps*|A|."""
import numpy as np
import pytest

import xtb_b200

pytestmark = pytest.mark.gpu
EPS = np.finfo(float).eps


def _check(a, tol=30.0):
    n = a.shape[0]
    w, z, _ = xtb_b200.symmetric_eigen(a)
    scale = max(np.abs(a).max(), 1e-300) * n * EPS
    assert np.all(np.diff(w) >= 0)
    assert np.max(np.abs(w - np.linalg.eigvalsh(a))) <= tol * scale
    assert np.max(np.abs(z.T @ z - np.eye(n))) <= tol * n * EPS
    assert np.max(np.abs(a @ z - z * w)) <= tol * scale


@pytest.mark.parametrize("n", [1, 2, 3, 7, 63, 64, 65, 127, 128, 129, 200, 515])
def test_random_symmetric(n):
    rng = np.random.default_rng(n)
    a = rng.standard_normal((n, n))
    _check((a + a.T) / 2)


def test_clustered_spectrum_of_weakly_coupled_identical_blocks():
    """The spectrum of a cluster of identical molecules: bands of nearly degenerate levels (heavy deflation)."""
    rng = np.random.default_rng(11)
    b = rng.standard_normal((6, 6))
    a = np.kron(np.eye(80), b + b.T) + 1e-4 * rng.standard_normal((480, 480))
    _check((a + a.T) / 2)


def test_exactly_degenerate_and_diagonal():
    _check(np.eye(150))
    _check(np.diag(np.repeat([1.0, -2.0, 3.0], 60)))
    rng = np.random.default_rng(3)
    a = rng.standard_normal((170, 170))
    _check(np.diag(np.arange(170.0)) + 1e-11 * (a + a.T))


def test_graded_matrix():
    rng = np.random.default_rng(4)
    n = 220
    s = 10.0 ** (-6 * rng.random(n))
    a = rng.standard_normal((n, n))
    _check(s[:, None] * (a + a.T) * s[None, :])


def test_larger_than_one_l2_slice():
    rng = np.random.default_rng(5)
    a = rng.standard_normal((1500, 1500))
    _check((a + a.T) / 2)


def test_panel_kernel_hands_over_to_the_resident_kernel():
    """n = 2100: the first panels run through k_sytrd_panel, the trailing matrix then fits into shared memory."""
    rng = np.random.default_rng(6)
    a = rng.standard_normal((2100, 2100))
    _check((a + a.T) / 2)


@pytest.mark.parametrize("n", [1926, 2101, 2432, 3006])
def test_both_forms_of_the_panel_kernel(n):
    """Even orders take the symmetric trailing-matrix product (lower triangle through the tensor-map TMA ring, tile sums
    gathered in the next column update), odd orders the two-triangle form (a tile line must start 16-byte aligned);
    1926: a single panel before the resident kernel; 2432 = 19 * 128: whole row chunks; 3006: ragged last chunk and strip."""
    rng = np.random.default_rng(n)
    a = rng.standard_normal((n, n))
    if n == 2432:   # bands of nearly degenerate levels on top of the random matrix
        a = a * 1e-3 + np.kron(np.eye(n // 8), rng.standard_normal((8, 8)))
    _check((a + a.T) / 2)


def test_order_beyond_the_shared_memory_deflation_scan():
    """n = 9700 (deflation scan of the top levels in global memory): matrix with a known spectrum, A = H3 H2 H1 D H1 H2 H3
    with three Householder reflectors, so that no LAPACK call of this size is needed; sampled residuals."""
    n = 9700
    rng = np.random.default_rng(8)
    dvals = np.sort(rng.standard_normal(n) * 3.0)
    dvals[100:140] = dvals[100]                      # a cluster of exactly degenerate levels
    a = np.diag(dvals)
    for _ in range(3):
        v = rng.standard_normal(n)
        v /= np.linalg.norm(v)
        a -= 2.0 * np.outer(v, v @ a)
        a -= 2.0 * np.outer(a @ v, v)
    a = (a + a.T) / 2
    w, z, ms = xtb_b200.symmetric_eigen(a)
    scale = np.abs(a).max() * n * EPS
    assert np.max(np.abs(w - np.sort(dvals))) <= 30 * scale
    cols = rng.choice(n, 64, replace=False)
    zs = z[:, cols]
    assert np.max(np.abs(a @ zs - zs * w[cols])) <= 30 * scale
    assert np.max(np.abs(zs.T @ z - np.eye(n)[cols])) <= 30 * n * EPS
