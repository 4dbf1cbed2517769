"""xtb_hessian through the C ABI: numerical Hessian by central differences of analytical gradients (src/api/interface.f90:312-456,
src/type/calculator.f90:116-229), the 6 N displaced geometries evaluated as one ensemble on the GPU.  Checked against the same
finite-difference recipe applied to the CPU oracle's gradients (the reference holds no GFN2 Hessian golden values for the API)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _oracle_hessian(s, step, restart=None):
    from oracle.oracle import Oracle
    x0 = np.asarray(s["xyz"], dtype=float).reshape(-1, 3)
    n3 = x0.size
    H = np.zeros((n3, n3))
    D = np.zeros((n3, 3))
    for ii in range(n3):
        g, d = [], []
        for sgn in (+1, -1):
            x = x0.copy()
            x.reshape(-1)[ii] += sgn * step
            o = Oracle(s["numbers"], x)
            if restart is not None:
                o.set_restart(*restart)
            assert o.singlepoint() == 0
            g.append(o.get("gradient"))
            d.append(o.get("dipole"))
        H[:, ii] = (g[0] - g[1]) / (2 * step)
        D[ii] = (d[0] - d[1]) / (2 * step)
    return 0.5 * (H + H.T), D


def test_hessian_matches_finite_differences_of_the_oracle(kat):
    import xtb_b200
    s = kat["c3h4_api"]
    calc = xtb_b200.Calculator(xtb_b200.Param.GFN2xTB, s["numbers"], s["xyz"])
    calc.set_verbosity(0)
    hess, dipgrad = calc.hessian()                       # fresh results handle: every point starts from the EEQ guess
    Href, Dref = _oracle_hessian(s, 0.005)
    assert np.max(np.abs(hess - hess.T)) == 0.0
    assert np.max(np.abs(hess - Href)) < 1e-6
    assert np.max(np.abs(dipgrad - Dref)) < 1e-6
    # translational invariance: rows sum to zero over the atoms (acoustic sum rule), within the SCC noise
    n = len(s["numbers"])
    assert np.max(np.abs(hess.reshape(3 * n, n, 3).sum(axis=1))) < 1e-4
    w = np.linalg.eigvalsh(hess)
    assert np.sum(np.abs(w) < 1e-8) == 3                 # the three translations (the geometry is not a stationary point)


def test_hessian_restarts_from_the_cached_wavefunction(kat):
    import xtb_b200
    from oracle.oracle import Oracle
    s = kat["c3h4_api"]
    calc = xtb_b200.Calculator(xtb_b200.Param.GFN2xTB, s["numbers"], s["xyz"])
    calc.set_verbosity(0)
    res = calc.singlepoint()
    hess, _ = calc.hessian(res, step=0.01)
    o = Oracle(s["numbers"], s["xyz"])
    assert o.singlepoint() == 0
    Href, _ = _oracle_hessian(s, 0.01, restart=(o.get("qsh"), o.get("dipm"), o.get("qp")))
    assert np.max(np.abs(hess - Href)) < 1e-6
