"""BASELINE configs[3] at full size: the (H2O)1000 cluster (3000 atoms, 6000 AOs) of bench.py, full single point +
gradient on the GPU against the CPU oracle with the north-star tolerances (E 1e-7 Eh, gradient 1e-6 Eh/bohr, charges 1e-6 e),
SCC iteration count and screening list lengths identical.  The oracle needs about a minute on the box's host cores."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

E_TOL, G_TOL, Q_TOL = 1e-7, 1e-6, 1e-6


def _cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


@pytest.mark.timeout(1500)
def test_water1000_full_single_point_matches_oracle():
    import bench
    import xtb_b200
    from oracle.oracle import Oracle
    if xtb_b200.device_count() < 1:
        pytest.fail("no CUDA device: the engine has no CPU fallback")
    Z, xyz = bench.make_water_cluster(1000)
    b = xtb_b200.BatchCalculator([Z])
    r = b.singlepoint(xyz)
    assert r["all_converged"]
    o = Oracle(Z, xyz, threads=_cores())
    assert o.singlepoint() == 0
    res, info, e = b.result(0), o.info, o.energies
    assert b.sizes(0)["nao"] == info["nao"] == 6000
    assert abs(res["energy"] - e["e_total"]) < E_TOL
    assert int(res["niter"]) == info["niter"], "SCC iteration count differs"
    cnt = b.counts(0)
    assert (cnt["ndp"], cnt["nqp"]) == (info["ndp"], info["nqp"])
    assert np.max(np.abs(r["gradients"] - o.get("gradient").reshape(-1, 3))) < G_TOL
    assert np.max(np.abs(r["charges"] - o.get("q"))) < Q_TOL
    for key in ("e_rep", "e_es", "e_aes", "e_axc", "e_disp", "e_atm"):
        assert abs(res[key] - e[key]) < E_TOL, key
    assert abs(res["hl_gap"] - e["hl_gap"]) < 1e-6
    print("\n(H2O)1000: dE = %.2e Eh, max|dg| = %.2e, max|dq| = %.2e, niter %d, ndp %d, nqp %d" % (
        res["energy"] - e["e_total"], np.max(np.abs(r["gradients"] - o.get("gradient").reshape(-1, 3))),
        np.max(np.abs(r["charges"] - o.get("q"))), info["niter"], info["ndp"], info["nqp"]))


@pytest.mark.timeout(900)
def test_polypeptide_full_single_point_matches_oracle():
    """BASELINE configs[1] stand-in: the covalently connected 648-atom polypeptide with crambin's sequence
    (xtb_b200/polypeptide.py, 1659 AOs): E, gradient, charges, SCC iteration count, screening list lengths."""
    import xtb_b200
    from oracle.oracle import Oracle
    from xtb_b200.polypeptide import make_polypeptide
    if xtb_b200.device_count() < 1:
        pytest.fail("no CUDA device: the engine has no CPU fallback")
    Z, xyz, info = make_polypeptide()
    assert info["formula"] == "C202H321N55O64S6" and len(Z) == 648
    b = xtb_b200.BatchCalculator([Z])
    r = b.singlepoint(xyz)
    assert r["all_converged"]
    o = Oracle(Z, xyz, threads=_cores())
    assert o.singlepoint() == 0
    res, oi, e = b.result(0), o.info, o.energies
    assert b.sizes(0)["nao"] == oi["nao"] == 1659
    assert abs(res["energy"] - e["e_total"]) < E_TOL
    assert int(res["niter"]) == oi["niter"], "SCC iteration count differs"
    cnt = b.counts(0)
    assert (cnt["ndp"], cnt["nqp"]) == (oi["ndp"], oi["nqp"])
    dg = np.max(np.abs(r["gradients"] - o.get("gradient").reshape(-1, 3)))
    dq = np.max(np.abs(r["charges"] - o.get("q")))
    assert dg < G_TOL and dq < Q_TOL
    print("\npolypeptide: dE = %.2e Eh, max|dg| = %.2e, max|dq| = %.2e, niter %d" % (res["energy"] - e["e_total"], dg, dq, oi["niter"]))
