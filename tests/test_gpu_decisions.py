"""Discrete decisions of the path compared bit-exactly with the CPU oracle: surviving primitive
pairs of the integral and gradient kernels, atom pairs inside every cutoff, non-zero entries of the D4 self-consistent matrix,
ATM triples, bond-order pairs.  The engine counts them where its kernels take the decisions (xtbb200_setDecisionCounting),
the oracle where its loops do (the reference's own loop structure)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _compare(xb, Z, xyz, charge=0.0, uhf=0):
    from oracle.oracle import Oracle
    b = xb.BatchCalculator([Z], charges=[charge], uhf=[uhf])
    b.set_decision_counting(True)
    r = b.singlepoint(xyz)
    assert r["all_converged"]
    got = b.decision_counts(0)
    o = Oracle(np.asarray(Z, dtype=np.int32), np.asarray(xyz, dtype=np.float64).reshape(-1, 3), charge=charge, uhf=uhf)
    assert o.singlepoint() == 0
    ref = dict(zip(xb.BatchCalculator.DECISION_COUNTS, o.get("decisions").astype(np.int64).tolist()))
    assert got == ref
    cnt = b.counts(0)
    assert (cnt["ndp"], cnt["nqp"]) == (o.info["ndp"], o.info["nqp"])
    # a second call counts afresh (the counters are per single point), and switching the counting off leaves results alone
    e0 = r["energies"][0]
    b.singlepoint(xyz)
    assert b.decision_counts(0) == ref
    b.set_decision_counting(False)
    assert b.singlepoint(xyz)["energies"][0] == e0
    return ref


@pytest.fixture(scope="module")
def xb():
    import xtb_b200
    if xtb_b200.device_count() < 1:
        pytest.fail("no CUDA device: the engine has no CPU fallback")
    return xtb_b200


@pytest.mark.parametrize("name", ["caffeine", "rivaroxaban", "feco5", "mindless02", "taxol"])
def test_decision_counts_match_oracle(xb, kat, name):
    s = kat[name]
    ref = _compare(xb, s["numbers"], s["xyz"], charge=s.get("charge", 0), uhf=s.get("uhf", 0))
    n = len(s["numbers"])
    assert ref["repulsion_pairs"] <= n * (n - 1) // 2 and ref["sdq_primitive_pairs"] > 0


def test_decision_counts_where_every_cutoff_bites(xb):
    """The 648-atom polypeptide is 100 bohr long: the 40 bohr (CN, repulsion, ATM, gradient integrals), 60 bohr (D4 gradient),
    72 bohr (D4 matrix), 10 bohr (bond orders) and sqrt(2000) bohr (integrals) cutoffs all remove pairs."""
    from xtb_b200.polypeptide import make_polypeptide
    Z, xyz, _ = make_polypeptide()
    ref = _compare(xb, Z, xyz)
    n = len(Z)
    npair = n * (n - 1) // 2
    assert ref["repulsion_pairs"] < npair and ref["d4_gradient_pairs"] < npair and ref["cn_d4_pairs"] < npair
    assert ref["repulsion_pairs"] < ref["d4_gradient_pairs"]
    assert ref["wbo_pairs"] < ref["repulsion_pairs"]
    assert ref["atm_triples"] < n * (n - 1) * (n - 2) // 6


def test_decision_counts_in_a_ragged_batch(xb, kat):
    from oracle.oracle import Oracle
    names = ["caffeine", "mindless01", "rivaroxaban"]
    b = xb.BatchCalculator([kat[k]["numbers"] for k in names], charges=[kat[k].get("charge", 0) for k in names],
                           uhf=[kat[k].get("uhf", 0) for k in names])
    b.set_decision_counting(True)
    b.singlepoint(np.concatenate([np.asarray(kat[k]["xyz"]).reshape(-1, 3) for k in names]))
    for i, k in enumerate(names):
        s = kat[k]
        o = Oracle(np.asarray(s["numbers"], dtype=np.int32), np.asarray(s["xyz"]).reshape(-1, 3), charge=s.get("charge", 0),
                   uhf=s.get("uhf", 0))
        assert o.singlepoint() == 0
        assert list(b.decision_counts(i).values()) == o.get("decisions").astype(np.int64).tolist()
