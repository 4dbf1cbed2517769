"""GPU parity of the point charge embedding (xtb_setExternalCharges / xtb_releaseExternalCharges /
xtb_getPCGradient; src/embedding.f90:225-280,329-365) through the C ABI: against the reference's own golden values
(test/api/c_api_example.c:298-419, tolerances as stated there) and against the CPU oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def xb():
    import xtb_b200
    if xtb_b200.device_count() < 1:
        pytest.fail("no CUDA device: the engine has no CPU fallback")
    return xtb_b200


def test_c_api_example_second_test(xb, kat):
    """testSecond() of the reference's C API example replayed through the same entry points."""
    s = kat["h2o2_pcem_api"]
    k = s["kat"]
    pcx = np.array(s["pc_xyz"]).reshape(-1, 3)
    calc = xb.Calculator(xb.Param.GFN2xTB, s["numbers"], s["xyz"])
    calc.set_verbosity(0)
    calc.set_external_charges(s["pc_numbers"], s["pc_charges"], pcx)
    res = calc.singlepoint()
    assert abs(res.get_energy() - k["energy"]) < k["tol_energy"]
    g, pg = res.get_gradient(), res.get_pc_gradient(len(pcx))
    for key, val in k.items():
        if key.startswith("g_"):
            i, c = map(int, key[2:].split("_"))
            assert abs(g[i, c] - val) < k["tol_grad"], key
        if key.startswith("pcg_"):
            i, c = map(int, key[4:].split("_"))
            assert abs(pg[i, c] - val) < k["tol_grad"], key
    # xtb_releaseExternalCharges: back to the gas-phase dimer, and no point-charge gradient in fresh results
    calc.release_external_charges()
    res2 = calc.singlepoint()
    from oracle.oracle import Oracle
    o = Oracle(s["numbers"], s["xyz"])
    assert o.singlepoint() == 0
    assert abs(res2.get_energy() - o.energies["e_total"]) < 1e-7
    with pytest.raises(xb.XTBException):
        res2.get_pc_gradient(len(pcx))


def test_batched_point_charges_match_oracle(xb, kat):
    """Ragged batch: caffeine without charges, the water dimer with six, rivaroxaban with two."""
    from oracle.oracle import Oracle
    rng = np.random.default_rng(2)
    mols = [kat["caffeine"], kat["h2o2_pcem_api"], kat["rivaroxaban"]]
    pcs = [dict(numbers=[], q=[], xyz=np.zeros((0, 3))),
           dict(numbers=mols[1]["pc_numbers"], q=mols[1]["pc_charges"], xyz=np.array(mols[1]["pc_xyz"]).reshape(-1, 3)),
           dict(numbers=[11, 17], q=[1.0, -1.0], xyz=np.array([[14.0, 0.5, -0.3], [-15.0, 1.0, 2.0]]))]
    b = xb.BatchCalculator([m["numbers"] for m in mols])
    b.set_external_charges([len(p["q"]) for p in pcs], np.concatenate([p["numbers"] for p in pcs]).astype(np.int32),
                           np.concatenate([p["q"] for p in pcs]), np.concatenate([p["xyz"] for p in pcs]))
    r = b.singlepoint(np.concatenate([m["xyz"] for m in mols]))
    assert r["all_converged"]
    pcg = b.array("pcgradient").reshape(-1, 3)
    off, poff = 0, 0
    for i, (m, p) in enumerate(zip(mols, pcs)):
        o = Oracle(m["numbers"], m["xyz"])
        o.set_external_charges(p["numbers"], p["q"], p["xyz"])
        assert o.singlepoint() == 0
        n = len(m["numbers"])
        assert abs(r["energies"][i] - o.energies["e_total"]) < 1e-7
        assert np.max(np.abs(r["gradients"][off:off + n] - o.get("gradient").reshape(-1, 3))) < 1e-6
        assert np.max(np.abs(r["charges"][off:off + n] - o.get("q"))) < 1e-6
        assert int(b.result(i)["niter"]) == o.info["niter"]
        if len(p["q"]):
            assert np.max(np.abs(pcg[poff:poff + len(p["q"])] - o.get("pcgradient").reshape(-1, 3))) < 1e-6
        off += n
        poff += len(p["q"])
    # release: the second molecule returns to its gas-phase energy
    b.set_external_charges(None)
    r2 = b.singlepoint(np.concatenate([m["xyz"] for m in mols]))
    o = Oracle(mols[1]["numbers"], mols[1]["xyz"])
    assert o.singlepoint() == 0
    assert abs(r2["energies"][1] - o.energies["e_total"]) < 1e-7
    assert abs(r2["energies"][1] - r["energies"][1]) > 1e-4
