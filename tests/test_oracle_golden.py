"""Pin the CPU oracle against every known-answer value the reference's own tests hold for the
GFN2-xTB single point + gradient path."""
import numpy as np
import pytest

from oracle.oracle import Oracle


def _run(sys_, **kw):
    o = Oracle(sys_["numbers"], sys_["xyz"], charge=sys_["charge"], uhf=sys_["uhf"],
               guess=sys_.get("guess", "eeq"), maxiter=sys_.get("maxiter", 250), **kw)
    assert o.singlepoint() == 0
    return o


def test_h2o_scc(kat):
    # test/unit/test_gfn2.f90:48-170
    s = kat["h2o_scc"]
    k = s["kat"]
    o = _run(s)
    info, e = o.info, o.energies
    assert (info["nshell"], info["nao"], info["ihomo"]) == (k["nshell"], k["nao"], k["ihomo"])
    thr, thr2 = 1e-7, 1e-5
    assert abs(e["e_total"] - k["energy"]) < thr
    assert abs(e["gnorm"] - k["gnorm"]) < thr
    assert abs(e["hl_gap"] - k["hl_gap"]) < 1e-4
    assert abs(e["e_elec"] - k["e_elec"]) < thr
    assert abs(e["e_es"] - k["e_es"]) < thr * 10
    assert abs(e["e_aes"] - k["e_aes"]) < thr * 10
    assert abs(e["e_axc"] - k["e_axc"]) < thr
    assert abs(e["e_disp"] - k["e_disp"]) < thr
    assert abs(e["e_rep"] - k["e_rep"]) < thr
    q, qsh = o.get("q"), o.get("qsh")
    dipm, qp = o.get("dipm").reshape(-1, 3), o.get("qp").reshape(-1, 6)
    assert abs(q[1] - k["q1"]) < thr2 and abs(q[2] - q[1]) < thr2
    assert abs(qsh[0] - k["qsh0"]) < thr2
    assert abs(dipm[1, 1]) < thr2 and abs(dipm[2, 0] - k["dipm_2_0"]) < thr2
    assert abs(qp[0, 5] - k["qp_0_5"]) < thr2 and abs(qp[2, 0] - k["qp_2_0"]) < thr2
    assert abs(o.get("emo")[info["ihomo"] - 1] - k["emo_homo"]) < thr2
    assert abs(o.get("focc")[info["ihomo"] - 1] - 2.0) < thr2


def test_c3h4_api(kat):
    # test/unit/test_gfn2.f90:172-235 and test/api/c_api_example.c:87-112
    s = kat["c3h4_api"]
    k = s["kat"]
    o = _run(s)
    e = o.energies
    g = o.get("gradient").reshape(-1, 3)
    assert abs(e["e_total"] - k["energy"]) < 1e-9
    assert abs(e["hl_gap"] - k["hl_gap"]) < 1e-9
    assert abs(e["gnorm"] - k["gnorm"]) < 1e-9
    assert abs(g[2, 1] - k["g_2_1"]) < 1e-9
    assert abs(g[0, 2] - k["g_0_2"]) < 1e-9
    assert abs(g[3, 0] - k["g_3_0"]) < 1e-9
    assert abs(g[6, 2] - k["g_6_2"]) < 1e-9
    assert abs(o.get("q")[5] - k["q5"]) < 1e-8
    assert abs(o.get("dipole")[2] - k["dipole2"]) < 1e-6
    assert abs(o.get("wbo")[9] - k["wbo9"]) < 1e-8


@pytest.mark.parametrize("name", ["mindless%02d" % i for i in range(1, 11)] + ["feco5", "co_cnx6", "fe_cnx6"])
def test_molstock_kat(kat, name):
    # test/unit/test_gfn2.f90:667-742 (mindless), 831-901 (feco5 gas phase), 987-1091 (wbo, open shell)
    s = kat[name]
    k = s["kat"]
    o = _run(s)
    e = o.energies
    assert abs(e["e_total"] - k["energy"]) < k["tol"]
    assert abs(e["gnorm"] - k["gnorm"]) < k["tol"]
    assert abs(e["hl_gap"] - k["hl_gap"]) < k["tol"]
    if "wbo" in k:
        n = len(s["numbers"])
        w = o.get("wbo").reshape(n, n)
        for key, val in k["wbo"].items():
            i, j = map(int, key.split(","))
            assert abs(w[i, j] - val) < k["tol"]


def test_gradient_matches_finite_difference(kat):
    # the reference's component tests check analytic vs numerical gradients
    s = kat["c3h4_api"]
    o = Oracle(s["numbers"], s["xyz"], acc=0.01)
    assert o.singlepoint() == 0
    g = o.get("gradient").reshape(-1, 3)
    h = 1e-4
    for (i, k) in [(0, 2), (3, 0), (6, 2), (4, 1)]:
        ep = []
        for sgn in (+1, -1):
            x = s["xyz"].copy()
            x[i, k] += sgn * h
            o2 = Oracle(s["numbers"], x, acc=0.01)
            assert o2.singlepoint() == 0
            ep.append(o2.energies["e_total"])
        fd = (ep[0] - ep[1]) / (2 * h)
        assert abs(fd - g[i, k]) < 2e-7


def test_restart_converges_immediately(kat):
    s = kat["caffeine"]
    o = _run(s)
    n1 = o.info["niter"]
    e1 = o.energies["e_total"]
    o.set_restart(o.get("qsh"), o.get("dipm"), o.get("qp"))
    assert o.singlepoint() == 0
    assert o.info["niter"] <= 3 < n1
    assert abs(o.energies["e_total"] - e1) < 1e-7


def test_external_point_charges_api_kat(kat):
    # test/api/c_api_example.c:298-419: water dimer embedded in the point charges of a second dimer;
    # energy 1e-9, gradient and point-charge gradient components 1e-8 (the reference's own tolerances)
    s = kat["h2o2_pcem_api"]
    k = s["kat"]
    o = Oracle(s["numbers"], s["xyz"])
    o.set_external_charges(s["pc_numbers"], s["pc_charges"], np.array(s["pc_xyz"]).reshape(-1, 3))
    assert o.singlepoint() == 0
    assert abs(o.energies["e_total"] - k["energy"]) < k["tol_energy"]
    g = o.get("gradient").reshape(-1, 3)
    pg = o.get("pcgradient").reshape(-1, 3)
    for key, val in k.items():
        if key.startswith("g_"):
            i, c = map(int, key[2:].split("_"))
            assert abs(g[i, c] - val) < k["tol_grad"], key
        if key.startswith("pcg_"):
            i, c = map(int, key[4:].split("_"))
            assert abs(pg[i, c] - val) < k["tol_grad"], key
    # releasing the charges gives back the gas-phase dimer
    o.set_external_charges([], [], np.zeros((0, 3)))
    assert o.singlepoint() == 0
    o2 = Oracle(s["numbers"], s["xyz"])
    assert o2.singlepoint() == 0
    assert abs(o.energies["e_total"] - o2.energies["e_total"]) < 1e-12


def test_external_point_charges_gradient_is_consistent(kat):
    # finite differences of the embedded energy with respect to an atom and to a point charge position
    s = kat["h2o2_pcem_api"]
    pcx = np.array(s["pc_xyz"]).reshape(-1, 3)

    def run(xyz, pc):
        o = Oracle(s["numbers"], xyz, acc=0.01)
        o.set_external_charges(s["pc_numbers"], s["pc_charges"], pc)
        assert o.singlepoint() == 0
        return o

    o = run(s["xyz"], pcx)
    g = o.get("gradient").reshape(-1, 3)
    pg = o.get("pcgradient").reshape(-1, 3)
    h = 1e-4
    for which, i, c in [("atom", 1, 1), ("pc", 3, 0)]:
        e = []
        for sgn in (+1, -1):
            x, p = s["xyz"].copy(), pcx.copy()
            (x if which == "atom" else p)[i, c] += sgn * h
            e.append(run(x, p).energies["e_total"])
        fd = (e[0] - e[1]) / (2 * h)
        assert abs(fd - (g if which == "atom" else pg)[i, c]) < 2e-7


# ---- implicit solvation (GBSA / ALPB), SURVEY 8f row f2 ----

AATOAU = 1.0 / 0.52917726


def test_c3h4_api_gbsa_alpb(kat):
    # test/api/c_api_example.c:114-163: gas phase, GBSA(h2o), ALPB(h2o), each restarted from the previous results
    s = kat["c3h4_api"]
    ks = s["kat_solv"]
    o = _run(s)
    for model in ("gbsa", "alpb"):
        k = ks[model]
        o.set_restart(o.get("qsh"), o.get("dipm"), o.get("qp"))
        o.set_solvent(k["solvent"], model)
        assert o.singlepoint() == 0
        assert abs(o.energies["e_total"] - k["energy"]) < ks["tol_energy"]
        assert abs(o.get("q")[5] - k["q5"]) < ks["tol_q"]
        assert abs(o.get("dipole")[2] - k["dipole2"]) < ks["tol_dipole"]
        assert abs(o.get("wbo")[9] - k["wbo9"]) < ks["tol_wbo"]
    # releasing the solvent gives the gas phase again
    o.set_solvent(None)
    o.set_restart(o.get("qsh"), o.get("dipm"), o.get("qp"))
    assert o.singlepoint() == 0
    assert abs(o.energies["e_total"] - s["kat"]["energy"]) < 1e-8


@pytest.mark.parametrize("name", ["gbsa_h2o_api", "gbsa_salt_api"])
def test_gbsa_energy_and_gradient(kat, name):
    # test/unit/test_gfn2.f90:237-311 (GBSA water) and 313-385 (GBSA CH2Cl2 with salt screening)
    s = kat[name]
    k = s["kat"]
    o = Oracle(s["numbers"], s["xyz"], maxiter=s["maxiter"])
    o.set_solvent(s["solvent"], s["model"], s["kernel"], ion_strength=s.get("ion_strength", 0.0),
                  ion_rad=s.get("ion_rad_angstrom", 0.0) * AATOAU)
    assert o.singlepoint() == 0
    e = o.energies
    g = o.get("gradient").reshape(-1, 3)
    assert abs(e["e_total"] - k["energy"]) < k["tol"]
    assert abs(e["gnorm"] - k["gnorm"]) < k["tol"]
    assert abs(e["hl_gap"] - k["hl_gap"]) < k["tol_gap"]
    for key, val in k.items():
        if key.startswith("g_"):
            _, i, c = key.split("_")
            assert abs(g[int(i), int(c)] - val) < k["tol"]
    if name == "gbsa_h2o_api":
        assert abs(g[10, 2] + g[9, 2]) < k["tol"]   # gradient(3,11) = -gradient(3,10)


@pytest.mark.parametrize("i", range(1, 11))
def test_mindless_solvation(kat, i):
    # test/unit/test_gfn2.f90:745-828: ten molecules, ten solvents, GBSA / ALPB alternating, Still kernel
    s = kat["mindless%02d" % i]
    k = s["kat_solv"]
    o = Oracle(s["numbers"], s["xyz"], charge=s["charge"], uhf=s["uhf"])
    o.set_solvent(k["solvent"], k["model"], k["kernel"])
    assert o.singlepoint() == 0
    e = o.energies
    assert abs(e["e_total"] - k["energy"]) < k["tol"]
    assert abs(e["gnorm"] - k["gnorm"]) < k["tol"]
    assert abs(e["hl_gap"] - k["hl_gap"]) < k["tol"]


@pytest.mark.parametrize("model", ["gbsa", "alpb"])
def test_feco5_solvation(kat, model):
    # test/unit/test_gfn2.f90:862-900
    s = kat["feco5"]
    k = s["kat_solv"]
    o = Oracle(s["numbers"], s["xyz"])
    o.set_solvent(k["solvent"], model, k["kernel"])
    assert o.singlepoint() == 0
    e = o.energies
    assert abs(e["e_total"] - k[model]["energy"]) < k["tol"]
    assert abs(e["gnorm"] - k[model]["gnorm"]) < k["tol"]
    assert abs(e["hl_gap"] - k[model]["hl_gap"]) < k["tol"]


def test_alpb_p16_gradient_is_consistent(kat):
    # ALPB with the P16 kernel (what xtb_setSolvent(..., "alpb") selects): analytic gradient against central differences
    s = kat["gbsa_h2o_api"]
    x0 = np.array(s["xyz"]).reshape(-1, 3)

    def run(xyz):
        o = Oracle(s["numbers"], xyz, acc=0.01)
        o.set_solvent("water", "alpb")
        assert o.singlepoint() == 0
        return o

    g = run(x0).get("gradient").reshape(-1, 3)
    h = 1e-4
    for i, c in [(0, 0), (4, 1), (9, 2)]:
        e = []
        for sgn in (+1, -1):
            x = x0.copy()
            x[i, c] += sgn * h
            e.append(run(x).energies["e_total"])
        assert abs((e[0] - e[1]) / (2 * h) - g[i, c]) < 3e-7
