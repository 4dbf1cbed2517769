"""Round-2 regression tests of the CUDA engine (through the C ABI): restart paths of open-shell systems, the results-info
extension, solver mode 5, failed-wavefunction handling."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def xb():
    import xtb_b200
    if xtb_b200.device_count() < 1:
        pytest.fail("no CUDA device: the engine has no CPU fallback")
    return xtb_b200


@pytest.mark.parametrize("name", ["mindless02", "fe_cnx6"])
def test_open_shell_restart_keeps_eigenvector_guess(xb, kat, name):
    """The open-shell bond-order path (alpha / beta P S, src/scf_module.F90:869-879) must not clobber the eigenvectors that
    warm-start the next restarted call: a second and third single point with restart=True on the same batch reproduce the
    cold result (Zm was overwritten by P_beta S)."""
    s = kat[name]
    assert s["uhf"] != 0
    kw = dict(charges=[s["charge"]], uhf=[s["uhf"]])
    cold = xb.BatchCalculator([s["numbers"]], **kw)
    r0 = cold.singlepoint(s["xyz"])
    assert r0["all_converged"]
    wbo0 = cold.array("wbo").copy()
    b = xb.BatchCalculator([s["numbers"]], **kw)
    b.singlepoint(s["xyz"])
    for _ in range(2):
        r = b.singlepoint(s["xyz"], restart=True)
        assert r["all_converged"]
        assert np.all(np.isfinite(r["energies"])) and np.all(np.isfinite(r["gradients"]))
        assert abs(r["energies"][0] - r0["energies"][0]) < 1e-7
        assert np.max(np.abs(r["gradients"] - r0["gradients"])) < 1e-6
        # bond orders are quadratic in P: two SCC solutions converged to 1e-6 e differ by ~1e-5 there
        assert np.max(np.abs(b.array("wbo") - wbo0)) < 5e-5
        assert int(r["niter"][0]) <= int(r0["niter"][0])
    # displaced geometry, restarted, against a cold calculation of the same geometry
    x1 = s["xyz"] + 0.02 * np.random.default_rng(1).standard_normal(s["xyz"].shape)
    r1 = b.singlepoint(x1, restart=True)
    c1 = cold.singlepoint(x1)
    assert r1["all_converged"] and c1["all_converged"]
    assert abs(r1["energies"][0] - c1["energies"][0]) < 2e-7
    assert np.max(np.abs(r1["gradients"] - c1["gradients"])) < 2e-6


def test_open_shell_md_short(xb, kat):
    """A few MD steps of a radical (every step is a restarted single point): finite, energy-conserving to the usual noise."""
    s = kat["mindless02"]
    b = xb.BatchCalculator([s["numbers"]], charges=[s["charge"]], uhf=[s["uhf"]])
    b.md_init(s["xyz"], tstep_fs=0.5, temperature=100.0, thermostat=False, shake=0, hmass=0, seed=3)
    r = b.md_run(6)
    assert r["steps"] == 6
    etot = r["epot"][:, 0] + r["ekin"][:, 0]
    assert np.all(np.isfinite(etot))
    # the first entry holds the start velocities (t = 0), the later ones the leap-frog half-step average: conservation is
    # judged from the second step on
    assert np.max(np.abs(etot[1:] - etot[1])) < 1e-4


def test_mixed_batch_open_and_closed_restart(xb, kat):
    """Closed-shell members of a batch that contains a radical take the alpha / beta bond-order path too."""
    a, c = kat["mindless02"], kat["mindless01"]
    b = xb.BatchCalculator([a["numbers"], c["numbers"]], charges=[a["charge"], c["charge"]], uhf=[a["uhf"], c["uhf"]])
    pos = np.concatenate([a["xyz"], c["xyz"]])
    r0 = b.singlepoint(pos)
    r1 = b.singlepoint(pos, restart=True)
    assert r0["all_converged"] and r1["all_converged"]
    assert np.max(np.abs(r0["energies"] - r1["energies"])) < 1e-7
    assert np.max(np.abs(r0["gradients"] - r1["gradients"])) < 1e-6


def test_results_info_and_parameter_file_warning(xb, kat):
    """xtbb200_getResultsInfo returns the engine's gap / iteration count; a parameter file name is answered with a warning
    (the file is not read), any other file name with an error (src/api/calculator.f90:309-322)."""
    s = kat["c3h4_api"]
    calc = xb.Calculator(xb.Param.GFN2xTB, s["numbers"], s["xyz"])
    calc.set_verbosity(0)
    res = calc.singlepoint()
    info = res.get_info()
    b = xb.BatchCalculator([s["numbers"]])
    b.singlepoint(s["xyz"])
    rb = b.result(0)
    assert abs(info["hl_gap"] - rb["hl_gap"]) < 1e-9
    assert info["niter"] == int(rb["niter"])
    L = xb.load_library()
    env = ctypes.c_void_p(L.xtb_newEnvironment())
    nat = ctypes.c_int(len(s["numbers"]))
    num = np.ascontiguousarray(s["numbers"], dtype=np.int32)
    xyz = np.ascontiguousarray(s["xyz"], dtype=np.float64)
    mol = ctypes.c_void_p(L.xtb_newMolecule(env, ctypes.byref(nat), num.ctypes.data_as(ctypes.POINTER(ctypes.c_int)),
                                            xyz.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), None, None, None, None))
    c = ctypes.c_void_p(L.xtb_newCalculator())
    L.xtb_loadGFN2xTB(env, mol, c, b"param_gfn2-xtb.txt")
    assert L.xtb_checkEnvironment(env) == 0           # a warning, not an error
    L.xtb_loadGFN2xTB(env, mol, c, b"my_params.txt")
    assert L.xtb_checkEnvironment(env) == 1
    buf = ctypes.create_string_buffer(512)
    n = ctypes.c_int(512)
    L.xtb_getError(env, buf, ctypes.byref(n))
    assert b"not supported" in buf.value
    L.xtb_delCalculator(ctypes.byref(c)); L.xtb_delMolecule(ctypes.byref(mol)); L.xtb_delEnvironment(ctypes.byref(env))


def test_scc_iteration_table_at_verbosity_one(xb, kat, tmp_path):
    """xtb_singlepoint at verbosity >= 1 prints the reference's SCC iteration table (header src/scf_module.F90:612-616, rows
    src/scc_core.f90:621-622, format (i4,F15.7,E14.6,E11.3,f8.2,2x,f8.1,l3)): one row per iteration, the electronic energy
    of the last row is the converged one, dE and RMSdq of the last two rows are below the convergence thresholds."""
    from oracle.oracle import Oracle
    s = kat["caffeine"]
    calc = xb.Calculator(xb.Param.GFN2xTB, s["numbers"], s["xyz"])
    log = tmp_path / "xtb.out"
    calc.set_output(str(log))
    calc.set_verbosity(1)
    res = calc.singlepoint()
    calc.release_output()
    text = log.read_text().splitlines()
    head = [k for k, ln in enumerate(text) if ln.startswith(" iter      E             dE          RMSdq      gap      omega  full diag")]
    assert len(head) == 1
    rows = []
    for ln in text[head[0] + 1:]:
        if not ln.strip():
            break
        assert len(ln) == 4 + 15 + 14 + 11 + 8 + 2 + 8 + 3 and ln.endswith("  T")
        rows.append((int(ln[:4]), float(ln[4:19]), float(ln[19:33]), float(ln[33:44]), float(ln[44:52]), float(ln[54:62])))
    o = Oracle(s["numbers"], s["xyz"])
    assert o.singlepoint() == 0
    assert [r[0] for r in rows] == list(range(1, o.info["niter"] + 1))
    assert abs(rows[-1][1] - o.energies["eel"]) < 5e-7
    assert all(abs(r[2]) < 1e-6 and r[3] < 1e-4 for r in rows[-2:])          # two consecutive converged iterations
    assert abs(rows[-1][4] - res.get_info()["hl_gap"]) < 0.006
    assert rows[0][5] == 0.0 and all(r[5] >= 1.0 for r in rows[1:])          # Broyden weights: none in the first iteration
