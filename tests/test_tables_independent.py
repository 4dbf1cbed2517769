"""The compiled-in GFN2 element tables (xtb_b200/csrc/tables/gfn2_tables.h, generated from the reference's Fortran data
statements by tools/gen_tables.py) against an independent source: the reference's parameter FILE param_gfn2-xtb.txt, parsed
by tests/golden/make_table_golden.py into tests/golden/gfn2_param_file.json.  Oracle and engine include the same header, so
this is the check that is not common-mode: all 86 elements, every per-element parameter of the Hamiltonian."""
import json
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 5.0e-7          # the file holds six decimals


def _header_arrays():
    text = open(os.path.join(ROOT, "xtb_b200", "csrc", "tables", "gfn2_tables.h")).read()
    arrays = {}
    for m in re.finditer(r"static const (double|int) (\w+)((?:\[\d+\])+) = \{(.*?)\};", text, re.S):
        typ, name, dims, body = m.groups()
        shape = [int(x) for x in re.findall(r"\[(\d+)\]", dims)]
        vals = [float(x) for x in re.findall(r"[-+]?\d+\.?\d*(?:[eE][-+]?\d+)?", body)]
        arrays[name] = np.array(vals).reshape(shape)
    return arrays


def test_element_tables_match_the_parameter_file():
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "gfn2_param_file.json")))
    A = _header_arrays()
    assert len(ref["elements"]) == 86
    for zs, e in ref["elements"].items():
        i = int(zs) - 1
        nsh = len(e["ang"])
        assert int(A["gfn2_nshell"][i]) == nsh, zs
        assert A["gfn2_ang_shell"][i][:nsh].astype(int).tolist() == e["ang"], zs
        assert A["gfn2_pqn"][i][:nsh].astype(int).tolist() == e["pqn"], zs
        np.testing.assert_allclose(A["gfn2_self_energy"][i][:nsh], e["lev"], atol=TOL, rtol=0, err_msg=zs)
        np.testing.assert_allclose(A["gfn2_slater_exponent"][i][:nsh], e["exp"], atol=TOL, rtol=0, err_msg=zs)
        assert abs(A["gfn2_chemical_hardness"][i] - e["gam"][0]) < TOL, zs
        assert abs(A["gfn2_third_order_atom"][i] - 0.1 * e["gam3"][0]) < TOL, zs
        assert abs(A["gfn2_dip_kernel"][i] - 0.01 * e["dpol"][0]) < TOL, zs
        assert abs(A["gfn2_quad_kernel"][i] - 0.01 * e["qpol"][0]) < TOL, zs
        assert abs(A["gfn2_rep_alpha"][i] - e["repa"][0]) < TOL, zs
        assert abs(A["gfn2_rep_zeff"][i] - e["repb"][0]) < TOL, zs
        for l, key in enumerate(("kcns", "kcnp", "kcnd")):
            assert abs(A["gfn2_kcn_l"][i][l] - 0.1 * e.get(key, [0.0])[0]) < TOL, (zs, key)
        for l, key in enumerate(("polys", "polyp", "polyd")):
            assert abs(A["gfn2_shell_poly_l"][i][l] - e.get(key, [0.0])[0]) < TOL, (zs, key)
        for l, key in enumerate(("lpars", "lparp", "lpard")):
            assert abs(A["gfn2_shell_hardness_l"][i][l] - 0.1 * e.get(key, [0.0])[0]) < TOL, (zs, key)


def test_global_parameters_match_the_parameter_file():
    """Globals of $globpar against the values the engine and the oracle compile in."""
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "gfn2_param_file.json")))["globals"]
    assert (g["ks"], g["kp"], g["kd"], g["ksd"], g["kpd"], g["kdiff"]) == (1.85, 2.23, 2.23, 2.0, 2.0, 2.0)
    assert (g["a1"], g["a2"], g["s8"], g["s9"]) == (0.52, 5.0, 2.7, 5.0)
    assert (g["aesdmp3"], g["aesdmp5"], g["aesshift"], g["aesexp"], g["aesrmax"]) == (3.0, 4.0, 1.2, 4.0, 5.0)
    assert (g["gam3s"], g["gam3p"], g["gam3d1"], g["gam3d2"]) == (1.0, 0.5, 0.25, 0.25)
    # where the engine holds them: the Hamiltonian scaling constants of k_sdqh0
    src = open(os.path.join(ROOT, "xtb_b200", "csrc", "kernels_int.cuh")).read()
    assert "kshell[3] = {1.85, 2.23, 2.23}" in src
