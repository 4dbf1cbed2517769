"""CPU-only checks of the drop-in boundary: the shared library loads, exports every symbol declared in
include/*.h, and the host-side (no compute) part of the xtb.h contract behaves like the reference
(src/api/*.f90): handle ownership, validation, error log."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import xtb_b200
    if not os.path.exists(xtb_b200.library_path()):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "xtb_b200", "csrc")])
    return xtb_b200.load_library()


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    names = set(re.findall(r"\b(xtb_[A-Za-z0-9]+|xtbb200_[A-Za-z0-9]+)\s*(?:XTB_API_SUFFIX__VERSION_\w+\s*)?\(", src))
    return {n for n in names if n not in ("xtb_delete",)}


def test_every_declared_symbol_is_exported(lib):
    import xtb_b200
    out = subprocess.check_output(["nm", "-D", "--defined-only", xtb_b200.library_path()]).decode()
    exported = {ln.split()[-1] for ln in out.splitlines() if ln.strip()}
    declared = _declared("xtb.h") | _declared("xtb_b200.h")
    assert len(declared) > 50
    missing = sorted(declared - exported)
    assert not missing, missing
    for n in declared:
        getattr(lib, n)


def test_api_version_and_handles(lib):
    assert lib.xtb_getAPIVersion() == 20000
    vp = ctypes.c_void_p
    env = vp(lib.xtb_newEnvironment())
    assert env.value and lib.xtb_checkEnvironment(env) == 0
    res = vp(lib.xtb_newResults())
    e = ctypes.c_double()
    lib.xtb_getEnergy(env, res, ctypes.byref(e))       # error, not a crash (c_api_example.c:77-81)
    assert lib.xtb_checkEnvironment(env) == 1
    buf = ctypes.create_string_buffer(512)
    n = ctypes.c_int(512)
    lib.xtb_getError(env, buf, ctypes.byref(n))
    assert b"not" in buf.value.lower() and lib.xtb_checkEnvironment(env) == 0
    lib.xtb_delResults(ctypes.byref(res))
    assert res.value is None
    lib.xtb_delResults(ctypes.byref(res))              # tolerates NULL
    lib.xtb_delEnvironment(ctypes.byref(env))
    assert env.value is None


def test_molecule_validation_is_host_side():
    import xtb_b200
    with pytest.raises(xtb_b200.XTBException):
        xtb_b200.Molecule([1, 1], np.zeros((2, 3)))
    with pytest.raises(xtb_b200.XTBException):
        xtb_b200.Molecule([119], np.zeros((1, 3)))
    m = xtb_b200.Molecule([8, 1, 1], [[0, 0, 0], [0, 0, 1.8], [1.7, 0, -0.5]], charge=0.0, uhf=0)
    m.update([[0, 0, 0], [0, 0, 1.9], [1.7, 0, -0.5]])
    with pytest.raises(ValueError):
        m.update(np.zeros((2, 3)))


def test_engine_fails_loudly_without_gpu():
    """No CPU fallback: without a device, loading a calculator raises through the environment."""
    import xtb_b200
    if xtb_b200.device_count() > 0:
        pytest.skip("a GPU is visible here")
    with pytest.raises(xtb_b200.XTBException):
        xtb_b200.Calculator(xtb_b200.Param.GFN2xTB, [1, 1], [[0, 0, 0], [0, 0, 1.4]])
    with pytest.raises(xtb_b200.XTBException):
        xtb_b200.BatchCalculator([[1, 1]])


def test_product_does_not_import_oracle():
    for dp, _, fs in os.walk(os.path.join(ROOT, "xtb_b200")):
        for f in fs:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".h")) and "tables" not in dp:
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.lower() or f == "interface.py" and "oracle" not in txt, (dp, f)


def test_external_charges_error_convention(lib):
    """src/api/calculator.f90:470-540: errors go to the environment log, nothing is returned; releasing on a calculator
    without charges is not an error (545-580)."""
    vp = ctypes.c_void_p
    env, calc = vp(lib.xtb_newEnvironment()), vp(lib.xtb_newCalculator())
    n = ctypes.c_int(1)
    nu, q, x = np.array([8], dtype=np.int32), np.array([0.5]), np.zeros(3)
    ip, dp = ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_double)
    lib.xtb_setExternalCharges(env, calc, ctypes.byref(n), nu.ctypes.data_as(ip), q.ctypes.data_as(dp), x.ctypes.data_as(dp))
    assert lib.xtb_checkEnvironment(env) == 1           # no parametrisation loaded
    buf, ln = ctypes.create_string_buffer(512), ctypes.c_int(512)
    lib.xtb_getError(env, buf, ctypes.byref(ln))
    assert b"No calculator loaded" in buf.value
    n0 = ctypes.c_int(0)
    lib.xtb_setExternalCharges(env, calc, ctypes.byref(n0), nu.ctypes.data_as(ip), q.ctypes.data_as(dp), x.ctypes.data_as(dp))
    assert lib.xtb_checkEnvironment(env) == 1
    lib.xtb_getError(env, buf, ctypes.byref(ln))
    lib.xtb_releaseExternalCharges(env, calc)
    assert lib.xtb_checkEnvironment(env) == 0
    res = vp(lib.xtb_newResults())
    g = np.zeros(3)
    lib.xtb_getPCGradient(env, res, g.ctypes.data_as(dp))
    assert lib.xtb_checkEnvironment(env) == 1
    lib.xtb_getError(env, buf, ctypes.byref(ln))
    lib.xtb_delResults(ctypes.byref(res))
    lib.xtb_delCalculator(ctypes.byref(calc))
    lib.xtb_delEnvironment(ctypes.byref(env))


def test_solvent_calls_without_calculator(lib):
    """src/api/calculator.f90:341-476: xtb_setSolvent on an unloaded calculator is an environment error; xtb_releaseSolvent
    on a calculator without a solvation model is not."""
    vp = ctypes.c_void_p
    lib.xtb_setSolvent.argtypes = [vp, vp, ctypes.c_char_p, vp, vp, vp, ctypes.c_char_p]
    env, calc = vp(lib.xtb_newEnvironment()), vp(lib.xtb_newCalculator())
    lib.xtb_setSolvent(env, calc, b"h2o", None, None, None, None)
    assert lib.xtb_checkEnvironment(env) == 1
    buf, ln = ctypes.create_string_buffer(512), ctypes.c_int(512)
    lib.xtb_getError(env, buf, ctypes.byref(ln))
    assert b"No calculator loaded to add solvation model" in buf.value
    lib.xtb_releaseSolvent(env, calc)
    assert lib.xtb_checkEnvironment(env) == 0
    lib.xtb_delCalculator(ctypes.byref(calc))
    lib.xtb_delEnvironment(ctypes.byref(env))
