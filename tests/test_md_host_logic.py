"""Host logic of the MD driver without a GPU: the SHAKE constraint selection and the masses the engine sets up
(xtb_b200/csrc/md_host.h through xtbb200_shakeList) against the CPU restatement of init_shake (oracle/md_oracle.py;
src/shake_module.f90:50-148, src/dynamic.f90:237-252) on bond orders from the oracle."""
import ctypes

import numpy as np
import pytest

from oracle.md_oracle import AMUTOAU, MASS, init_shake
from oracle.oracle import Oracle


@pytest.fixture(scope="module")
def lib():
    import xtb_b200
    L = xtb_b200.load_library()
    ip, dp = ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_double)
    L.xtbb200_shakeList.restype = ctypes.c_longlong
    L.xtbb200_shakeList.argtypes = [ctypes.c_int, ip, dp, dp, ctypes.c_int, ctypes.c_int, ip, ctypes.c_longlong, dp]
    return L


def _shake_list(lib, numbers, xyz, wbo, mode, hmass=4):
    ip, dp = ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_double)
    n = len(numbers)
    z = np.ascontiguousarray(numbers, dtype=np.int32)
    x = np.ascontiguousarray(xyz, dtype=np.float64)
    w = np.ascontiguousarray(wbo, dtype=np.float64)
    pairs = np.zeros(2 * 4 * n, dtype=np.int32)
    mass = np.zeros(n)
    nc = lib.xtbb200_shakeList(n, z.ctypes.data_as(ip), x.ctypes.data_as(dp), w.ctypes.data_as(dp), mode, hmass,
                               pairs.ctypes.data_as(ip), 4 * n, mass.ctypes.data_as(dp))
    assert nc >= 0
    return [tuple(p) for p in pairs[:2 * nc].reshape(-1, 2).tolist()], mass


@pytest.mark.parametrize("name", ["caffeine", "rivaroxaban", "feco5", "mindless01"])
def test_constraint_lists_match_init_shake(lib, kat, name):
    s = kat[name]
    n = len(s["numbers"])
    o = Oracle(s["numbers"], s["xyz"], charge=s["charge"], uhf=s["uhf"])
    assert o.singlepoint() == 0
    wbo = o.get("wbo").reshape(n, n)
    for mode in (1, 2):
        ref, _ = init_shake(s["numbers"], s["xyz"], wbo, mode)
        got, mass = _shake_list(lib, s["numbers"], s["xyz"], wbo, mode)
        assert got == [tuple(p) for p in ref], (name, mode)
    if name == "feco5":      # metal-ligand bonds are left free, the five C-O bonds are constrained
        assert len(got) == 5 and all(0 not in p for p in got)
    ref_mass = np.array([MASS[z - 1] for z in s["numbers"]]) * AMUTOAU
    ref_mass[np.asarray(s["numbers"]) == 1] = 4.0 * AMUTOAU
    assert np.max(np.abs(mass - ref_mass) / ref_mass) < 1e-15
    _, mass0 = _shake_list(lib, s["numbers"], s["xyz"], wbo, 0, hmass=0)
    assert np.max(np.abs(mass0 - np.array([MASS[z - 1] for z in s["numbers"]]) * AMUTOAU)) < 1e-9


def test_shake_list_errors(lib, kat):
    s = kat["caffeine"]
    ip, dp = ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_double)
    z = np.ascontiguousarray(s["numbers"], dtype=np.int32)
    x = np.ascontiguousarray(s["xyz"], dtype=np.float64)
    assert lib.xtbb200_shakeList(len(z), z.ctypes.data_as(ip), x.ctypes.data_as(dp), None, 2, 4, None, 0, None) == -1   # no bond orders
    assert lib.xtbb200_shakeList(len(z), z.ctypes.data_as(ip), x.ctypes.data_as(dp), None, 3, 4, None, 0, None) == -1   # unknown mode
    assert lib.xtbb200_shakeList(len(z), z.ctypes.data_as(ip), x.ctypes.data_as(dp), None, 1, 4, None, 0, None) == 10   # X-H bonds
