"""Host-side set-up of the implicit solvation model (xtb_b200/csrc/solvation.h through xtbb200_solventParameters; no GPU):
solvent name normalisation, parameter-set selection by kernel, derived constants (src/solv/model.f90:185-246,498-528,
src/solv/gbsa.f90:254-260, src/solv/state.f90:55-85) against the numbers of the reference's parameter files
(include/param_gbsa_h2o.fh:67-75 gfn2_h2o, include/param_alpb_water.fh:67-75 gfn2_alpb_water)."""
import ctypes
import math

import numpy as np
import pytest

AATOAU = 1.0 / 0.52917726
KCALTOAU = 1.0 / 627.50947428
KB = 3.166808578545117e-06


@pytest.fixture(scope="module")
def lib():
    import xtb_b200
    L = xtb_b200.load_library()
    L.xtbb200_solventParameters.restype = ctypes.c_int
    L.xtbb200_solventParameters.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                            ctypes.POINTER(ctypes.c_double)]
    L.xtbb200_lastError.restype = ctypes.c_char_p
    return L


def _par(lib, name, alpb, kernel, state=1, temp=298.15):
    out = np.zeros(8)
    rc = lib.xtbb200_solventParameters(name.encode(), alpb, kernel, state, temp, out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
    return rc, out


def test_water_sets(lib):
    # gfn2_h2o (GBSA, Still kernel): epsv 80.2, c1 1.55243817, rprobe 0.97531379, gshift 1.16556316, soset 0.13032636
    rc, p = _par(lib, "h2o", 0, 1)
    assert rc == 0
    assert p[0] == 80.2 and p[2] == 0.0 and abs(p[1] - (1.0 / 80.2 - 1.0)) < 1e-15
    assert abs(p[4] - 1.55243817) < 1e-12 and abs(p[5] - 0.13032636 * 0.1 * AATOAU) < 1e-12
    assert abs(p[6] - 0.97531379 * AATOAU) < 1e-12 and abs(p[3] - 1.16556316 * KCALTOAU) < 1e-14 and p[7] == 1.0
    # gfn2_alpb_water (P16 kernel) with ALPB: alpha * beta = 0.571412 / eps enters keps
    rc, q = _par(lib, "Water", 1, 2)
    assert rc == 0
    ab = 0.571412 / 80.2
    assert abs(q[2] - ab) < 1e-16 and abs(q[1] - (1.0 / 80.2 - 1.0) / (1.0 + ab)) < 1e-15
    assert abs(q[4] - 1.47438678) < 1e-12 and abs(q[3] - 0.67818695 * KCALTOAU) < 1e-14
    assert abs(q[6] - 1.13409020 * AATOAU) < 1e-12 and q[5] == 0.0
    # the parameter set follows the kernel, not the alpb flag (the reference's unit tests use alpb + Still)
    rc, r = _par(lib, "water", 1, 1)
    assert rc == 0 and abs(r[4] - 1.55243817) < 1e-12 and abs(r[2] - ab) < 1e-16


def test_name_normalisation_and_missing_sets(lib):
    for a, b in [("dcm", "ch2cl2"), ("dichloromethane", "CH2Cl2"), ("chloroform", "chcl3"), ("DMSO", "dimethylsulfoxide"),
                 ("thf", "tetrahydrofuran"), ("n-hexane", "hexane"), ("ether", "diethylether"), ("furan", "furane")]:
        for kernel in (1, 2):
            ra, pa = _par(lib, a, 0, kernel)
            rb, pb = _par(lib, b, 0, kernel)
            assert ra == rb and (ra != 0 or np.array_equal(pa, pb)), (a, b, kernel)
    assert _par(lib, "aniline", 1, 2)[0] == 0          # ALPB-only solvent
    rc, _ = _par(lib, "aniline", 0, 1)                  # ... has no GBSA set
    assert rc == -1 and b"solvent: 'aniline' is not parametrized" in lib.xtbb200_lastError()
    assert _par(lib, "unobtainium", 1, 2)[0] == -1
    assert _par(lib, "water", 0, 3)[0] == -1            # unknown kernel


def test_reference_state_shift(lib):
    # getStateShift, src/solv/state.f90:55-85; gfn2_alpb_water: smass 18.0, rhos 1.0
    base = _par(lib, "water", 1, 2, state=1)[1][3]
    t = 310.0
    dens = 1.0 / 18.0 * 8.92388e-2
    ref = t * KB * (math.log(24.79 * t / 298.15) + math.log(dens / 1.0e-3 * 1.0 / 18.0))
    bar = t * KB * math.log(24.79 * t / 298.15)
    assert abs(_par(lib, "water", 1, 2, state=2, temp=t)[1][3] - base - ref) < 1e-15
    assert abs(_par(lib, "water", 1, 2, state=3, temp=t)[1][3] - base - bar) < 1e-15
