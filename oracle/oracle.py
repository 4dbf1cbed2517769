"""ctypes wrapper of the CPU oracle (oracle/libgfn2_oracle.so).

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Never imported by xtb_b200/.
"""
import ctypes
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libgfn2_oracle.so")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(os.path.join(_HERE, "gfn2_oracle.cpp")):
        subprocess.check_call(["make", "-C", _HERE, "libgfn2_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        L.orc_create.restype = ctypes.c_void_p
        L.orc_create.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double, ctypes.c_int]
        L.orc_destroy.argtypes = [ctypes.c_void_p]
        L.orc_set_params.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.c_double, ctypes.c_int]
        L.orc_set_threads.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.orc_set_guess.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.orc_set_xyz.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.orc_set_pcem.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.orc_set_solvent.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_double, ctypes.c_double, ctypes.c_double]
        L.orc_set_restart.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        L.orc_singlepoint.argtypes = [ctypes.c_void_p]
        L.orc_len64.restype = ctypes.c_long
        L.orc_len64.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
        L.orc_get.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_long]
        L.orc_geti.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_long]
        L.orc_cn_erf.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.orc_eeq.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        _LIB = L
    return _LIB


INFO = ["niter", "converged", "nao", "nshell", "ndp", "nqp", "nel", "nopen", "ihomo", "ndim", "nbr"]
ENERGIES = ["e_total", "e_elec", "e_rep", "e_es", "e_aes", "e_axc", "e_disp", "hl_gap", "gnorm", "e_atm", "eel", "ed"]


class Oracle:
    """One GFN2-xTB calculation on the CPU oracle. Units: bohr, Hartree."""

    def __init__(self, numbers, positions, charge=0.0, uhf=0, acc=1.0, etemp=300.0, maxiter=250, guess="eeq",
                 threads=None):
        self.numbers = np.ascontiguousarray(numbers, dtype=np.int32)
        self.positions = np.ascontiguousarray(positions, dtype=np.float64).reshape(-1, 3)
        self.nat = len(self.numbers)
        L = lib()
        self.h = L.orc_create(self.nat, self.numbers.ctypes.data, self.positions.ctypes.data, float(charge), int(uhf))
        L.orc_set_params(self.h, float(acc), float(etemp), int(maxiter))
        L.orc_set_guess(self.h, {"eeq": 0, "sad": 1, "restart": 2}[guess])
        if threads is not None:
            L.orc_set_threads(self.h, int(threads))
        self.status = None

    def __del__(self):
        try:
            lib().orc_destroy(self.h)
        except Exception:
            pass

    def set_positions(self, positions):
        self.positions = np.ascontiguousarray(positions, dtype=np.float64).reshape(-1, 3)
        lib().orc_set_xyz(self.h, self.positions.ctypes.data)

    def set_external_charges(self, numbers, charges, positions):
        """xtb_setExternalCharges: element numbers (-> hardness), charges, positions [npc][3] in bohr; empty = release."""
        numbers = np.ascontiguousarray(numbers, dtype=np.int32)
        charges = np.ascontiguousarray(charges, dtype=np.float64)
        positions = np.ascontiguousarray(positions, dtype=np.float64).reshape(-1, 3)
        lib().orc_set_pcem(self.h, len(numbers), numbers.ctypes.data, charges.ctypes.data, positions.ctypes.data)

    def set_solvent(self, solvent, model="gbsa", kernel=None, state=1, temperature=298.15, ion_strength=0.0, ion_rad=0.0):
        """xtb_setSolvent: model "gbsa" (Still kernel) or "alpb" (P16 kernel) as the C API chooses them
        (src/api/calculator.f90:411-419); kernel "still"/"p16" overrides (the reference's unit tests combine
        alpb=True with the Still kernel); solvent None releases the model."""
        if solvent is None:
            lib().orc_set_solvent(self.h, None, 0, 0, 1, 298.15, 0.0, 0.0)
            return
        alpb = 1 if model == "alpb" else 0
        if kernel is None:
            kernel = "still" if model == "gbsa" else "p16"
        lib().orc_set_solvent(self.h, solvent.encode(), alpb, {"still": 1, "p16": 2}[kernel], int(state), float(temperature),
                              float(ion_strength), float(ion_rad))

    SOLVATION = ["g_solv", "g_born", "g_sasa", "g_hb", "g_shift", "g_scc", "g_born_grad", "adet"]

    def set_restart(self, qsh, dipm, qp):
        qsh = np.ascontiguousarray(qsh, dtype=np.float64)
        dipm = np.ascontiguousarray(dipm, dtype=np.float64)
        qp = np.ascontiguousarray(qp, dtype=np.float64)
        lib().orc_set_restart(self.h, qsh.ctypes.data, len(qsh), dipm.ctypes.data, qp.ctypes.data)
        lib().orc_set_guess(self.h, 2)

    def singlepoint(self):
        self.status = lib().orc_singlepoint(self.h)
        return self.status

    def get(self, name):
        L = lib()
        n = L.orc_len64(self.h, name.encode())
        if n < 0:
            raise KeyError(name)
        if name in ("mdlst", "mqlst", "info"):
            buf = np.zeros(n, dtype=np.int32)
            L.orc_geti(self.h, name.encode(), buf.ctypes.data, n)
        else:
            buf = np.zeros(n, dtype=np.float64)
            L.orc_get(self.h, name.encode(), buf.ctypes.data, n)
        return buf

    @property
    def info(self):
        return dict(zip(INFO, self.get("info").tolist()))

    @property
    def energies(self):
        return dict(zip(ENERGIES, self.get("energies").tolist()))

    def cn_erf(self):
        buf = np.zeros(self.nat)
        lib().orc_cn_erf(self.h, buf.ctypes.data)
        return buf

    def eeq(self):
        buf = np.zeros(self.nat)
        rc = lib().orc_eeq(self.h, buf.ctypes.data)
        if rc:
            raise RuntimeError("EEQ solve failed")
        return buf
