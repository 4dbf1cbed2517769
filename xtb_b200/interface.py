"""ctypes binding of the C ABI (include/xtb.h, include/xtb_b200.h).

Class and method names follow the reference's Python binding (xtb-python: xtb.interface.Calculator /
Results / Param / XTBException, itself a thin wrapper of include/xtb.h) so that code written against
it reads the same here.  Units: Bohr, Hartree.
"""
import ctypes
import enum
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_c_double_p = ctypes.POINTER(ctypes.c_double)
_c_int_p = ctypes.POINTER(ctypes.c_int)


class XTBException(Exception):
    """Error raised from the environment log of the C API (xtb_checkEnvironment != 0)."""


def library_path():
    # XTBB200_LIB: diagnostic builds of the same engine (e.g. clock64-instrumented kernels); never a different backend
    return os.environ.get("XTBB200_LIB") or os.path.join(_HERE, "lib", "libxtb_b200.so")


def load_library():
    """Load the CUDA engine; fails loudly when it has not been built (no fallback path exists)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise XTBException("CUDA engine %s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)" % path)
    L = ctypes.CDLL(path)
    vp, cp, ci, cd = ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_double
    sig = {
        "xtb_getAPIVersion": (ci, []),
        "xtb_newEnvironment": (vp, []), "xtb_delEnvironment": (None, [ctypes.POINTER(vp)]),
        "xtb_checkEnvironment": (ci, [vp]), "xtb_showEnvironment": (None, [vp, cp]),
        "xtb_getError": (None, [vp, cp, _c_int_p]), "xtb_setOutput": (None, [vp, cp]),
        "xtb_releaseOutput": (None, [vp]), "xtb_setVerbosity": (None, [vp, ci]),
        "xtb_newMolecule": (vp, [vp, _c_int_p, _c_int_p, _c_double_p, _c_double_p, _c_int_p, _c_double_p, vp]),
        "xtb_delMolecule": (None, [ctypes.POINTER(vp)]), "xtb_updateMolecule": (None, [vp, vp, _c_double_p, _c_double_p]),
        "xtb_newCalculator": (vp, []), "xtb_delCalculator": (None, [ctypes.POINTER(vp)]),
        "xtb_loadGFN0xTB": (None, [vp, vp, vp, cp]), "xtb_loadGFN1xTB": (None, [vp, vp, vp, cp]),
        "xtb_loadGFN2xTB": (None, [vp, vp, vp, cp]), "xtb_loadGFNFF": (None, [vp, vp, vp, cp]),
        "xtb_setSolvent": (None, [vp, vp, cp, _c_int_p, _c_double_p, _c_int_p, cp]),
        "xtb_releaseSolvent": (None, [vp, vp]),
        "xtb_setExternalCharges": (None, [vp, vp, _c_int_p, _c_int_p, _c_double_p, _c_double_p]),
        "xtb_releaseExternalCharges": (None, [vp, vp]),
        "xtb_setAccuracy": (None, [vp, vp, cd]), "xtb_setMaxIter": (None, [vp, vp, ci]),
        "xtb_setElectronicTemp": (None, [vp, vp, cd]),
        "xtb_cpcmx_calc": (None, [vp, vp, vp, vp]), "xtb_singlepoint": (None, [vp, vp, vp, vp]),
        "xtb_hessian": (None, [vp, vp, vp, vp, _c_double_p, _c_int_p, _c_int_p, _c_double_p, _c_double_p]),
        "xtb_newResults": (vp, []), "xtb_delResults": (None, [ctypes.POINTER(vp)]), "xtb_copyResults": (vp, [vp]),
        "xtb_getEnergy": (None, [vp, vp, _c_double_p]), "xtb_getSolvationEnergy": (None, [vp, vp, _c_double_p]),
        "xtb_getGradient": (None, [vp, vp, _c_double_p]), "xtb_getPCGradient": (None, [vp, vp, _c_double_p]),
        "xtb_getVirial": (None, [vp, vp, _c_double_p]), "xtb_getDipole": (None, [vp, vp, _c_double_p]),
        "xtb_getCharges": (None, [vp, vp, _c_double_p]), "xtb_getBondOrders": (None, [vp, vp, _c_double_p]),
        "xtb_getNao": (None, [vp, vp, _c_int_p]), "xtb_getOrbitalEigenvalues": (None, [vp, vp, _c_double_p]),
        "xtb_getOrbitalOccupations": (None, [vp, vp, _c_double_p]),
        "xtb_getOrbitalCoefficients": (None, [vp, vp, _c_double_p]),
        # batched extension
        "xtbb200_lastError": (cp, []), "xtbb200_deviceCount": (ci, []),
        "xtbb200_newBatch": (vp, [ci, _c_int_p, _c_int_p, _c_double_p, _c_int_p, ci]),
        "xtbb200_delBatch": (None, [ctypes.POINTER(vp)]),
        "xtbb200_setOptions": (ci, [vp, cd, cd, ci, ci]),
        "xtbb200_setSolver": (ci, [vp, ci]),
        "xtbb200_setSolvent": (ci, [vp, cp, ci, ci, ci, cd, cd, cd]),
        "xtbb200_mdInit": (ci, [vp, _c_double_p, cd, cd, ci, ci, ci, _c_double_p, ctypes.c_ulonglong, ci]),
        "xtbb200_mdRun": (ci, [vp, ci, _c_double_p, _c_double_p, _c_double_p]),
        "xtbb200_mdConstraints": (ctypes.c_longlong, [vp]),
        "xtbb200_setExternalCharges": (ci, [vp, _c_int_p, _c_int_p, _c_double_p, _c_double_p]),
        "xtbb200_symmetricEigen": (ci, [ci, _c_double_p, _c_double_p, _c_double_p, ctypes.POINTER(ctypes.c_float), ci]),
        "xtbb200_setWavefunction": (ci, [vp, _c_double_p, _c_double_p, _c_double_p]),
        "xtbb200_shardUniqueId": (ci, [ctypes.c_char_p, ctypes.c_size_t]),
        "xtbb200_setShard": (ci, [vp, ci, ci, ctypes.c_char_p, ctypes.c_size_t]),
        "xtbb200_shardOwner": (ci, [ci, ctypes.c_longlong, ci]),
        "xtbb200_getCommTime": (ci, [vp, ctypes.POINTER(ctypes.c_float)]),
        "xtbb200_singlepoint": (ci, [vp, _c_double_p, ci, _c_double_p, _c_double_p, _c_double_p, _c_int_p, _c_int_p]),
        "xtbb200_singlepointDevice": (ci, [vp, vp, ci, vp]),
        "xtbb200_getNMol": (ci, [vp]), "xtbb200_getSizes": (ci, [vp, ci, _c_int_p, _c_int_p, _c_int_p]),
        "xtbb200_getElectrons": (ci, [vp, ci, _c_int_p, _c_int_p]),
        "xtbb200_getDeviceBytes": (ctypes.c_size_t, [vp]),
        "xtbb200_getResultsInfo": (ci, [vp, _c_double_p, _c_int_p]),
        "xtbb200_getResult": (ci, [vp, ci, _c_double_p]), "xtbb200_getCounts": (ci, [vp, ci, _c_int_p, _c_int_p, _c_int_p]),
        "xtbb200_setDecisionCounting": (ci, [vp, ci]),
        "xtbb200_getDecisionCounts": (ci, [vp, ci, ctypes.POINTER(ctypes.c_longlong)]),
        "xtbb200_getArrayLen": (ci, [vp, cp, ctypes.POINTER(ctypes.c_size_t)]),
        "xtbb200_getArray": (ci, [vp, cp, _c_double_p, ctypes.c_size_t]),
        "xtbb200_getOffsets": (ci, [vp, ci, ctypes.POINTER(ctypes.c_longlong)]),
        "xtbb200_getMask": (ci, [vp, ctypes.c_void_p, ctypes.c_size_t]),
        "xtbb200_getTimings": (ci, [vp, ctypes.POINTER(ctypes.c_float), _c_int_p]),
        "xtbb200_getCallTime": (ci, [vp, ctypes.POINTER(ctypes.c_float)]),
        "xtbb200_getFactorisationTime": (ci, [vp, ctypes.POINTER(ctypes.c_float)]),
        "xtbb200_setProfile": (ci, [vp, ctypes.c_ulonglong]), "xtbb200_kernelClassCount": (ci, []),
        "xtbb200_kernelClassName": (cp, [ci]),
        "xtbb200_getKernelTimes": (ci, [vp, ctypes.POINTER(ctypes.c_float), _c_int_p, _c_int_p, _c_int_p]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _LIB = L
    return L


EXPORTED_SYMBOLS = None  # filled lazily by tests from the headers


def device_count():
    return load_library().xtbb200_deviceCount()


def symmetric_eigen(a, vectors=True, device=0):
    """Eigen-decomposition of a dense symmetric matrix with the engine's large-matrix solver (xtbb200_symmetricEigen):
    returns (w ascending, Z or None, {"sytrd_ms", "dc_ms", "backtransform_ms"})."""
    L = load_library()
    a = np.asfortranarray(a, dtype=np.float64)
    n = a.shape[0]
    w = np.empty(n)
    z = np.empty((n, n), order="F") if vectors else None
    ms = (ctypes.c_float * 3)()
    rc = L.xtbb200_symmetricEigen(n, _dp(a), _dp(w), _dp(z) if vectors else None, ms, device)
    if rc != 0:
        raise XTBException("xtbb200_symmetricEigen failed (%d): %s" % (rc, L.xtbb200_lastError().decode()))
    return w, z, {"sytrd_ms": ms[0], "dc_ms": ms[1], "backtransform_ms": ms[2]}


def _dp(a):
    return a.ctypes.data_as(_c_double_p)


def _ip(a):
    return a.ctypes.data_as(_c_int_p)


class Param(enum.Enum):
    """Parametrisations selectable through the API (include/xtb.h:154-180)."""
    GFN2xTB = "xtb_loadGFN2xTB"
    GFN1xTB = "xtb_loadGFN1xTB"
    GFN0xTB = "xtb_loadGFN0xTB"
    GFNFF = "xtb_loadGFNFF"


VERBOSITY_FULL, VERBOSITY_MINIMAL, VERBOSITY_MUTED = 2, 1, 0


class Environment:
    """Calculation environment (xtb_TEnvironment): error log + output unit + verbosity."""

    def __init__(self):
        self._lib = load_library()
        self._env = ctypes.c_void_p(self._lib.xtb_newEnvironment())

    def __del__(self):
        try:
            self._lib.xtb_delEnvironment(ctypes.byref(self._env))
        except Exception:
            pass

    def check(self):
        return self._lib.xtb_checkEnvironment(self._env)

    def get_error(self, message=None):
        buf = ctypes.create_string_buffer(2048)
        n = ctypes.c_int(2048)
        self._lib.xtb_getError(self._env, buf, ctypes.byref(n))
        msg = buf.value.decode()
        return msg if message is None else "%s:\n%s" % (message, msg)

    def raise_on_error(self, message):
        if self.check() != 0:
            raise XTBException(self.get_error(message))

    def set_output(self, filename):
        self._lib.xtb_setOutput(self._env, filename.encode())
        self.raise_on_error("Failed to open output file")

    def release_output(self):
        self._lib.xtb_releaseOutput(self._env)

    def set_verbosity(self, verbosity):
        self._lib.xtb_setVerbosity(self._env, int(verbosity))


class Molecule(Environment):
    """Molecular structure data (xtb_TMolecule); positions in Bohr."""

    def __init__(self, numbers, positions, charge=None, uhf=None, lattice=None, periodic=None):
        Environment.__init__(self)
        self._numbers = np.ascontiguousarray(numbers, dtype=np.int32)
        pos = np.ascontiguousarray(positions, dtype=np.float64)
        if pos.size != 3 * self._numbers.size:
            raise ValueError("Dimension missmatch between numbers and positions")
        self._natoms = int(self._numbers.size)
        n = ctypes.c_int(self._natoms)
        chrg = ctypes.c_double(charge) if charge is not None else None
        u = ctypes.c_int(uhf) if uhf is not None else None
        per = None
        if periodic is not None:
            per = (ctypes.c_bool * 3)(*[bool(p) for p in periodic])
        lat = _dp(np.ascontiguousarray(lattice, dtype=np.float64)) if lattice is not None else None
        self._mol = ctypes.c_void_p(self._lib.xtb_newMolecule(
            self._env, ctypes.byref(n), _ip(self._numbers), _dp(pos),
            ctypes.byref(chrg) if chrg is not None else None, ctypes.byref(u) if u is not None else None,
            lat, ctypes.cast(per, ctypes.c_void_p) if per is not None else None))
        self.raise_on_error("Setup of molecular structure failed")

    def __del__(self):
        try:
            self._lib.xtb_delMolecule(ctypes.byref(self._mol))
        except Exception:
            pass
        Environment.__del__(self)

    def __len__(self):
        return self._natoms

    def update(self, positions, lattice=None):
        pos = np.ascontiguousarray(positions, dtype=np.float64)
        if pos.size != 3 * self._natoms:
            raise ValueError("Dimension missmatch for positions")
        self._lib.xtb_updateMolecule(self._env, self._mol, _dp(pos), None)
        self.raise_on_error("Update of molecular structure failed")


class Results(Environment):
    """Single point results (xtb_TResults); also carries the restart wavefunction."""

    def __init__(self, res):
        Environment.__init__(self)
        if isinstance(res, Results):
            self._res = ctypes.c_void_p(self._lib.xtb_copyResults(res._res))
            self._natoms = res._natoms
        else:
            self._res = ctypes.c_void_p(self._lib.xtb_newResults())
            self._natoms = len(res)

    def __del__(self):
        try:
            self._lib.xtb_delResults(ctypes.byref(self._res))
        except Exception:
            pass
        Environment.__del__(self)

    def __len__(self):
        return self._natoms

    def _scalar(self, fn, msg):
        v = ctypes.c_double(0.0)
        fn(self._env, self._res, ctypes.byref(v))
        self.raise_on_error(msg)
        return v.value

    def _array(self, fn, shape, msg):
        a = np.zeros(shape, dtype=np.float64)
        fn(self._env, self._res, _dp(a))
        self.raise_on_error(msg)
        return a

    def get_energy(self):
        return self._scalar(self._lib.xtb_getEnergy, "Energy is not available")

    def get_info(self):
        """HOMO-LUMO gap [eV] and number of SCC iterations of the stored single point (xtbb200_getResultsInfo; the
        reference prints both but xtb.h has no getter for them)."""
        gap, nit = ctypes.c_double(0.0), ctypes.c_int(0)
        if self._lib.xtbb200_getResultsInfo(self._res, ctypes.byref(gap), ctypes.byref(nit)) != 0:
            raise XTBException(self._lib.xtbb200_lastError().decode())
        return dict(hl_gap=gap.value, niter=nit.value)

    def get_gradient(self):
        return self._array(self._lib.xtb_getGradient, (self._natoms, 3), "Gradient is not available")

    def get_pc_gradient(self, npc):
        """Gradient on the external point charges of the last single point (xtb_getPCGradient), shape (npc, 3)."""
        return self._array(self._lib.xtb_getPCGradient, (int(npc), 3), "Gradient on external charges is not available")

    def get_virial(self):
        return self._array(self._lib.xtb_getVirial, (3, 3), "Virial is not available")

    def get_dipole(self):
        return self._array(self._lib.xtb_getDipole, (3,), "Dipole is not available")

    def get_charges(self):
        return self._array(self._lib.xtb_getCharges, (self._natoms,), "Charges are not available")

    def get_bond_orders(self):
        return self._array(self._lib.xtb_getBondOrders, (self._natoms, self._natoms), "Bond orders are not available")

    def get_number_of_orbitals(self):
        n = ctypes.c_int(0)
        self._lib.xtb_getNao(self._env, self._res, ctypes.byref(n))
        self.raise_on_error("Number of orbitals is not available")
        return n.value

    def get_orbital_eigenvalues(self):
        return self._array(self._lib.xtb_getOrbitalEigenvalues, (self.get_number_of_orbitals(),), "Not available")

    def get_orbital_occupations(self):
        return self._array(self._lib.xtb_getOrbitalOccupations, (self.get_number_of_orbitals(),), "Not available")

    def get_orbital_coefficients(self):
        n = self.get_number_of_orbitals()
        return self._array(self._lib.xtb_getOrbitalCoefficients, (n, n), "Not available")


class Calculator(Molecule):
    """Single point calculator (xtb_TCalculator) bound to one molecule.

    >>> calc = Calculator(Param.GFN2xTB, numbers, positions)
    >>> res = calc.singlepoint()
    >>> res.get_energy(), res.get_gradient()
    """

    def __init__(self, param, numbers, positions, charge=None, uhf=None, lattice=None, periodic=None):
        Molecule.__init__(self, numbers, positions, charge, uhf, lattice, periodic)
        self._calc = ctypes.c_void_p(self._lib.xtb_newCalculator())
        getattr(self._lib, param.value)(self._env, self._mol, self._calc, None)
        self.raise_on_error("Could not load parametrisation data")

    def __del__(self):
        try:
            self._lib.xtb_delCalculator(ctypes.byref(self._calc))
        except Exception:
            pass
        Molecule.__del__(self)

    def set_accuracy(self, accuracy):
        self._lib.xtb_setAccuracy(self._env, self._calc, float(accuracy))
        self.raise_on_error("Could not set accuracy")

    def set_max_iterations(self, maxiter):
        self._lib.xtb_setMaxIter(self._env, self._calc, int(maxiter))
        self.raise_on_error("Could not set max iterations")

    def set_electronic_temperature(self, etemp):
        self._lib.xtb_setElectronicTemp(self._env, self._calc, float(etemp))
        self.raise_on_error("Could not set electronic temperature")

    def set_solvent(self, solvent=None, model=None, state=None, temperature=None, grid=None):
        """xtb_setSolvent / xtb_releaseSolvent: implicit solvation, model "gbsa" (default) or "alpb"; None releases it."""
        if solvent is None:
            self._lib.xtb_releaseSolvent(self._env, self._calc)
        else:
            st = ctypes.byref(ctypes.c_int(int(state))) if state is not None else None
            te = ctypes.byref(ctypes.c_double(float(temperature))) if temperature is not None else None
            gr = ctypes.byref(ctypes.c_int(int(grid))) if grid is not None else None
            self._lib.xtb_setSolvent(self._env, self._calc, str(solvent).encode(), st, te, gr,
                                     model.encode() if model is not None else None)
        self.raise_on_error("Failed to set solvent model")

    def set_external_charges(self, numbers, charges, positions):
        n = ctypes.c_int(len(numbers))
        nu = np.ascontiguousarray(numbers, dtype=np.int32)
        ch = np.ascontiguousarray(charges, dtype=np.float64)
        po = np.ascontiguousarray(positions, dtype=np.float64)
        self._lib.xtb_setExternalCharges(self._env, self._calc, ctypes.byref(n), _ip(nu), _dp(ch), _dp(po))
        self.raise_on_error("Failed to set external charges")
        self._npc = len(nu)

    def release_external_charges(self):
        self._lib.xtb_releaseExternalCharges(self._env, self._calc)
        self.raise_on_error("Failed to release external charges")
        self._npc = 0

    def hessian(self, res=None, step=None):
        """Numerical Hessian (xtb_hessian): central differences of analytical gradients, the 6 N displaced geometries run
        as one ensemble on the GPU.  Returns (hessian (3N, 3N), dipole gradient (3N, 3)).  `res`: results of a previous
        single point whose wavefunction is the SCC start of every point (None: EEQ guess)."""
        _res = Results(self) if res is None else res
        n3 = 3 * len(self)
        hess = np.zeros((n3, n3))
        dipgrad = np.zeros((n3, 3))
        st = ctypes.c_double(step) if step is not None else None
        self._lib.xtb_hessian(self._env, self._mol, self._calc, _res._res, _dp(hess), None,
                              ctypes.cast(ctypes.byref(st), _c_int_p) if st is not None else None, _dp(dipgrad), None)
        self.raise_on_error("Hessian calculation failed")
        return hess, dipgrad

    def singlepoint(self, res=None, copy=False):
        _res = Results(self) if res is None else (Results(res) if copy else res)
        self._lib.xtb_singlepoint(self._env, self._mol, self._calc, _res._res)
        self.raise_on_error("Single point calculation failed")
        return _res


RESULT_FIELDS = ["energy", "e_elec", "e_rep", "e_es", "e_aes", "e_axc", "e_disp", "hl_gap", "gnorm", "e_atm",
                 "dipole_x", "dipole_y", "dipole_z", "niter", "converged", "fail"]


class BatchCalculator:
    """GFN2-xTB single points + gradients for an ensemble of molecules in one call (include/xtb_b200.h)."""

    def __init__(self, numbers_list, charges=None, uhf=None, device=0, accuracy=1.0, etemp=300.0, maxiter=250,
                 guess="eeq"):
        self._lib = load_library()
        self.natoms = np.array([len(n) for n in numbers_list], dtype=np.int32)
        self.nmol = len(numbers_list)
        numbers = np.ascontiguousarray(np.concatenate([np.asarray(n, dtype=np.int32) for n in numbers_list]), dtype=np.int32)
        ch = np.ascontiguousarray(charges, dtype=np.float64) if charges is not None else None
        uh = np.ascontiguousarray(uhf, dtype=np.int32) if uhf is not None else None
        self._b = ctypes.c_void_p(self._lib.xtbb200_newBatch(self.nmol, _ip(self.natoms), _ip(numbers),
                                                            _dp(ch) if ch is not None else None,
                                                            _ip(uh) if uh is not None else None, int(device)))
        if not self._b:
            raise XTBException(self._lib.xtbb200_lastError().decode())
        self.total_atoms = int(self.natoms.sum())
        self.offsets = np.concatenate([[0], np.cumsum(self.natoms)])
        self.set_options(accuracy, etemp, maxiter, guess)

    def __del__(self):
        try:
            self._lib.xtbb200_delBatch(ctypes.byref(self._b))
        except Exception:
            pass

    def _check(self, rc):
        if rc < 0:
            raise XTBException(self._lib.xtbb200_lastError().decode())
        return rc

    def set_options(self, accuracy=1.0, etemp=300.0, maxiter=250, guess="eeq"):
        self._check(self._lib.xtbb200_setOptions(self._b, float(accuracy), float(etemp), int(maxiter),
                                                 {"eeq": 0, "sad": 1}[guess]))

    def set_wavefunction(self, qsh=None, dipm=None, qp=None):
        """Restart wavefunction for the next singlepoint(restart=True); None drops the stored one."""
        if qsh is None:
            self._check(self._lib.xtbb200_setWavefunction(self._b, None, None, None))
            return
        a, b, c = (np.ascontiguousarray(x, dtype=np.float64) for x in (qsh, dipm, qp))
        self._check(self._lib.xtbb200_setWavefunction(self._b, _dp(a), _dp(b), _dp(c)))

    def set_external_charges(self, npc=None, numbers=None, charges=None, positions=None):
        """Point charges per molecule (xtbb200_setExternalCharges): npc[nmol], concatenated numbers / charges / positions
        (bohr); None releases them.  Gradient on the charges after a single point: array("pcgradient")."""
        if npc is None:
            self._check(self._lib.xtbb200_setExternalCharges(self._b, None, None, None, None))
            return
        n = np.ascontiguousarray(npc, dtype=np.int32)
        nu = np.ascontiguousarray(numbers, dtype=np.int32)
        ch = np.ascontiguousarray(charges, dtype=np.float64)
        po = np.ascontiguousarray(positions, dtype=np.float64)
        assert len(nu) == int(n.sum()) == len(ch) and po.size == 3 * len(ch)
        self._check(self._lib.xtbb200_setExternalCharges(self._b, _ip(n), _ip(nu), _dp(ch), _dp(po)))

    def set_solvent(self, solvent=None, model="gbsa", kernel=None, state=1, temperature=298.15, ion_strength=0.0, ion_rad=0.0):
        """Implicit solvation for all molecules (xtbb200_setSolvent): model "gbsa" or "alpb"; kernel "still" / "p16"
        (default: what xtb_setSolvent selects for the model); None releases it.  Energy parts after a single point:
        array("solvation").reshape(nmol, 5) = g_solv, g_born, g_sasa, g_hb, g_shift."""
        if solvent is None:
            self._check(self._lib.xtbb200_setSolvent(self._b, None, 0, 2, 1, 298.15, 0.0, 0.0))
            return
        if kernel is None:
            kernel = "still" if model == "gbsa" else "p16"
        self._check(self._lib.xtbb200_setSolvent(self._b, str(solvent).encode(), 1 if model == "alpb" else 0,
                                                 {"still": 1, "p16": 2}[kernel], int(state), float(temperature),
                                                 float(ion_strength), float(ion_rad)))

    def md_init(self, positions, tstep_fs=4.0, temperature=298.15, thermostat=True, shake=2, hmass=4, velocities=None,
                seed=1, restart=False):
        """Start molecular dynamics of all molecules as independent replicas (xtbb200_mdInit; defaults as xtb's $md block:
        4 fs, 298.15 K, Berendsen thermostat, SHAKE on all bonds, hydrogen mass 4 amu).  Returns the number of constraints."""
        pos = np.ascontiguousarray(positions, dtype=np.float64)
        assert pos.size == 3 * self.total_atoms
        vel = np.ascontiguousarray(velocities, dtype=np.float64) if velocities is not None else None
        self._check(self._lib.xtbb200_mdInit(self._b, _dp(pos), float(tstep_fs), float(temperature), int(bool(thermostat)),
                                             int(shake), int(hmass), _dp(vel) if vel is not None else None, int(seed),
                                             int(bool(restart))))
        return int(self._lib.xtbb200_mdConstraints(self._b))

    def md_run(self, nsteps):
        """Propagate nsteps MD steps (xtbb200_mdRun).  Returns dict(steps, epot, ekin, temperature) with [steps][nmol] arrays;
        positions / velocities afterwards: array("xyz") / array("velocities")."""
        ep, ek, te = (np.zeros((int(nsteps), self.nmol)) for _ in range(3))
        n = self._check(self._lib.xtbb200_mdRun(self._b, int(nsteps), _dp(ep), _dp(ek), _dp(te)))
        return dict(steps=n, epot=ep[:n], ekin=ek[:n], temperature=te[:n])

    def set_solver(self, mode):
        """0: automatic, 1: scalar Jacobi kernel, 2: X from eig(S), 3: two-level Jacobi for large matrices, 4: batched
        Householder + D&C for ensemble-sized matrices, 5: Householder + D&C one matrix at a time (include/xtb_b200.h)."""
        self._check(self._lib.xtbb200_setSolver(self._b, int(mode)))

    def set_shard(self, rank, nranks, unique_id=None):
        """Make this batch rank `rank` of `nranks` processes that work on the SAME molecules (one per GPU): integral /
        gradient blocks are split, contributions summed by NCCL allreduce (include/xtb_b200.h).  Collective call."""
        uid = bytes(unique_id) if unique_id is not None else None
        self._check(self._lib.xtbb200_setShard(self._b, int(rank), int(nranks), uid, len(uid) if uid else 0))

    def comm_ms(self):
        ms = ctypes.c_float(0)
        self._check(self._lib.xtbb200_getCommTime(self._b, ctypes.byref(ms)))
        return ms.value

    def singlepoint(self, positions, restart=False):
        """positions: (total_atoms, 3) Bohr. Returns dict(energies, gradients, charges, niter, status)."""
        pos = np.ascontiguousarray(positions, dtype=np.float64)
        if pos.size != 3 * self.total_atoms:
            raise ValueError("Dimension missmatch for positions")
        e = np.zeros(self.nmol)
        g = np.zeros((self.total_atoms, 3))
        q = np.zeros(self.total_atoms)
        it = np.zeros(self.nmol, dtype=np.int32)
        st = np.zeros(self.nmol, dtype=np.int32)
        rc = self._check(self._lib.xtbb200_singlepoint(self._b, _dp(pos), int(bool(restart)), _dp(e), _dp(g), _dp(q),
                                                       _ip(it), _ip(st)))
        return dict(energies=e, gradients=g, charges=q, niter=it, status=st, all_converged=(rc == 0))

    def singlepoint_device(self, positions_ptr, restart=False):
        """positions_ptr: integer device address of (total_atoms,3) float64 on this batch's GPU."""
        return self._check(self._lib.xtbb200_singlepointDevice(self._b, ctypes.c_void_p(int(positions_ptr)),
                                                               int(bool(restart)), None))

    def result(self, imol):
        out = np.zeros(16)
        self._check(self._lib.xtbb200_getResult(self._b, int(imol), _dp(out)))
        return dict(zip(RESULT_FIELDS, out.tolist()))

    def counts(self, imol):
        a, b, c = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        self._check(self._lib.xtbb200_getCounts(self._b, int(imol), ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        return dict(ndp=a.value, nqp=b.value, sweeps=c.value)

    def electrons(self, imol):
        a, b = ctypes.c_int(), ctypes.c_int()
        self._check(self._lib.xtbb200_getElectrons(self._b, int(imol), ctypes.byref(a), ctypes.byref(b)))
        return dict(nel=a.value, nopen=b.value)

    DECISION_COUNTS = ("sdq_shell_pairs", "sdq_primitive_pairs", "grad_shell_pairs", "grad_primitive_pairs", "cn_pairs",
                       "cn_d4_pairs", "repulsion_pairs", "d4_matrix_entries", "d4_gradient_pairs", "atm_triples", "wbo_pairs", "born_stale_pairs")

    def set_decision_counting(self, on=True):
        """Diagnostic: let the kernels count their cutoff / screening decisions during the next single points."""
        self._check(self._lib.xtbb200_setDecisionCounting(self._b, int(bool(on))))

    def decision_counts(self, imol):
        out = (ctypes.c_longlong * 12)()
        self._check(self._lib.xtbb200_getDecisionCounts(self._b, int(imol), out))
        return dict(zip(self.DECISION_COUNTS, list(out)))

    def sizes(self, imol):
        a, b, c = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        self._check(self._lib.xtbb200_getSizes(self._b, int(imol), ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        return dict(natoms=a.value, nshell=b.value, nao=c.value)

    def offsets_of(self, imol):
        o = (ctypes.c_longlong * 5)()
        self._check(self._lib.xtbb200_getOffsets(self._b, int(imol), o))
        return dict(atom=o[0], shell=o[1], ao=o[2], mat=o[3], nn=o[4])

    def array(self, name):
        n = ctypes.c_size_t(0)
        self._check(self._lib.xtbb200_getArrayLen(self._b, name.encode(), ctypes.byref(n)))
        out = np.zeros(n.value)
        self._check(self._lib.xtbb200_getArray(self._b, name.encode(), _dp(out), n.value))
        return out

    def matrix(self, name, imol, planes=1):
        """nao x nao matrix (column-major on the device == symmetric here) of one molecule."""
        full = self.array(name)
        off = self.offsets_of(imol)
        nao = self.sizes(imol)["nao"]
        tot = full.size // planes
        out = [full[p * tot + off["mat"]: p * tot + off["mat"] + nao * nao].reshape(nao, nao).T for p in range(planes)]
        return out[0] if planes == 1 else np.stack(out)

    def mask(self):
        n = ctypes.c_size_t(0)
        self._check(self._lib.xtbb200_getArrayLen(self._b, b"S", ctypes.byref(n)))
        out = np.zeros(n.value, dtype=np.uint8)
        self._check(self._lib.xtbb200_getMask(self._b, out.ctypes.data, n.value))
        return out

    def timings(self):
        t = (ctypes.c_float * 3)()
        n = ctypes.c_int()
        self._check(self._lib.xtbb200_getTimings(self._b, t, ctypes.byref(n)))
        f = ctypes.c_float()
        self._check(self._lib.xtbb200_getFactorisationTime(self._b, ctypes.byref(f)))
        return dict(setup_ms=t[0], scc_ms=t[1], grad_ms=t[2], launches=n.value, factorisation_ms=f.value)

    def call_ms(self):
        """Device time of the last host-buffer singlepoint call, H2D of positions through D2H of results."""
        t = ctypes.c_float()
        self._check(self._lib.xtbb200_getCallTime(self._b, ctypes.byref(t)))
        return t.value

    def set_profile(self, classes="all"):
        """Bracket the launches of the given kernel classes (names, ids, "all" or None) with CUDA events."""
        n = self._lib.xtbb200_kernelClassCount()
        names = [self._lib.xtbb200_kernelClassName(i).decode() for i in range(n)]
        if classes is None:
            mask = 0
        elif classes == "all":
            mask = (1 << n) - 1
        else:
            mask = 0
            for c in classes:
                mask |= 1 << (names.index(c) if isinstance(c, str) else int(c))
        self._check(self._lib.xtbb200_setProfile(self._b, mask))

    def kernel_times(self):
        """{class name: dict(ms, timed_launches, launches)} accumulated since set_profile, plus scc_passes."""
        n = self._lib.xtbb200_kernelClassCount()
        ms = (ctypes.c_float * n)()
        tl = (ctypes.c_int * n)()
        nl = (ctypes.c_int * n)()
        sp = ctypes.c_int()
        self._check(self._lib.xtbb200_getKernelTimes(self._b, ms, tl, nl, ctypes.byref(sp)))
        out = {self._lib.xtbb200_kernelClassName(i).decode(): dict(ms=ms[i], timed_launches=tl[i], launches=nl[i])
               for i in range(n)}
        return out, sp.value

    def device_bytes(self):
        return int(self._lib.xtbb200_getDeviceBytes(self._b))
