"""B200-native GFN2-xTB single point + gradient engine behind the xtb calculator API.

The compute path is hand-written CUDA (sm_100a) in xtb_b200/csrc, exposed through the C ABI of
include/xtb.h / include/xtb_b200.h (xtb_b200/lib/libxtb_b200.so).  This package is the Python
host-side mirror of that ABI; there is no CPU fallback.
"""
from .interface import (BatchCalculator, Calculator, Environment, Molecule, Param, Results, XTBException,
                        device_count, library_path, load_library, symmetric_eigen)

__all__ = ["BatchCalculator", "Calculator", "Environment", "Molecule", "Param", "Results", "XTBException",
           "device_count", "library_path", "load_library", "symmetric_eigen"]
