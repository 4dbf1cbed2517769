/* xtb.h -- C ABI of the B200-native GFN2-xTB engine.
 *
 * Drop-in boundary: the entry points below are exactly the ones the reference exports from
 * include/xtb.h (grimme-lab/xtb v6.7.1, API version 2.0.0; Fortran side in src/api/*.f90).
 * Each declaration cites the reference interface it replaces.  Only the GFN2-xTB single point
 * path is implemented on the GPU; the other calculators / models exist as symbols and raise an
 * error on the environment (never a missing symbol, never a CPU fallback).
 *
 * Conventions (reference: include/xtb.h:120-128,211-226,272-340; src/api/molecule.f90:45-92):
 *   - positions[natoms][3] in Bohr, energies in Hartree, gradients Hartree/Bohr
 *   - scalars natoms/charge/uhf by pointer (NULL = 0), accuracy/maxiter/temperature by value
 *   - no return codes: failures are logged on the environment, poll xtb_checkEnvironment()
 *   - xtb_del* take the address of the handle, free it and set it to NULL
 */
#pragma once
#ifdef __cplusplus
extern "C" {
#else
#include <stdbool.h>
#endif

#define XTB_API_VERSION 20000        /* include/xtb.h:25 */
#define XTB_VERBOSITY_FULL    2      /* include/xtb.h:32-34 */
#define XTB_VERBOSITY_MINIMAL 1
#define XTB_VERBOSITY_MUTED   0

/* opaque handles, include/xtb.h:55-64 */
typedef struct _xtb_TEnvironment* xtb_TEnvironment;
typedef struct _xtb_TMolecule* xtb_TMolecule;
typedef struct _xtb_TCalculator* xtb_TCalculator;
typedef struct _xtb_TResults* xtb_TResults;

#ifndef __cplusplus
/* generic destructor, include/xtb.h:37-42 */
#define xtb_delete(ptr) _Generic((ptr), \
       xtb_TEnvironment: xtb_delEnvironment, \
          xtb_TMolecule: xtb_delMolecule, \
        xtb_TCalculator: xtb_delCalculator, \
           xtb_TResults: xtb_delResults \
                                 )(&ptr)
#endif

/* include/xtb.h:71-72 ; src/api/version.f90 */
int xtb_getAPIVersion(void);

/* environment: include/xtb.h:79-113 ; src/api/environment.f90:46-246 */
xtb_TEnvironment xtb_newEnvironment(void);
void xtb_delEnvironment(xtb_TEnvironment* env);
int xtb_checkEnvironment(xtb_TEnvironment env);
void xtb_showEnvironment(xtb_TEnvironment env, const char* message);
void xtb_getError(xtb_TEnvironment env, char* buffer, const int* buffersize);
void xtb_setOutput(xtb_TEnvironment env, const char* filename);
void xtb_releaseOutput(xtb_TEnvironment env);
void xtb_setVerbosity(xtb_TEnvironment env, int verbosity);

/* molecule: include/xtb.h:120-141 ; src/api/molecule.f90:45-170 */
xtb_TMolecule xtb_newMolecule(xtb_TEnvironment env, const int* natoms, const int* numbers,
                              const double* positions, const double* charge, const int* uhf,
                              const double* lattice, const bool* periodic);
void xtb_delMolecule(xtb_TMolecule* mol);
void xtb_updateMolecule(xtb_TEnvironment env, xtb_TMolecule mol, const double* positions,
                        const double* lattice);

/* calculator: include/xtb.h:146-252 ; src/api/calculator.f90:57-696, src/api/interface.f90:40-156 */
xtb_TCalculator xtb_newCalculator(void);
void xtb_delCalculator(xtb_TCalculator* calc);
void xtb_loadGFN0xTB(xtb_TEnvironment env, xtb_TMolecule mol, xtb_TCalculator calc, char* filename);
void xtb_loadGFN1xTB(xtb_TEnvironment env, xtb_TMolecule mol, xtb_TCalculator calc, char* filename);
void xtb_loadGFN2xTB(xtb_TEnvironment env, xtb_TMolecule mol, xtb_TCalculator calc, char* filename);
void xtb_loadGFNFF(xtb_TEnvironment env, xtb_TMolecule mol, xtb_TCalculator calc, char* filename);
void xtb_setSolvent(xtb_TEnvironment env, xtb_TCalculator calc, char* solvent, int* state,
                    double* temp, int* grid, char* solvent_model);
void xtb_releaseSolvent(xtb_TEnvironment env, xtb_TCalculator calc);
void xtb_setExternalCharges(xtb_TEnvironment env, xtb_TCalculator calc, int* n, int* numbers,
                            double* charges, double* positions);
void xtb_releaseExternalCharges(xtb_TEnvironment env, xtb_TCalculator calc);
void xtb_setAccuracy(xtb_TEnvironment env, xtb_TCalculator calc, double accuracy);
void xtb_setMaxIter(xtb_TEnvironment env, xtb_TCalculator calc, int iterations);
void xtb_setElectronicTemp(xtb_TEnvironment env, xtb_TCalculator calc, double temperature);
void xtb_cpcmx_calc(xtb_TEnvironment env, xtb_TMolecule mol, xtb_TCalculator calc, xtb_TResults res);
void xtb_singlepoint(xtb_TEnvironment env, xtb_TMolecule mol, xtb_TCalculator calc, xtb_TResults res);
/* Argument order as in the reference's C header (include/xtb.h:244-252).  The reference's Fortran side declares
 * (hessian, step_size, atom_index_list) (src/api/interface.f90:310-311), so its binary reads the 6th pointer as the step
 * and the 7th as the list; this library follows the header.  Portable callers pass NULL for both (all atoms, 0.005 bohr);
 * a non-NULL atom_index_list is refused with an environment error, a non-NULL step_size is read as a double. */
void xtb_hessian(xtb_TEnvironment env, xtb_TMolecule mol, xtb_TCalculator calc, xtb_TResults res,
                 double* hessian, int* atom_index_list, int* step_size, double* dipole_gradient,
                 double* polarizability_gradient);

/* results: include/xtb.h:259-340 ; src/api/results.f90:37-504 */
xtb_TResults xtb_newResults(void);
void xtb_delResults(xtb_TResults* res);
xtb_TResults xtb_copyResults(xtb_TResults res);
void xtb_getEnergy(xtb_TEnvironment env, xtb_TResults res, double* energy);
void xtb_getSolvationEnergy(xtb_TEnvironment env, xtb_TResults res, double* energy);
void xtb_getGradient(xtb_TEnvironment env, xtb_TResults res, double* gradient);
void xtb_getPCGradient(xtb_TEnvironment env, xtb_TResults res, double* gradient);
void xtb_getVirial(xtb_TEnvironment env, xtb_TResults res, double* virial);
void xtb_getDipole(xtb_TEnvironment env, xtb_TResults res, double* dipole);
void xtb_getCharges(xtb_TEnvironment env, xtb_TResults res, double* charges);
void xtb_getBondOrders(xtb_TEnvironment env, xtb_TResults res, double* wbo);
void xtb_getNao(xtb_TEnvironment env, xtb_TResults res, int* nao);
void xtb_getOrbitalEigenvalues(xtb_TEnvironment env, xtb_TResults res, double* emo);
void xtb_getOrbitalOccupations(xtb_TEnvironment env, xtb_TResults res, double* focc);
void xtb_getOrbitalCoefficients(xtb_TEnvironment env, xtb_TResults res, double* c);

#ifdef __cplusplus
}
#endif
