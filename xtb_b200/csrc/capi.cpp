// C ABI of the B200 engine: the reference's include/xtb.h entry points (opaque handles, error log
// on the environment) plus the batched extension of include/xtb_b200.h.
//
// Reference semantics followed (file:line in the reference):
//   src/api/environment.f90:46-246   environment, error stack, verbosity, output unit
//   src/api/molecule.f90:45-170      molecule creation / update / validation
//   src/api/calculator.f90:57-696    calculator loading and setters (clamps, error messages)
//   src/api/interface.f90:40-156     xtb_singlepoint (cache check, restart from res%chk)
//   src/api/results.f90:37-504       getters (errors when data is missing)
#include <cmath>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/xtb.h"
#include "../../include/xtb_b200.h"
#include "engine.h"
#include "solvation.h"
#include "md_host.h"

using xtbb200::Batch;
using xtbb200::MolResult;

namespace {
thread_local std::string g_last_error;
}

struct _xtb_TEnvironment {
    std::vector<std::string> log;   // message stack; error pending iff has_error
    bool has_error = false;
    int verbosity = 1;              // src/api/environment.f90:56
    FILE* unit = nullptr;           // nullptr -> stdout
    bool owns_unit = false;
    void error(const std::string& msg, const char* source) {
        log.push_back(std::string("-") + source + "- " + msg);
        has_error = true;
    }
    void warning(const std::string& msg, const char* source) { log.push_back(std::string("#") + source + "# " + msg); }
    FILE* out() { return unit ? unit : stdout; }
};
struct _xtb_TMolecule {
    int n = 0;
    std::vector<int> at;
    std::vector<double> xyz;
    double chrg = 0;
    int uhf = 0;
};
struct _xtb_TCalculator {
    std::shared_ptr<Batch> batch;   // engine for the element list the calculator was loaded with
    std::vector<int> at;
    double chrg = 0;
    int uhf = 0;
    double accuracy = 1.0, etemp = 300.0;
    int maxiter = 250;
    bool loaded = false;
    // external point charges (src/type/pcem.f90): element numbers (-> hardness), charges, positions in bohr
    std::vector<int> pc_numbers;
    std::vector<double> pc_q, pc_xyz;
    // implicit solvation (xtb_setSolvent): applied to the engine at once, kept here for the Hessian ensemble
    bool solv_on = false, solv_alpb = false;
    std::string solv_name;
    int solv_kernel = 2, solv_state = 1;
    double solv_temp = 298.15;
    // ensemble of displaced copies of the molecule for xtb_hessian (created on first use, reused)
    std::shared_ptr<Batch> hess_batch;
    int hess_copies = 0;
};
struct _xtb_TResults {
    bool has_energy = false, has_gradient = false, has_wfn = false, has_dipole = false, has_pcgradient = false;
    std::vector<double> pcgradient;
    int n = 0, nao = 0, nshell = 0;
    double energy = 0, egap = 0;
    std::vector<double> gradient, q, qsh, dipm, qp, wbo, emo, focc, C;
    double dipole[3] = {0, 0, 0};
    int niter = 0;
};

namespace {

bool check_env(xtb_TEnvironment env) { return env != nullptr; }

void unsupported(xtb_TEnvironment env, const char* what, const char* source) {
    if (!check_env(env)) return;
    env->error(std::string(what) + " is not available in the B200 GFN2-xTB engine", source);
}

}  // namespace

extern "C" {

int xtb_getAPIVersion(void) { return XTB_API_VERSION; }

// ---------------------------------------------------------------- environment
xtb_TEnvironment xtb_newEnvironment(void) { return new _xtb_TEnvironment(); }
void xtb_delEnvironment(xtb_TEnvironment* env) {
    if (!env || !*env) return;
    if ((*env)->owns_unit && (*env)->unit) fclose((*env)->unit);
    delete *env;
    *env = nullptr;
}
int xtb_checkEnvironment(xtb_TEnvironment env) {
    if (!env) return 2;
    return env->has_error ? 1 : 0;
}
void xtb_showEnvironment(xtb_TEnvironment env, const char* message) {
    if (!env) return;
    FILE* f = env->out();
    if (!env->log.empty()) {
        if (message) fprintf(f, "%s\n", message);
        for (auto it = env->log.rbegin(); it != env->log.rend(); ++it) fprintf(f, "%s\n", it->c_str());
        fflush(f);
    }
    env->log.clear();
    env->has_error = false;
}
void xtb_getError(xtb_TEnvironment env, char* buffer, const int* buffersize) {
    if (!env || !buffer) return;
    int max_length = buffersize ? *buffersize : 0;
    std::string msg;
    for (auto it = env->log.rbegin(); it != env->log.rend(); ++it) {
        if (!msg.empty()) msg += "\n";
        msg += *it;
    }
    if (max_length > 0) {
        int n = std::min<int>((int)msg.size(), max_length - 1);
        std::memcpy(buffer, msg.data(), n);
        buffer[n] = '\0';
    }
    env->log.clear();
    env->has_error = false;
}
void xtb_setOutput(xtb_TEnvironment env, const char* filename) {
    if (!env) return;
    if (env->owns_unit && env->unit) fclose(env->unit);
    env->unit = nullptr; env->owns_unit = false;
    if (!filename) return;
    if (std::strcmp(filename, "-") == 0 || std::strcmp(filename, "STDOUT") == 0) return;
    FILE* f = fopen(filename, "w");
    if (!f) { env->error(std::string("Could not open output file '") + filename + "'", "xtb_api_setOutput"); return; }
    env->unit = f; env->owns_unit = true;
}
void xtb_releaseOutput(xtb_TEnvironment env) {
    if (!env) return;
    if (env->owns_unit && env->unit) fclose(env->unit);
    env->unit = nullptr; env->owns_unit = false;
}
void xtb_setVerbosity(xtb_TEnvironment env, int verbosity) {
    if (!env) return;
    env->verbosity = verbosity < 0 ? 0 : (verbosity > 2 ? 2 : verbosity);
}

// ---------------------------------------------------------------- molecule
static bool verify_molecule(const _xtb_TMolecule& m) {  // src/api/utils.f90:63-72
    for (int i = 0; i < m.n; ++i)
        for (int j = 0; j < i; ++j) {
            double r2 = 0;
            for (int k = 0; k < 3; ++k) { double d = m.xyz[3 * i + k] - m.xyz[3 * j + k]; r2 += d * d; }
            if (std::sqrt(r2) < 1.0e-9) return false;
        }
    return true;
}
xtb_TMolecule xtb_newMolecule(xtb_TEnvironment env, const int* natoms, const int* numbers, const double* positions,
                              const double* charge, const int* uhf, const double* lattice, const bool* periodic) {
    const char* source = "xtb_api_newMolecule";
    if (!check_env(env)) return nullptr;
    if (!natoms || *natoms <= 0) { env->error("Number of atoms must be positive", source); return nullptr; }
    if (!numbers || !positions) { env->error("Atomic numbers and positions are required", source); return nullptr; }
    for (int i = 0; i < *natoms; ++i)
        if (numbers[i] < 1 || numbers[i] > 118) { env->error("Invalid atomic number present", source); return nullptr; }
    if (periodic && (periodic[0] || periodic[1] || periodic[2])) {
        env->error("Periodic boundary conditions are not available in the B200 GFN2-xTB engine", source);
        return nullptr;
    }
    (void)lattice;
    auto* m = new _xtb_TMolecule();
    m->n = *natoms;
    m->at.assign(numbers, numbers + m->n);
    m->xyz.assign(positions, positions + 3 * m->n);
    m->chrg = charge ? *charge : 0.0;
    m->uhf = uhf ? *uhf : 0;
    if (!verify_molecule(*m)) {
        delete m;
        env->error("Could not generate molecular structure", source);
        return nullptr;
    }
    return m;
}
void xtb_delMolecule(xtb_TMolecule* mol) {
    if (!mol || !*mol) return;
    delete *mol;
    *mol = nullptr;
}
void xtb_updateMolecule(xtb_TEnvironment env, xtb_TMolecule mol, const double* positions, const double* lattice) {
    const char* source = "xtb_api_updateMolecule";
    if (!check_env(env)) return;
    if (!mol) { env->error("Molecular structure data is not allocated", source); return; }
    if (!positions) { env->error("Positions are required", source); return; }
    (void)lattice;
    std::vector<double> old = mol->xyz;
    mol->xyz.assign(positions, positions + 3 * mol->n);
    if (!verify_molecule(*mol)) {
        env->error("Could not update molecular structure", source);
        mol->xyz = old;
    }
}

// ---------------------------------------------------------------- calculator
xtb_TCalculator xtb_newCalculator(void) { return new _xtb_TCalculator(); }
void xtb_delCalculator(xtb_TCalculator* calc) {
    if (!calc || !*calc) return;
    delete *calc;
    *calc = nullptr;
}
void xtb_loadGFN2xTB(xtb_TEnvironment env, xtb_TMolecule mol, xtb_TCalculator calc, char* filename) {
    const char* source = "xtb_api_loadGFN2xTB";
    if (!check_env(env)) return;
    if (!mol) { env->error("Molecular structure data is not allocated", source); return; }
    if (!calc) { env->error("Calculator object is not allocated", source); return; }
    // parameter files are not parsed: the compiled-in GFN2-xTB tables are the reference's own fallback
    // (src/xtb/calculator.f90:147-160); a custom file would change the arithmetic, so refuse it loudly
    if (filename && filename[0] != '\0') {
        if (std::strstr(filename, "param_gfn2-xtb.txt") == nullptr) {
            env->error("Custom parameter files are not supported, only the built-in GFN2-xTB parametrisation", source);
            return;
        }
        // the distributed param_gfn2-xtb.txt holds the same numbers as the compiled-in tables; it is not read, say so
        env->warning(std::string("Parameter file '") + filename + "' is not read: the built-in GFN2-xTB tables "
                     "(identical to the distributed param_gfn2-xtb.txt) are used", source);
    }
    try {
        double chrg = mol->chrg;
        int uhf = mol->uhf;
        calc->batch = std::make_shared<Batch>(1, &mol->n, mol->at.data(), &chrg, &uhf, 0);
        calc->at = mol->at; calc->chrg = chrg; calc->uhf = uhf;
        calc->accuracy = 1.0; calc->etemp = 300.0; calc->maxiter = 250;   // src/xtb/calculator.f90:138-145
        calc->loaded = true;
    } catch (const std::exception& e) {
        calc->loaded = false;
        calc->batch.reset();
        env->error(std::string("Could not load parametrisation data: ") + e.what(), source);
    }
}
void xtb_loadGFN0xTB(xtb_TEnvironment env, xtb_TMolecule, xtb_TCalculator, char*) { unsupported(env, "GFN0-xTB", "xtb_api_loadGFN0xTB"); }
void xtb_loadGFN1xTB(xtb_TEnvironment env, xtb_TMolecule, xtb_TCalculator, char*) { unsupported(env, "GFN1-xTB", "xtb_api_loadGFN1xTB"); }
void xtb_loadGFNFF(xtb_TEnvironment env, xtb_TMolecule, xtb_TCalculator, char*) { unsupported(env, "GFN-FF", "xtb_api_loadGFNFF"); }
// src/api/calculator.f90:341-445: solvent name, optional reference state (1 gsolv, 2 reference [X=1], 3 mol1bar),
// temperature [K], angular grid size (only the default 230 point Lebedev grid is built in), model "gbsa" (default; Still
// kernel, GBSA parameter sets) or "alpb" (P16 kernel, ALPB parameter sets).  COSMO, TM-COSMO and CPCM-X are refused.
void xtb_setSolvent(xtb_TEnvironment env, xtb_TCalculator calc, char* solvent, int* state, double* temperature, int* grid, char* model) {
    const char* source = "xtb_api_setSolvent";
    if (!check_env(env)) return;
    if (!calc) { env->error("Singlepoint calculator is not allocated", source); return; }
    if (!calc->loaded || !calc->batch) { env->error("No calculator loaded to add solvation model", source); return; }
    if (!solvent) { env->error("Solvent name is not given", source); return; }
    const int st = state ? *state : 1;
    const double temp = temperature ? *temperature : 298.15;
    const int nang = grid ? *grid : 230;
    const std::string mdl = model ? std::string(model) : std::string("gbsa");
    bool alpb = false;
    int kernel = 2;
    if (mdl == "gbsa") kernel = 1;
    else if (mdl == "alpb") alpb = true;
    else if (mdl == "cosmo" || mdl == "tmcosmo" || mdl == "cpcmx") { unsupported(env, "The COSMO family of solvation models", source); return; }
    else { env->error("Unknown solvation model", source); return; }
    if (nang != 230) { env->error("Only the 230 point angular grid is available for the surface integration", source); return; }
    try {
        calc->batch->set_solvent(solvent, alpb, kernel, st, temp, 0.0, 0.0);
        calc->solv_on = true; calc->solv_alpb = alpb; calc->solv_name = solvent; calc->solv_kernel = kernel; calc->solv_state = st;
        calc->solv_temp = temp;
    } catch (const std::exception& e) {
        env->error(e.what(), "solv_model_loadInternalParam");
        env->error(std::string("Could not add solvation model for '") + solvent + "'", source);
    }
}
// src/api/calculator.f90:449-476: no error on calculators without a solvation model
void xtb_releaseSolvent(xtb_TEnvironment env, xtb_TCalculator calc) {
    if (!check_env(env)) return;
    if (!calc) { env->error("Singlepoint calculator is not allocated", "xtb_api_setRelease"); return; }
    calc->solv_on = false;
    if (calc->batch) calc->batch->release_solvent();
}
// src/api/calculator.f90:470-540
void xtb_setExternalCharges(xtb_TEnvironment env, xtb_TCalculator calc, int* n, int* numbers, double* charges, double* positions) {
    const char* source = "xtb_api_setExternalCharges";
    if (!check_env(env)) return;
    if (!n || !numbers || !charges || !positions) { env->error("External charges are not fully specified", source); return; }
    if (*n <= 0) { env->error("Negative number of point charges provided", source); return; }
    if (!calc) { env->error("Singlepoint calculator is not allocated", source); return; }
    if (!calc->loaded) { env->error("No calculator loaded to add external potential", source); return; }
    for (int i = 0; i < *n; ++i)
        if (numbers[i] < 1 || numbers[i] > 86) { env->error("Element of an external charge is not parametrised in GFN2-xTB", source); return; }
    calc->pc_numbers.assign(numbers, numbers + *n);
    calc->pc_q.assign(charges, charges + *n);
    calc->pc_xyz.assign(positions, positions + 3 * (size_t)*n);
}
// src/api/calculator.f90:545-580: no errors on calculators without point charges
void xtb_releaseExternalCharges(xtb_TEnvironment env, xtb_TCalculator calc) {
    if (!check_env(env)) return;
    if (!calc) { env->error("Singlepoint calculator is not allocated", "xtb_api_releaseExternalCharges"); return; }
    calc->pc_numbers.clear(); calc->pc_q.clear(); calc->pc_xyz.clear();
}
void xtb_setAccuracy(xtb_TEnvironment env, xtb_TCalculator calc, double accuracy) {
    const char* source = "xtb_api_setAccuracy";
    if (!check_env(env)) return;
    if (!calc) { env->error("Calculator object is not allocated", source); return; }
    if (!calc->loaded) { env->error("Calculator is not allocated", source); return; }
    if (accuracy < 1.e-4) { env->warning("We cannot provide this level of accuracy, resetted accuracy to 0.0001", source); calc->accuracy = 1.e-4; }
    else if (accuracy > 1.e+3) { env->warning("We cannot provide this level of accuracy, resetted accuracy to 1000", source); calc->accuracy = 1.e+3; }
    else calc->accuracy = accuracy;
}
void xtb_setMaxIter(xtb_TEnvironment env, xtb_TCalculator calc, int iterations) {
    const char* source = "xtb_api_setMaxIter";
    if (!check_env(env)) return;
    if (!calc) { env->error("Calculator object is not allocated", source); return; }
    if (!calc->loaded) { env->error("Calculator is not allocated", source); return; }
    calc->maxiter = iterations < 1 ? 1 : iterations;
}
void xtb_setElectronicTemp(xtb_TEnvironment env, xtb_TCalculator calc, double temperature) {
    const char* source = "xtb_api_setElectronicTemp";
    if (!check_env(env)) return;
    if (!calc) { env->error("Calculator object is not allocated", source); return; }
    if (!calc->loaded) { env->error("Calculator is not allocated", source); return; }
    calc->etemp = temperature < 1.e-6 ? 1.e-6 : temperature;
}
void xtb_cpcmx_calc(xtb_TEnvironment env, xtb_TMolecule, xtb_TCalculator, xtb_TResults) { unsupported(env, "CPCM-X", "xtb_api_cpcmx_calc"); }
// Numerical Hessian by central differences of analytical gradients (src/api/interface.f90:312-456 ->
// TCalculator%hessian, src/type/calculator.f90:116-229).  The reference loops over 6 N displaced single points (OpenMP over
// the points); here they are one conformer ensemble: every displaced geometry is a member of a batch, all of them start
// from the wavefunction cached on the results handle (or from the EEQ guess when there is none, as newWavefunction does),
// and the batch runs through the same kernels as xtbb200_singlepoint.  hessian: (3N x 3N) symmetrised as in
// interface.f90:438-444; dipole_gradient (3, 3N) optional.  atom_index_list / step_size: NULL = all atoms / 0.005 bohr
// (the reference's defaults); a non-NULL step_size is read as a double, which is what the reference's Fortran side does
// with this argument (interface.f90:330).  Deliberate deviation, documented in include/xtb.h and INTEGRATION.md: the
// argument ORDER follows the C header (include/xtb.h:244-252: hessian, atom_index_list, step_size), whereas the
// reference's Fortran binding declares (c_hess, c_step, c_list) (interface.f90:310-311), i.e. its binary reads the 6th
// pointer as the step and the 7th as the list.  Callers that pass NULL for both (the only portable use) are unaffected.  Polarizabilities are not available for GFN2-xTB: a non-NULL
// polarizability_gradient is an error.
void xtb_hessian(xtb_TEnvironment env, xtb_TMolecule mol, xtb_TCalculator calc, xtb_TResults res, double* hess, int* atom_list,
                 int* step_size, double* dipgrad, double* polgrad) {
    const char* source = "xtb_api_hessian";
    if (!check_env(env)) return;
    if (!mol) { env->error("Molecular structure data is not allocated", source); return; }
    if (!calc) { env->error("Singlepoint calculator is not allocated", source); return; }
    if (!calc->loaded || !calc->batch) { env->error("No calculator loaded for single point", source); return; }
    if (!res) { env->error("Calculation results are not allocated", source); return; }
    if (!hess) { env->error("Hessian array is not allocated", source); return; }
    if (atom_list) { env->error("Atom index lists are not supported, pass NULL to displace all atoms", source); return; }
    if (polgrad) { env->error("Polarizability gradients are not available for GFN2-xTB", source); return; }
    if (calc->at != mol->at) { env->error("Calculator was loaded for a different molecule", source); return; }
    try {
        const int n = mol->n, n3 = 3 * n;
        const double step = step_size ? *reinterpret_cast<const double*>(step_size) : 0.005;
        if (!(step > 0.0)) { env->error("Step size for numerical differentiation must be positive", source); return; }
        const int nao = calc->batch->nao(0), nsh = calc->batch->nshell(0);
        const bool cached = res->has_wfn && res->n == n && res->nao == nao && res->nshell == nsh;   // interface.f90:369-388
        // points per pass: bounded by a memory estimate of ~40 nao^2 doubles per member
        const double per = 40.0 * 8.0 * (double)nao * nao + 1.0e6;
        const int chunk = (int)std::max(1.0, std::min((double)(2 * n3), std::floor(6.0e9 / per)));
        if (!calc->hess_batch || calc->hess_copies != chunk) {
            std::vector<int> nat(chunk, n), num;
            std::vector<double> chg(chunk, calc->chrg);
            std::vector<int> uhf(chunk, calc->uhf);
            num.reserve((size_t)chunk * n);
            for (int c = 0; c < chunk; ++c) num.insert(num.end(), mol->at.begin(), mol->at.end());
            calc->hess_batch.reset(new Batch(chunk, nat.data(), num.data(), chg.data(), uhf.data(), 0));
            calc->hess_copies = chunk;
        }
        Batch& b = *calc->hess_batch;
        b.set_options(calc->accuracy, calc->etemp, calc->maxiter);
        b.set_guess(0);
        // central differences with a step of 0.005 bohr multiply the noise of gradients and dipoles by 100: the eigenvectors of
        // the displaced geometries are converged further than a single point needs (quadratic stop)
        b.set_jacobi_stop(1.0e-9);
        if (calc->solv_on) b.set_solvent(calc->solv_name.c_str(), calc->solv_alpb, calc->solv_kernel, calc->solv_state, calc->solv_temp, 0.0, 0.0);
        else if (b.has_solvent()) b.release_solvent();
        const int npc = (int)calc->pc_q.size();
        if (npc > 0) {
            std::vector<int> cnt(chunk, npc), pnum;
            std::vector<double> pq, px;
            for (int c = 0; c < chunk; ++c) {
                pnum.insert(pnum.end(), calc->pc_numbers.begin(), calc->pc_numbers.end());
                pq.insert(pq.end(), calc->pc_q.begin(), calc->pc_q.end());
                px.insert(px.end(), calc->pc_xyz.begin(), calc->pc_xyz.end());
            }
            b.set_external_charges(cnt.data(), pnum.data(), pq.data(), px.data());
        } else if (b.external_charge_count() > 0) b.release_external_charges();
        std::vector<double> H((size_t)n3 * n3, 0.0), dg((size_t)3 * n3, 0.0);
        std::vector<double> pos((size_t)chunk * n3), grad((size_t)chunk * n3), qsh, dipm, qp;
        if (cached) {
            for (int c = 0; c < chunk; ++c) {
                qsh.insert(qsh.end(), res->qsh.begin(), res->qsh.end());
                dipm.insert(dipm.end(), res->dipm.begin(), res->dipm.end());
                qp.insert(qp.end(), res->qp.begin(), res->qp.end());
            }
        }
        const double step2 = 0.5 / step;
        int nbad = 0;
        // point index p = 2 * ii + (0: +step, 1: -step), ii = 3 * iat + ic
        for (int p0 = 0; p0 < 2 * n3; p0 += chunk) {
            const int np = std::min(chunk, 2 * n3 - p0);
            for (int c = 0; c < chunk; ++c) {
                const int p = std::min(p0 + c, 2 * n3 - 1);          // members beyond the last point repeat it
                std::copy(mol->xyz.begin(), mol->xyz.end(), pos.begin() + (size_t)c * n3);
                pos[(size_t)c * n3 + p / 2] += (p & 1) ? -step : step;
            }
            b.set_positions(pos.data());
            if (cached) b.set_wavefunction(qsh.data(), dipm.data(), qp.data());
            else b.invalidate_wavefunction();
            b.singlepoint(true);
            b.fetch("gradient", grad.data());
            for (int c = 0; c < np; ++c) {
                const int p = p0 + c, ii = p / 2;
                const double sgn = (p & 1) ? -step2 : step2;
                MolResult r = b.result(c);
                if (r.fail || !r.converged) ++nbad;
                for (int jj = 0; jj < n3; ++jj) H[jj + (size_t)n3 * ii] += sgn * grad[(size_t)c * n3 + jj];
                for (int k = 0; k < 3; ++k) dg[k + (size_t)3 * ii] += sgn * r.dipole[k];
            }
        }
        for (int i = 0; i < n3; ++i)                                 // interface.f90:438-444
            for (int j = i + 1; j < n3; ++j) {
                const double v = 0.5 * (H[i + (size_t)n3 * j] + H[j + (size_t)n3 * i]);
                H[i + (size_t)n3 * j] = v; H[j + (size_t)n3 * i] = v;
            }
        std::copy(H.begin(), H.end(), hess);
        if (dipgrad) std::copy(dg.begin(), dg.end(), dipgrad);
        if (nbad) env->error("Self consistent charge iterator did not converge in " + std::to_string(nbad) + " displaced geometries", "scf");
    } catch (const std::exception& e) {
        env->error(std::string("Hessian calculation failed: ") + e.what(), source);
    }
}

// Fortran Ew.d edit descriptor: 0.dddddE+ee, right-adjusted in w columns
static std::string fortran_e(double v, int w, int d) {
    char buf[64];
    if (v == 0.0 || !(v == v) || std::fabs(v) > 1e300) {
        snprintf(buf, sizeof buf, "%*.*E", w, d, v);
        if (v == 0.0) snprintf(buf, sizeof buf, "%*s", w, (std::string("0.") + std::string(d, '0') + "E+00").c_str());
        return buf;
    }
    int e = (int)std::floor(std::log10(std::fabs(v))) + 1;
    double m = v / std::pow(10.0, e);
    char mant[48];
    snprintf(mant, sizeof mant, "%.*f", d, std::fabs(m));
    if (mant[0] == '1') { ++e; m /= 10.0; snprintf(mant, sizeof mant, "%.*f", d, std::fabs(m)); }   // rounding carried into the leading digit
    char body[64];
    snprintf(body, sizeof body, "%s%s%c%02d", v < 0 ? "-" : "", mant, e < 0 ? '-' : '+', std::abs(e));
    std::string t(body);
    t.insert(t.find_last_of("+-"), "E");
    snprintf(buf, sizeof buf, "%*s", w, t.c_str());
    return buf;
}

void xtb_singlepoint(xtb_TEnvironment env, xtb_TMolecule mol, xtb_TCalculator calc, xtb_TResults res) {
    const char* source = "xtb_api_singlepoint";
    if (!check_env(env)) return;
    if (!mol) { env->error("Molecular structure data is not allocated", source); return; }
    if (!calc) { env->error("Calculator object is not allocated", source); return; }
    if (!calc->loaded || !calc->batch) { env->error("Calculator is not allocated", source); return; }
    if (!res) { env->error("Results object is not allocated", source); return; }
    if (calc->at != mol->at) { env->error("Calculator was loaded for a different molecule", source); return; }
    try {
        Batch& b = *calc->batch;
        const int n = mol->n, nao = b.nao(0), nsh = b.nshell(0);
        // restart cache check, src/api/interface.f90:82-101
        bool cached = res->has_wfn && res->n == n && res->nao == nao && res->nshell == nsh;
        if (!cached) res->has_wfn = false;
        b.set_options(calc->accuracy, calc->etemp, calc->maxiter);
        b.set_guess(0);
        b.set_positions(mol->xyz.data());
        const int npc = (int)calc->pc_q.size();
        if (npc > 0) b.set_external_charges(&npc, calc->pc_numbers.data(), calc->pc_q.data(), calc->pc_xyz.data());
        else if (b.external_charge_count() > 0) b.release_external_charges();
        if (cached) b.set_wavefunction(res->qsh.data(), res->dipm.data(), res->qp.data());
        else b.invalidate_wavefunction();
        b.set_iteration_log(env->verbosity > 0);
        int nfail = b.singlepoint(true);
        MolResult r = b.result(0);
        res->n = n; res->nao = nao; res->nshell = nsh;
        res->energy = r.energy; res->egap = r.hl_gap; res->niter = r.niter;
        res->gradient.resize(3 * n); b.fetch("gradient", res->gradient.data());
        res->q.resize(n); b.fetch("q", res->q.data());
        res->qsh.resize(nsh); b.fetch("qsh", res->qsh.data());
        res->dipm.resize(3 * n); b.fetch("dipm", res->dipm.data());
        res->qp.resize(6 * n); b.fetch("qp", res->qp.data());
        res->wbo.resize((size_t)n * n); b.fetch("wbo", res->wbo.data());
        res->emo.resize(nao); b.fetch("emo", res->emo.data());
        res->focc.resize(nao); b.fetch("focc", res->focc.data());
        res->C.resize((size_t)nao * nao); b.fetch("C", res->C.data());
        for (int k = 0; k < 3; ++k) res->dipole[k] = r.dipole[k];
        res->has_pcgradient = npc > 0;   // src/api/interface.f90:111-113,148-150
        res->pcgradient.clear();
        if (npc > 0) { res->pcgradient.resize(3 * (size_t)npc); b.fetch("pcgradient", res->pcgradient.data()); }
        res->has_energy = res->has_gradient = res->has_wfn = res->has_dipole = true;
        if (env->verbosity > 0) {
            FILE* f = env->out();
            // the SCC iteration table of the reference (header src/scf_module.F90:612-616, rows src/scc_core.f90:621-622,
            // format (i4,F15.7,E14.6,E11.3,f8.2,2x,f8.1,l3); the engine always diagonalises fully)
            std::vector<double> lg(5 * (size_t)std::max(r.niter, 1));
            const int rows = b.iteration_log(lg.data(), std::max(r.niter, 1));
            fprintf(f, "\n iter      E             dE          RMSdq      gap      omega  full diag\n");
            for (int k = 0; k < rows; ++k) {
                const double* x = &lg[5 * (size_t)k];
                fprintf(f, "%4d%15.7f%s%s%8.2f  %8.1f  T\n", k + 1, x[0], fortran_e(x[1], 14, 6).c_str(), fortran_e(x[2], 11, 3).c_str(),
                        x[3], x[4]);
            }
            fprintf(f, "\n");
            fprintf(f, "   *** GFN2-xTB (B200): %s after %d iterations ***\n",
                    r.converged ? "convergence criteria satisfied" : "convergence criteria cannot be satisfied", r.niter);
            if (env->verbosity > 1) {
                fprintf(f, "          total energy      %18.12f Eh\n", r.energy);
                fprintf(f, "          gradient norm     %18.12f Eh/a0\n", r.gnorm);
                fprintf(f, "          HOMO-LUMO gap     %18.12f eV\n", r.hl_gap);
            }
            fflush(f);
        }
        if (r.fail) env->error("Electronic structure method terminated (device error code " + std::to_string(r.fail) + ")", source);
        else if (nfail) env->error("Self consistent charge iterator did not converge", "scf");   // src/scf_module.F90:806-808
    } catch (const std::exception& e) {
        env->error(std::string("Single point calculation failed: ") + e.what(), source);
    }
}

// ---------------------------------------------------------------- results
xtb_TResults xtb_newResults(void) { return new _xtb_TResults(); }
void xtb_delResults(xtb_TResults* res) {
    if (!res || !*res) return;
    delete *res;
    *res = nullptr;
}
xtb_TResults xtb_copyResults(xtb_TResults res) {
    if (!res) return nullptr;
    return new _xtb_TResults(*res);
}
#define XTB_GETTER_PROLOGUE(src)                                                     \
    const char* source = src;                                                        \
    if (!check_env(env)) return;                                                     \
    if (!res) { env->error("Results object is not allocated", source); return; }
void xtb_getEnergy(xtb_TEnvironment env, xtb_TResults res, double* energy) {
    XTB_GETTER_PROLOGUE("xtb_api_getEnergy")
    if (!res->has_energy) { env->error("Energy is not available in results", source); return; }
    if (energy) *energy = res->energy;
}
void xtb_getSolvationEnergy(xtb_TEnvironment env, xtb_TResults res, double*) {
    XTB_GETTER_PROLOGUE("xtb_api_getSolvationEnergy")
    env->error("Solvation energy is not available in results", source);
}
void xtb_getGradient(xtb_TEnvironment env, xtb_TResults res, double* gradient) {
    XTB_GETTER_PROLOGUE("xtb_api_getGradient")
    if (!res->has_gradient) { env->error("Gradient is not available in results", source); return; }
    if (gradient) std::memcpy(gradient, res->gradient.data(), res->gradient.size() * sizeof(double));
}
void xtb_getPCGradient(xtb_TEnvironment env, xtb_TResults res, double* gradient) {
    XTB_GETTER_PROLOGUE("xtb_api_getPCGradient")
    if (!res->has_pcgradient) { env->error("Gradient on external charges is not available in results", source); return; }
    if (gradient) std::copy(res->pcgradient.begin(), res->pcgradient.end(), gradient);
}
void xtb_getVirial(xtb_TEnvironment env, xtb_TResults res, double* virial) {
    XTB_GETTER_PROLOGUE("xtb_api_getVirial")
    if (!res->has_gradient) { env->error("Virial is not available in results", source); return; }
    // GFN1/2 never fill sigma in the reference (src/xtb/calculator.f90:247; scf has no sigma argument)
    if (virial) for (int k = 0; k < 9; ++k) virial[k] = 0.0;
}
void xtb_getDipole(xtb_TEnvironment env, xtb_TResults res, double* dipole) {
    XTB_GETTER_PROLOGUE("xtb_api_getDipole")
    if (!res->has_dipole) { env->error("Dipole is not available in results", source); return; }
    if (dipole) for (int k = 0; k < 3; ++k) dipole[k] = res->dipole[k];
}
void xtb_getCharges(xtb_TEnvironment env, xtb_TResults res, double* charges) {
    XTB_GETTER_PROLOGUE("xtb_api_getCharges")
    if (!res->has_wfn) { env->error("Partial charges are not available in results", source); return; }
    if (charges) std::memcpy(charges, res->q.data(), res->q.size() * sizeof(double));
}
void xtb_getBondOrders(xtb_TEnvironment env, xtb_TResults res, double* wbo) {
    XTB_GETTER_PROLOGUE("xtb_api_getBondOrders")
    if (!res->has_wfn) { env->error("Bond orders are not available in results", source); return; }
    if (wbo) std::memcpy(wbo, res->wbo.data(), res->wbo.size() * sizeof(double));
}
void xtb_getNao(xtb_TEnvironment env, xtb_TResults res, int* nao) {
    XTB_GETTER_PROLOGUE("xtb_api_getNao")
    if (!res->has_wfn) { env->error("Wavefunction is not available in results", source); return; }
    if (nao) *nao = res->nao;
}
void xtb_getOrbitalEigenvalues(xtb_TEnvironment env, xtb_TResults res, double* emo) {
    XTB_GETTER_PROLOGUE("xtb_api_getOrbitalEigenvalues")
    if (!res->has_wfn) { env->error("Orbital eigenvalues are not available in results", source); return; }
    if (emo) for (int i = 0; i < res->nao; ++i) emo[i] = res->emo[i] * (1.0 / 27.21138505);   // src/api/results.f90:429
}
void xtb_getOrbitalOccupations(xtb_TEnvironment env, xtb_TResults res, double* focc) {
    XTB_GETTER_PROLOGUE("xtb_api_getOrbitalOccupations")
    if (!res->has_wfn) { env->error("Occupation numbers are not available in results", source); return; }
    if (focc) std::memcpy(focc, res->focc.data(), res->focc.size() * sizeof(double));
}
void xtb_getOrbitalCoefficients(xtb_TEnvironment env, xtb_TResults res, double* c) {
    XTB_GETTER_PROLOGUE("xtb_api_getOrbitalCoefficients")
    if (!res->has_wfn) { env->error("Orbital coefficients are not available in results", source); return; }
    if (c) std::memcpy(c, res->C.data(), res->C.size() * sizeof(double));
}

// ================================================================= batched extension
struct _xtbb200_TBatch {
    std::unique_ptr<Batch> b;
    std::vector<double> scratch;
};

const char* xtbb200_lastError(void) { return g_last_error.c_str(); }
int xtbb200_deviceCount(void) { return xtbb200::device_count(); }

#define XTBB_TRY try {
#define XTBB_CATCH(rc)                   \
    }                                    \
    catch (const std::exception& e) {    \
        g_last_error = e.what();         \
        return rc;                       \
    }

xtbb200_TBatch xtbb200_newBatch(int nmol, const int* natoms, const int* numbers, const double* charges, const int* uhf, int device) {
    XTBB_TRY
    if (nmol <= 0 || !natoms || !numbers) throw std::runtime_error("invalid batch description");
    auto* h = new _xtbb200_TBatch();
    try { h->b.reset(new Batch(nmol, natoms, numbers, charges, uhf, device)); }
    catch (...) { delete h; throw; }
    g_last_error.clear();
    return h;
    XTBB_CATCH(nullptr)
}
void xtbb200_delBatch(xtbb200_TBatch* batch) {
    if (!batch || !*batch) return;
    delete *batch;
    *batch = nullptr;
}
int xtbb200_setOptions(xtbb200_TBatch batch, double accuracy, double etemp, int maxiter, int guess) {
    XTBB_TRY
    if (!batch) throw std::runtime_error("batch is not allocated");
    accuracy = std::min(1.e+3, std::max(1.e-4, accuracy));
    batch->b->set_options(accuracy, std::max(etemp, 1.e-6), std::max(maxiter, 1));
    batch->b->set_guess(guess);
    return 0;
    XTBB_CATCH(-1)
}
int xtbb200_setWavefunction(xtbb200_TBatch batch, const double* qsh, const double* dipm, const double* qp) {
    XTBB_TRY
    if (!batch) throw std::runtime_error("batch is not allocated");
    if (!qsh || !dipm || !qp) { batch->b->invalidate_wavefunction(); return 0; }
    batch->b->set_wavefunction(qsh, dipm, qp);
    return 0;
    XTBB_CATCH(-1)
}
int xtbb200_setExternalCharges(xtbb200_TBatch batch, const int* npc, const int* numbers, const double* charges, const double* positions) {
    XTBB_TRY
    if (!batch) throw std::runtime_error("batch is not allocated");
    batch->b->set_external_charges(npc, numbers, charges, positions);
    return 0;
    XTBB_CATCH(-1)
}
int xtbb200_setSolvent(xtbb200_TBatch batch, const char* solvent, int alpb, int kernel, int state, double temperature,
                       double ion_strength, double ion_rad) {
    XTBB_TRY
    if (!batch) throw std::runtime_error("batch is not allocated");
    if (!solvent) { batch->b->release_solvent(); return 0; }
    batch->b->set_solvent(solvent, alpb != 0, kernel, state, temperature, ion_strength, ion_rad);
    return 0;
    XTBB_CATCH(-1)
}
int xtbb200_solventParameters(const char* solvent, int alpb, int kernel, int state, double temperature, double* out8) {
    XTBB_TRY
    if (!solvent || !out8) throw std::runtime_error("solvent name and output buffer are required");
    if (kernel != 1 && kernel != 2) throw std::runtime_error("solvation: kernel must be 1 (Still) or 2 (P16)");
    const xtbb200::SolvModel sm = xtbb200::make_solv_model(solvent, alpb != 0, kernel, state, temperature, 0.0, 0.0);
    out8[0] = sm.eps; out8[1] = sm.keps; out8[2] = sm.alpbet; out8[3] = sm.gshift; out8[4] = sm.c1; out8[5] = sm.born_offset;
    out8[6] = sm.probe; out8[7] = sm.any_hb ? 1.0 : 0.0;
    return 0;
    XTBB_CATCH(-1)
}
long long xtbb200_shakeList(int natoms, const int* numbers, const double* positions, const double* wbo, int mode, int hmass,
                            int* pairs, long long capacity, double* masses) {
    XTBB_TRY
    if (natoms < 1 || !numbers || !positions) throw std::runtime_error("atoms and positions are required");
    if (mode < 0 || mode > 2) throw std::runtime_error("md: shake mode must be 0, 1 or 2");
    if (mode == 2 && !wbo) throw std::runtime_error("shake mode 2 needs the bond orders");
    for (int i = 0; i < natoms; ++i)
        if (numbers[i] < 1 || numbers[i] > 86) throw std::runtime_error("element outside 1..86");
    const auto list = xtbb200::shake_constraint_list(natoms, numbers, positions, wbo, mode);
    if (pairs)
        for (size_t c = 0; c < list.size() && (long long)c < capacity; ++c) { pairs[2 * c] = list[c].first; pairs[2 * c + 1] = list[c].second; }
    if (masses)
        for (int i = 0; i < natoms; ++i) masses[i] = xtbb200::md_mass(numbers[i], hmass);
    return (long long)list.size();
    XTBB_CATCH(-1)
}
int xtbb200_mdInit(xtbb200_TBatch batch, const double* positions, double tstep_fs, double temperature, int thermostat,
                   int shake_mode, int hmass, const double* velocities, unsigned long long seed, int restart) {
    XTBB_TRY
    if (!batch) throw std::runtime_error("batch is not allocated");
    if (!positions) throw std::runtime_error("positions are required");
    batch->b->md_init(positions, tstep_fs, temperature, thermostat != 0, shake_mode, hmass, velocities, seed, restart != 0);
    return 0;
    XTBB_CATCH(-1)
}
int xtbb200_mdRun(xtbb200_TBatch batch, int nsteps, double* epot, double* ekin, double* temperature) {
    XTBB_TRY
    if (!batch) throw std::runtime_error("batch is not allocated");
    if (nsteps < 0) throw std::runtime_error("negative number of steps");
    return batch->b->md_run(nsteps, epot, ekin, temperature);
    XTBB_CATCH(-1)
}
long long xtbb200_mdConstraints(xtbb200_TBatch batch) {
    XTBB_TRY
    if (!batch) throw std::runtime_error("batch is not allocated");
    return batch->b->md_constraints();
    XTBB_CATCH(-1)
}
int xtbb200_setSolver(xtbb200_TBatch batch, int mode) {
    XTBB_TRY
    if (!batch) throw std::runtime_error("batch is not allocated");
    batch->b->set_solver(mode);
    return 0;
    XTBB_CATCH(-1)
}
int xtbb200_symmetricEigen(int n, const double* a, double* w, double* z, float* ms3, int device) {
    XTBB_TRY
    return xtbb200::symmetric_eigen(n, a, w, z, ms3, device);
    XTBB_CATCH(-1)
}
int xtbb200_shardUniqueId(char* id128, size_t len) {
    XTBB_TRY
    if (!id128 || len < 128) throw std::runtime_error("need a 128-byte buffer for the NCCL unique id");
    xtbb200::shard_unique_id(id128);
    return 0;
    XTBB_CATCH(-1)
}
int xtbb200_setShard(xtbb200_TBatch batch, int rank, int nranks, const char* id128, size_t len) {
    XTBB_TRY
    if (!batch) throw std::runtime_error("batch is not allocated");
    if (nranks > 1 && (!id128 || len < 128)) throw std::runtime_error("need the 128-byte NCCL unique id of rank 0");
    batch->b->set_shard(rank, nranks, id128);
    return 0;
    XTBB_CATCH(-1)
}
int xtbb200_shardOwner(int kind, long long index, int nranks) {
    if (nranks < 1 || index < 0) return -1;
    if (kind == 0) return xtbb200::shard_atom_owner_host((int)index, nranks);
    if (kind == 1) return xtbb200::shard_pair_cta_owner_host(index, nranks);
    return -1;
}
int xtbb200_getCommTime(xtbb200_TBatch batch, float* ms) {
    XTBB_TRY
    if (!batch || !ms) throw std::runtime_error("batch is not allocated");
    *ms = batch->b->comm_ms();
    return 0;
    XTBB_CATCH(-1)
}
int xtbb200_singlepoint(xtbb200_TBatch batch, const double* positions, int restart, double* energies, double* gradients,
                        double* charges, int* niter, int* status) {
    XTBB_TRY
    if (!batch) throw std::runtime_error("batch is not allocated");
    if (!positions) throw std::runtime_error("positions are required");
    Batch& b = *batch->b;
    b.mark_call(0);
    b.set_positions(positions);
    int nfail = b.singlepoint(restart != 0);
    for (int i = 0; i < b.nmol(); ++i) {
        MolResult r = b.result(i);
        if (energies) energies[i] = r.energy;
        if (niter) niter[i] = r.niter;
        if (status) status[i] = r.fail ? r.fail : (r.converged ? 0 : 1);
    }
    if (gradients) b.fetch("gradient", gradients);
    if (charges) b.fetch("q", charges);
    b.mark_call(1);
    return nfail ? 1 : 0;
    XTBB_CATCH(-1)
}
int xtbb200_singlepointDevice(xtbb200_TBatch batch, const double* positions_dev, int restart, double*) {
    XTBB_TRY
    if (!batch) throw std::runtime_error("batch is not allocated");
    Batch& b = *batch->b;
    b.set_positions_device(positions_dev);
    int nfail = b.singlepoint(restart != 0);
    return nfail ? 1 : 0;
    XTBB_CATCH(-1)
}
int xtbb200_getNMol(xtbb200_TBatch batch) { return batch ? batch->b->nmol() : -1; }
int xtbb200_getSizes(xtbb200_TBatch batch, int imol, int* natoms, int* nshell, int* nao) {
    XTBB_TRY
    if (!batch || imol < 0 || imol >= batch->b->nmol()) throw std::runtime_error("invalid molecule index");
    if (natoms) *natoms = batch->b->natoms(imol);
    if (nshell) *nshell = batch->b->nshell(imol);
    if (nao) *nao = batch->b->nao(imol);
    return 0;
    XTBB_CATCH(-1)
}
int xtbb200_getElectrons(xtbb200_TBatch batch, int imol, int* nel, int* nopen) {
    XTBB_TRY
    if (!batch || imol < 0 || imol >= batch->b->nmol()) throw std::runtime_error("invalid molecule index");
    if (nel) *nel = batch->b->nel(imol);
    if (nopen) *nopen = batch->b->nopen(imol);
    return 0;
    XTBB_CATCH(-1)
}
size_t xtbb200_getDeviceBytes(xtbb200_TBatch batch) { return batch ? batch->b->device_bytes() : 0; }
int xtbb200_getResult(xtbb200_TBatch batch, int imol, double* out16) {
    XTBB_TRY
    if (!batch || imol < 0 || imol >= batch->b->nmol()) throw std::runtime_error("invalid molecule index");
    MolResult r = batch->b->result(imol);
    const double v[16] = {r.energy, r.e_elec, r.e_rep, r.e_es, r.e_aes, r.e_axc, r.e_disp, r.hl_gap, r.gnorm, r.e_atm,
                          r.dipole[0], r.dipole[1], r.dipole[2], (double)r.niter, (double)r.converged, (double)r.fail};
    std::memcpy(out16, v, sizeof(v));
    return 0;
    XTBB_CATCH(-1)
}
int xtbb200_getResultsInfo(void* xtb_results, double* hl_gap_eV, int* scc_iterations) {
    XTBB_TRY
    const _xtb_TResults* res = (const _xtb_TResults*)xtb_results;
    if (!res || !res->has_energy) throw std::runtime_error("results handle holds no single point");
    if (hl_gap_eV) *hl_gap_eV = res->egap;
    if (scc_iterations) *scc_iterations = res->niter;
    return 0;
    XTBB_CATCH(-1)
}
int xtbb200_getCounts(xtbb200_TBatch batch, int imol, int* ndp, int* nqp, int* sweeps) {
    XTBB_TRY
    if (!batch || imol < 0 || imol >= batch->b->nmol()) throw std::runtime_error("invalid molecule index");
    MolResult r = batch->b->result(imol);
    if (ndp) *ndp = r.ndp;
    if (nqp) *nqp = r.nqp;
    if (sweeps) *sweeps = r.sweeps;
    return 0;
    XTBB_CATCH(-1)
}
int xtbb200_setDecisionCounting(xtbb200_TBatch batch, int on) {
    XTBB_TRY
    if (!batch) throw std::runtime_error("batch is not allocated");
    batch->b->set_decision_counting(on != 0);
    return 0;
    XTBB_CATCH(-1)
}
int xtbb200_getDecisionCounts(xtbb200_TBatch batch, int imol, long long* counts12) {
    XTBB_TRY
    if (!batch || !counts12) throw std::runtime_error("batch is not allocated");
    batch->b->decision_counts(imol, counts12);
    return 0;
    XTBB_CATCH(-1)
}
int xtbb200_getArrayLen(xtbb200_TBatch batch, const char* name, size_t* len) {
    XTBB_TRY
    if (!batch) throw std::runtime_error("batch is not allocated");
    *len = batch->b->fetch_len(name);
    return 0;
    XTBB_CATCH(-1)
}
int xtbb200_getArray(xtbb200_TBatch batch, const char* name, double* out, size_t len) {
    XTBB_TRY
    if (!batch) throw std::runtime_error("batch is not allocated");
    if (len < batch->b->fetch_len(name)) throw std::runtime_error("output buffer too small");
    batch->b->fetch(name, out);
    return 0;
    XTBB_CATCH(-1)
}
int xtbb200_getOffsets(xtbb200_TBatch batch, int imol, long long* off5) {
    XTBB_TRY
    if (!batch || imol < 0 || imol >= batch->b->nmol()) throw std::runtime_error("invalid molecule index");
    off5[0] = batch->b->atom_offset(imol); off5[1] = batch->b->shell_offset(imol); off5[2] = batch->b->ao_offset(imol);
    off5[3] = batch->b->mat_offset(imol); off5[4] = batch->b->nn_offset(imol);
    return 0;
    XTBB_CATCH(-1)
}
int xtbb200_getMask(xtbb200_TBatch batch, unsigned char* out, size_t len) {
    XTBB_TRY
    if (!batch) throw std::runtime_error("batch is not allocated");
    if (len < batch->b->fetch_len("S")) throw std::runtime_error("output buffer too small");
    batch->b->fetch_mask(out);
    return 0;
    XTBB_CATCH(-1)
}
int xtbb200_getTimings(xtbb200_TBatch batch, float* ms3, int* launches) {
    XTBB_TRY
    if (!batch) throw std::runtime_error("batch is not allocated");
    if (ms3) batch->b->timings(ms3);
    if (launches) *launches = batch->b->launches();
    return 0;
    XTBB_CATCH(-1)
}
int xtbb200_getFactorisationTime(xtbb200_TBatch batch, float* ms) {
    XTBB_TRY
    if (!batch) throw std::runtime_error("batch is not allocated");
    *ms = batch->b->factorisation_ms();
    return 0;
    XTBB_CATCH(-1)
}
int xtbb200_getCallTime(xtbb200_TBatch batch, float* ms) {
    XTBB_TRY
    if (!batch) throw std::runtime_error("batch is not allocated");
    *ms = batch->b->call_ms();
    return 0;
    XTBB_CATCH(-1)
}
int xtbb200_setProfile(xtbb200_TBatch batch, unsigned long long class_mask) {
    XTBB_TRY
    if (!batch) throw std::runtime_error("batch is not allocated");
    batch->b->set_profile(class_mask);
    return 0;
    XTBB_CATCH(-1)
}
int xtbb200_kernelClassCount(void) { return Batch::kernel_classes(); }
const char* xtbb200_kernelClassName(int id) { return Batch::kernel_class_name(id); }
int xtbb200_getKernelTimes(xtbb200_TBatch batch, float* ms, int* timed_launches, int* launches, int* scc_passes) {
    XTBB_TRY
    if (!batch) throw std::runtime_error("batch is not allocated");
    batch->b->kernel_times(ms, timed_launches, launches);
    if (scc_passes) *scc_passes = batch->b->scc_passes();
    return 0;
    XTBB_CATCH(-1)
}

}  // extern "C"
