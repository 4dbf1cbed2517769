// Host-side set-up of the implicit solvation model (GBSA / ALPB): solvent name -> internal GFN2-xTB parameter set ->
// per-atom model parameters.  Geometry independent, done in xtb_setSolvent / xtbb200_setSolvent.
//
// Reference being replaced: src/solv/model.f90:185-207 (initSolvModel), 211-246 (normalizeSolventName), 250-474
// (loadInternalParam: Still kernel -> GBSA sets, P16 kernel -> ALPB sets), 498-528 (paramToModel),
// src/solv/state.f90:55-85 (getStateShift), src/solv/gbsa.f90:197-388 (initBorn).  Parameter files on XTBPATH
// (param_gbsa2_<solvent>.txt) are not read: the compiled-in sets are used.
#pragma once
#include <algorithm>
#include <cctype>
#include <cmath>
#include <stdexcept>
#include <string>
#include <vector>

#include "system.h"
#include "tables/solv_tables.h"

namespace xtbb200 {

constexpr double kKcalToAU = 1.0 / 627.50947428;   // src/mctc/convert.f90:31-32

struct SolvModel {
    bool on = false;
    std::string solvent;      // normalised name
    int kernel = 2;           // 1 Still, 2 P16
    bool alpb = false, salt = false;
    int state = 1;            // 1 gsolv, 2 reference, 3 mol1bar
    double temperature = 298.15, ion_strength = 0.0, ion_rad = 0.0;
    const SolvParamSet* ps = nullptr;
    // derived
    double eps = 1, keps = 0, alpbet = 0, gshift = 0, c1 = 1, kappa = 0, born_offset = 0, probe = 0;
    bool any_hb = false;
};

inline std::string solv_normalize_name(const char* name) {
    std::string s(name ? name : "");
    for (auto& c : s) c = (char)std::tolower((unsigned char)c);
    static const char* alias[][2] = {
        {"methylcyanide", "acetonitrile"}, {"methylenechloride", "ch2cl2"}, {"dichloromethane", "ch2cl2"}, {"dcm", "ch2cl2"},
        {"chloroform", "chcl3"}, {"trichloromethane", "chcl3"}, {"tcm", "chcl3"}, {"carbondisulfide", "cs2"},
        {"diethylether", "ether"}, {"dimethylformamide", "dmf"}, {"dimethylsulfoxide", "dmso"}, {"furan", "furane"},
        {"nhexan", "hexane"}, {"n-hexan", "hexane"}, {"nhexane", "hexane"}, {"n-hexane", "hexane"},
        {"tetrahydrofuran", "thf"}, {"h2o", "water"}};
    for (auto& a : alias)
        if (s == a[0]) s = a[1];
    return s;
}

// throws std::runtime_error("solvent: '<name>' is not parametrized") like loadInternalParam (model.f90:310-313)
inline SolvModel make_solv_model(const char* solvent, bool alpb, int kernel, int state, double temperature, double ion_strength,
                                 double ion_rad) {
    SolvModel sm;
    sm.solvent = solv_normalize_name(solvent);
    sm.kernel = kernel; sm.alpb = alpb; sm.state = state; sm.temperature = temperature;
    sm.ion_strength = ion_strength; sm.ion_rad = ion_rad;
    for (int i = 0; i < solv_nsets; ++i)
        if (solv_param_sets[i].kernel == kernel && sm.solvent == solv_param_sets[i].name) sm.ps = &solv_param_sets[i];
    if (!sm.ps) throw std::runtime_error("solvent: '" + sm.solvent + "' is not parametrized");
    const SolvParamSet& p = *sm.ps;
    const double molcm3toau = 8.92388e-2;
    const double density = p.rhos / p.smass * molcm3toau;
    sm.eps = p.epsv;
    sm.gshift = p.gshift * kKcalToAU;
    const double ideal = 24.79, ambient = 298.15;
    if (state == 2) sm.gshift += temperature * kBoltz * (std::log(ideal * temperature / ambient) + std::log(density / 1.0e-3 * 1.0 / p.smass));
    else if (state == 3) sm.gshift += temperature * kBoltz * std::log(ideal * temperature / ambient);
    sm.c1 = p.c1;
    sm.born_offset = p.soset * 0.1 * kAAtoAU;
    sm.probe = p.rprobe * kAAtoAU;
    sm.alpbet = alpb ? 0.571412 / sm.eps : 0.0;
    sm.keps = (1.0 / sm.eps - 1.0) / (1.0 + sm.alpbet);
    sm.salt = ion_strength > 0.0;
    if (sm.salt) sm.kappa = 1.0 / (std::sqrt(sm.eps * temperature * 0.7897e-3 / ion_strength) * kAAtoAU);
    for (int k = 0; k < 94; ++k) sm.any_hb = sm.any_hb || std::fabs(p.tmp[k]) > 1.e-3;
    sm.on = true;
    return sm;
}

// per-atom parameters in the layout of kernels_solv.cuh: vdwr, rho, svdw, vdwsa, wrp, gamsasa, hbmag, vdwr^3
inline void solv_atom_par(const SolvModel& sm, int Z, double* out) {
    if (Z < 1 || Z > 94) throw std::runtime_error("element without solvation parameters");
    const SolvParamSet& p = *sm.ps;
    const double w = 0.3 * kAAtoAU, w3 = w * (w * w), ah3 = -1.0 / (4.0 * w3);
    const double vdwr = kAAtoAU * solv_vdw_rad_d3_aa[Z - 1];
    const double v = vdwr + sm.probe;
    double r = v + w;
    double wr = (0.25 / w + 3.0 * ah3 * (0.2 * r * r - 0.5 * r * v + v * v / 3.0)) * r * r * r;
    r = v - w;
    wr -= (0.25 / w + 3.0 * ah3 * (0.2 * r * r - 0.5 * r * v + v * v / 3.0)) * r * r * r;
    out[0] = vdwr;
    out[1] = vdwr * p.sx[Z - 1];
    out[2] = vdwr - sm.born_offset;
    out[3] = v;
    out[4] = wr;
    out[5] = p.gamscale[Z - 1] * (4.0 * kPi) * 1.0e-5;
    out[6] = sm.any_hb ? -(p.tmp[Z - 1] * p.tmp[Z - 1]) * kKcalToAU : 0.0;
    out[7] = vdwr * vdwr * vdwr;
}

}  // namespace xtbb200
