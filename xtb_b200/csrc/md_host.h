// Host-side set-up of the molecular dynamics driver: SHAKE constraint selection and masses.  Pure host code, shared by
// Batch::md_init and the C ABI query xtbb200_shakeList (CPU tests).
//
// Reference being replaced: init_shake (src/shake_module.f90:50-148), masses of md (src/dynamic.f90:237-252).
#pragma once
#include <cmath>
#include <utility>
#include <vector>

#include "system.h"
#include "tables/md_tables.h"

namespace xtbb200 {

constexpr double kAmuToAU = 1.660539040e-27 * (1.0 / 9.10938356e-31);   // src/mctc/convert.f90:51-55
constexpr double kFsToAU = 1.0e-15 / 2.4188843265857e-17;               // src/mctc/convert.f90:58-60

// Constraint pairs (first, second) of one molecule in the order init_shake creates them.
//   mode 1: every hydrogen to its nearest neighbour (an H-H pair once, from its lower index)
//   mode 2: all bonds: pairs i < j closer than 1.2 x the sum of the atomic radii whose Wiberg bond order exceeds 0.5
//           (0.1 when a metal takes part); bonds to metals other than Li / Be are not constrained
// xyz: [nat][3] in bohr, wbo: nat x nat (symmetric), Z: atomic numbers (<= 86)
inline std::vector<std::pair<int, int>> shake_constraint_list(int nat, const int* Z, const double* xyz, const double* wbo, int mode) {
    std::vector<std::pair<int, int>> list;
    auto r2 = [&](int i, int j) { double s = 0; for (int k = 0; k < 3; ++k) { const double t = xyz[3 * i + k] - xyz[3 * j + k]; s += t * t; } return s; };
    if (mode == 1) {
        for (int i = 0; i < nat; ++i) {
            if (Z[i] != 1) continue;
            double minr = 1000.0; int jmin = -1;
            for (int j = 0; j < nat; ++j) if (j != i && r2(i, j) < minr) { minr = r2(i, j); jmin = j; }
            if (jmin < 0) continue;
            if (Z[jmin] == 1) { if (jmin > i) list.push_back({i, jmin}); }
            else list.push_back({i, jmin});
        }
    } else if (mode == 2) {
        for (int i = 0; i < nat; ++i)
            for (int j = i + 1; j < nat; ++j) {
                const int zi = Z[i], zj = Z[j];
                const double rco = (kAAtoAU * param_atomic_rad_aa[zi - 1] + kAAtoAU * param_atomic_rad_aa[zj - 1]) * kAutoAA;
                if (!(0.52917726 * std::sqrt(r2(i, j)) < 1.2 * rco)) continue;
                bool metalbond = md_metal[zi - 1] == 1 || md_metal[zj - 1] == 1;
                const double wthr = metalbond ? 0.1 : 0.5;
                if (!(wbo[i + (size_t)nat * j] > wthr)) continue;
                if (zi == 3 || zi == 4 || zj == 3 || zj == 4) metalbond = false;
                if (!metalbond) list.push_back({i, j});
            }
    }
    return list;
}

// mass of an atom in atomic units; hydrogens take hmass amu when hmass > 0
inline double md_mass(int Z, int hmass) {
    if (Z == 1 && hmass > 0) return (double)hmass * kAmuToAU;
    return md_atomic_mass_amu[Z - 1] * kAmuToAU;
}

}  // namespace xtbb200
