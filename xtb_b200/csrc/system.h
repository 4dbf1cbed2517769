// Host-side construction of the flat, device-ready description of a batch of molecules:
// GFN2-xTB element parameters, STO-nG shells, D4 reference systems.  Geometry independent,
// done once per calculator (reference: newXTBCalculator, src/xtb/calculator.f90:94-192).
//
// Reference being replaced (data + setup, not arithmetic on the hot path):
//   src/xtb/gfn2.f90:731-967   initData/initCoulomb/initHamiltonian
//   src/xtb/basisset.f90:33-308 newBasisset, src/slater.f90:373-483 slaterToGauss
//   src/disp/dftd4.F90:106-203  newD4Model
#pragma once
#include <cmath>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "tables/d4_tables.h"
#include "tables/eeq_tables.h"
#include "tables/gfn2_tables.h"
#include "tables/slater_tables.h"

namespace xtbb200 {

constexpr double kPi = 3.1415926535897932384626433832795029;
constexpr double kAutoAA = 0.52917726;          // src/mctc/convert.f90:25
constexpr double kAAtoAU = 1.0 / kAutoAA;
constexpr double kAutoEV = 27.21138505;         // src/mctc/convert.f90:28
constexpr double kEVtoAU = 1.0 / kAutoEV;
constexpr double kBoltz = 3.166808578545117e-06;  // src/mctc/mctc_constants.f90:34

// per-molecule descriptor (all offsets index the concatenated batch arrays)
struct MolDesc {
    int nat, nsh, nao, ndim, nbr, nel, nopen, ihomoa, ihomob, pad0;
    int at0, sh0, ao0, ref0;        // offsets: atoms, shells, AOs, D4 references
    long long mat0;                 // nao*nao matrices
    long long nn0;                  // nat*nat matrices
    long long ss0;                  // nsh*nsh matrices
    long long dd0;                  // ndim*ndim matrices
    long long br0;                  // Broyden history (maxhist*nbr)
    long long q0;                   // Broyden vector offset (nbr)
    long long hj0;                  // two-level Jacobi: first panel Gram matrix (128*128 doubles per block pair)
    double charge;
};

struct ShellPar {  // one contracted shell
    int atom;      // atom index within the molecule
    int l;         // 0,1,2
    int nprim;
    int ao0;       // first spherical AO within the molecule
    double alp[6];
    double cont[6];
    double zeta;     // Slater exponent
    double level;    // self energy (eV)
    double kcn;      // CN shift of the level
    double poly;     // shellPoly coefficient
    double hard;     // shell hardness
    double gam3;     // third-order shell parameter
    double refocc;   // reference occupation (zsh)
};

struct AtomPar {
    int Z;
    int etype;      // dense element type index inside this batch (for the C6 table)
    int nref;       // D4 references
    int sh0, nsh;   // shells (within molecule)
    int ao0, nao;   // AOs (within molecule)
    int ref0;       // first reference (within molecule)
    double zval;    // valence charge
    double rep_alpha, rep_zeff;
    double en;        // Pauling EN (param/paulingen.f90)
    double rcov;      // covalentRadD3 (bohr, *4/3)
    double rad;       // atomicRad (bohr)
    double r4r2;      // sqrtZr4r2
    double dipk, quadk, multirad, valcn;
    double eeq_en, eeq_gam, eeq_kappa, eeq_alpha;
    double nc_rcov, nc_en;   // ncoord.f90 single precision tables
    double d4_zeff, d4_gam;
};

struct RefPar {  // one D4 reference system of one atom
    int atom;      // atom within molecule
    int iref;      // index within the element
    int ncount;
    int pad;
    double covcn;  // refcovcn
    double refq;
};

inline double valence_z(int Z) {  // src/type/molecule.f90:560-597
    int nc;
    if (Z <= 2) nc = 0; else if (Z <= 10) nc = 2; else if (Z <= 18) nc = 10; else if (Z <= 29) nc = 18;
    else if (Z <= 36) nc = 28; else if (Z <= 47) nc = 36; else if (Z <= 54) nc = 46; else if (Z <= 71) nc = 54;
    else if (Z <= 79) nc = 68; else if (Z <= 86) nc = 78; else nc = 86;
    double z = (double)Z - (double)nc;
    if (Z > 57 && Z < 72) z = 3.0;
    return z;
}

inline int gfn2_nprim(int Z, int ish) {  // src/xtb/gfn2.f90:893-937
    int l = gfn2_ang_shell[Z - 1][ish], n = gfn2_pqn[Z - 1][ish];
    if (Z <= 2) return l == 0 ? 3 : 4;
    if (l <= 1) return n > 5 ? 6 : 4;
    if (l == 2) return 3;
    return 4;
}

inline void sto_ng(int ng, int n, int l, double zeta, double* alpha, double* coeff) {  // src/slater.f90:373-483
    static const double dfact[4] = {1., 1., 3., 15.};
    int ityp = ((l == 0) ? n : (l == 1) ? 4 + n : 7 + n) - 1;
    for (int p = 0; p < ng; ++p) {
        double a = 0, c = 0;
        if (ng == 3) { a = sto_alpha3[ityp][p]; c = sto_coeff3[ityp][p]; }
        else if (ng == 4) { a = sto_alpha4[ityp][p]; c = sto_coeff4[ityp][p]; }
        else if (ng == 6) {
            if (n == 6) { a = l == 0 ? sto_alpha6s[p] : sto_alpha6p[p]; c = l == 0 ? sto_coeff6s[p] : sto_coeff6p[p]; }
            else { a = sto_alpha6[ityp][p]; c = sto_coeff6[ityp][p]; }
        } else throw std::runtime_error("unsupported STO-nG expansion");
        alpha[p] = a * zeta * zeta;
        coeff[p] = c * std::pow(2.0 / kPi * alpha[p], 0.75) * std::pow(std::sqrt(4 * alpha[p]), (double)l) / std::sqrt(dfact[l]);
    }
}

// D4 reference polarizabilities and pair C6 (src/disp/dftd4.F90:106-203, 348-449)
struct D4Reference {
    double alpha[118][7][23];
    int ncount[118][7];
    bool have[118];
    D4Reference() { std::memset(have, 0, sizeof(have)); }
    static double zeta(double a, double c, double qref, double qmod) {
        if (qmod < 0.0) return std::exp(a);
        return std::exp(a * (1.0 - std::exp(c * (1.0 - qref / qmod))));
    }
    void ensure(int Z) {
        int ia = Z - 1;
        if (have[ia]) return;
        int cncount[19] = {0};
        cncount[0] = 1;
        for (int j = 0; j < d4_refn[ia]; ++j) {
            int is = d4_refsys[ia][j];
            double iz = d4_zeff[is - 1];
            double zf = zeta(3.0, d4_gam[is - 1] * 2.0, iz, d4_refh[ia][j] + iz);
            cncount[(int)std::lround(d4_refcn[ia][j])]++;
            for (int w = 0; w < 23; ++w) {
                double sec = d4_sscale[is - 1] * d4_secaiw[is - 1][w] * zf;
                alpha[ia][j][w] = std::max(d4_ascale[ia][j] * (d4_alphaiw[ia][j][w] - d4_hcount[ia][j] * sec), 0.0);
            }
        }
        for (int j = 0; j < d4_refn[ia]; ++j) {
            int icn = cncount[(int)std::lround(d4_refcn[ia][j])];
            ncount[ia][j] = icn * (icn + 1) / 2;
        }
        have[ia] = true;
    }
    void pair_c6(int zi, int zj, double* out49) {
        static const double freq[23] = {0.000001, 0.050000, 0.100000, 0.200000, 0.300000, 0.400000, 0.500000, 0.600000,
                                        0.700000, 0.800000, 0.900000, 1.000000, 1.200000, 1.400000, 1.600000, 1.800000,
                                        2.000000, 2.500000, 3.000000, 4.000000, 5.000000, 7.500000, 10.00000};
        ensure(zi); ensure(zj);
        double w[23];
        w[0] = 0.5 * (freq[1] - freq[0]);
        for (int k = 1; k < 22; ++k) w[k] = 0.5 * ((freq[k] - freq[k - 1]) + (freq[k + 1] - freq[k]));
        w[22] = 0.5 * (freq[22] - freq[21]);
        for (int k = 0; k < 49; ++k) out49[k] = 0.0;
        for (int ii = 0; ii < d4_refn[zi - 1]; ++ii)
            for (int jj = 0; jj < d4_refn[zj - 1]; ++jj) {
                double s = 0;
                for (int k = 0; k < 23; ++k) s += alpha[zi - 1][ii][k] * alpha[zj - 1][jj][k] * w[k];
                out49[ii * 7 + jj] = 3.0 / kPi * s;
            }
    }
};

// Flat host-side description of a batch
struct BatchSystem {
    std::vector<MolDesc> mol;
    std::vector<AtomPar> atoms;
    std::vector<ShellPar> shells;
    std::vector<RefPar> refs;
    std::vector<int> ao2sh, ao2at;      // per AO (indices within molecule)
    std::vector<int> etypes;            // element numbers of the dense types
    std::vector<double> c6ref;          // [ntype][ntype][7][7]
    long long tot_mat = 0, tot_nn = 0, tot_ss = 0, tot_dd = 0, tot_br = 0, tot_q = 0, tot_hj = 0;
    int max_nat = 0, max_nsh = 0, max_nao = 0, max_ndim = 0, max_nbr = 0;
    int maxhist = 0;

    int add_type(int Z) {
        for (size_t i = 0; i < etypes.size(); ++i) if (etypes[i] == Z) return (int)i;
        etypes.push_back(Z);
        return (int)etypes.size() - 1;
    }

    // append one molecule; throws on unsupported input
    void add_molecule(int nat, const int* Z, double charge, int uhf, int maxhist_) {
        if (nat <= 0) throw std::runtime_error("Number of atoms must be positive");
        maxhist = maxhist_;
        MolDesc m;
        std::memset(&m, 0, sizeof(m));
        m.nat = nat;
        m.at0 = (int)atoms.size(); m.sh0 = (int)shells.size(); m.ao0 = (int)ao2sh.size(); m.ref0 = (int)refs.size();
        m.charge = charge;
        int nsh = 0, nao = 0, ndim = 0;
        double sumz = 0;
        for (int i = 0; i < nat; ++i) {
            int z = Z[i];
            if (z < 1 || z > GFN2_MAX_ELEM) throw std::runtime_error("GFN2-xTB supports elements 1-86 only");
            AtomPar a;
            std::memset(&a, 0, sizeof(a));
            a.Z = z; a.etype = add_type(z); a.nref = d4_refn[z - 1];
            a.sh0 = nsh; a.nsh = gfn2_nshell[z - 1]; a.ao0 = nao; a.ref0 = ndim;
            a.zval = valence_z(z); sumz += a.zval;
            a.rep_alpha = gfn2_rep_alpha[z - 1]; a.rep_zeff = gfn2_rep_zeff[z - 1];
            a.en = param_pauling_en[z - 1];
            a.rcov = param_covalent_rad_d3_aa[z - 1] * kAAtoAU * 4.0 / 3.0;
            a.rad = kAAtoAU * param_atomic_rad_aa[z - 1];
            a.r4r2 = std::sqrt(0.5 * (param_r4_over_r2[z - 1] * std::sqrt((double)z)));
            a.dipk = gfn2_dip_kernel[z - 1]; a.quadk = gfn2_quad_kernel[z - 1];
            a.multirad = xtb_multi_rad[z - 1]; a.valcn = xtb_valence_cn[z - 1];
            a.eeq_en = eeq2019_en[z - 1]; a.eeq_gam = eeq2019_gam[z - 1];
            a.eeq_kappa = eeq2019_kappa[z - 1]; a.eeq_alpha = eeq2019_alpha[z - 1];
            a.nc_rcov = 4.0 / 3.0 * ncoord_rad_aa[z - 1] / 0.52917726; a.nc_en = ncoord_en[z - 1];
            a.d4_zeff = d4_zeff[z - 1]; a.d4_gam = d4_gam[z - 1];
            d4ref().ensure(z);
            double ntot = -1.0e-6;
            for (int s = 0; s < a.nsh; ++s) {
                ShellPar sp;
                std::memset(&sp, 0, sizeof(sp));
                sp.atom = i; sp.l = gfn2_ang_shell[z - 1][s]; sp.nprim = gfn2_nprim(z, s); sp.ao0 = nao;
                sp.zeta = gfn2_slater_exponent[z - 1][s]; sp.level = gfn2_self_energy[z - 1][s];
                sp.kcn = gfn2_kcn_l[z - 1][sp.l]; sp.poly = gfn2_shell_poly_l[z - 1][sp.l];
                sp.hard = gfn2_chemical_hardness[z - 1] * (1.0 + gfn2_shell_hardness_l[z - 1][sp.l]);
                static const double gam3shell[4] = {1.0, 0.5, 0.25, 0.25};  // same for both kinds, gfn2.f90:44-46
                sp.gam3 = gfn2_third_order_atom[z - 1] * gam3shell[sp.l];
                double zs = gfn2_reference_occ_l[z - 1][sp.l];  // setzshell, scc_core.f90:1509-1542
                ntot += zs;
                if (ntot > a.zval) zs = 0.0;
                sp.refocc = zs;
                sto_ng(sp.nprim, gfn2_pqn[z - 1][s], sp.l, sp.zeta, sp.alp, sp.cont);
                int na = sp.l == 0 ? 1 : sp.l == 1 ? 3 : 5;
                for (int k = 0; k < na; ++k) { ao2sh.push_back(nsh); ao2at.push_back(i); }
                nao += na; nsh++;
                shells.push_back(sp);
            }
            a.nao = nao - a.ao0;
            for (int r = 0; r < a.nref; ++r) {
                RefPar rp;
                rp.atom = i; rp.iref = r; rp.ncount = d4ref().ncount[z - 1][r]; rp.pad = 0;
                rp.covcn = d4_refcovcn[z - 1][r]; rp.refq = d4_refq[z - 1][r];
                refs.push_back(rp);
            }
            ndim += a.nref;
            atoms.push_back(a);
        }
        m.nsh = nsh; m.nao = nao; m.ndim = ndim; m.nbr = nsh + 9 * nat;
        m.nel = (int)std::lround(sumz - charge);   // scf_module.F90:334
        m.nopen = uhf;
        if (m.nopen == 0 && m.nel % 2 != 0) m.nopen = 1;
        if (m.nel > 2 * nao) throw std::runtime_error("Not enough basis functions for filling orbitals");
        if (m.nel < 0) throw std::runtime_error("Negative number of electrons");
        // alpha/beta electron counts of occu (scc_core.f90:1112-1170)
        if (m.nel % 2 == 0) { m.ihomoa = m.nel / 2 + m.nopen / 2; m.ihomob = m.nel / 2 - m.nopen / 2; }
        else { m.ihomoa = m.nel / 2 + (m.nopen - 1) / 2 + 1; m.ihomob = m.nel / 2 - (m.nopen - 1) / 2; }
        if (m.ihomoa > nao || m.ihomob < 0) throw std::runtime_error("Spin state not representable in this basis");
        m.mat0 = tot_mat; tot_mat += (long long)nao * nao;
        m.nn0 = tot_nn; tot_nn += (long long)nat * nat;
        m.ss0 = tot_ss; tot_ss += (long long)nsh * nsh;
        m.dd0 = tot_dd; tot_dd += (long long)ndim * ndim;
        m.br0 = tot_br; tot_br += (long long)maxhist * m.nbr;
        m.q0 = tot_q; tot_q += m.nbr;
        {
            int nb64 = (nao + 63) / 64;
            if (nb64 & 1) nb64++;
            if (nb64 < 2) nb64 = 2;
            m.hj0 = tot_hj; tot_hj += (long long)(nb64 / 2) * 128 * 128;
        }
        max_nat = std::max(max_nat, nat); max_nsh = std::max(max_nsh, nsh); max_nao = std::max(max_nao, nao);
        max_ndim = std::max(max_ndim, ndim); max_nbr = std::max(max_nbr, m.nbr);
        mol.push_back(m);
    }
    void finalize() {
        int nt = (int)etypes.size();
        c6ref.assign((size_t)nt * nt * 49, 0.0);
        for (int a = 0; a < nt; ++a)
            for (int b = 0; b < nt; ++b) d4ref().pair_c6(etypes[a], etypes[b], &c6ref[((size_t)a * nt + b) * 49]);
    }
    static D4Reference& d4ref() { static D4Reference r; return r; }
};

}  // namespace xtbb200
