// =====================================================================================
//  oracle/gfn2_oracle.cpp  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
//  CPU restatement (C++/OpenMP + host LAPACK) of the GFN2-xTB single point + gradient
//  path of grimme-lab/xtb v6.7.1.  It is the checker the CUDA engine in xtb_b200/ is
//  compared against and the "port" CPU baseline timed by bench.py.  Only tests/,
//  __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
//
//  Parity pin: validated against the reference's own known-answer tests
//  (tests/test_oracle_golden.py: test/unit/test_gfn2.f90, test/api/c_api_example.c,
//  test/unit/test_coulomb.f90, test_repulsion.f90, test_eeq.f90 values).
//
//  Every routine cites the reference file:line it follows (paths relative to the
//  reference root).  LAPACK/BLAS calls are the ones the reference makes
//  (dpotrf, dsygst, dsyevd, dtrsm, dgemm, dsymv, dsysv, dgetrf, dgetri), resolved from
//  the OpenBLAS that ships in the scipy wheel (symbols prefixed scipy_).
// =====================================================================================
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <string>
#include <map>
#include <algorithm>
#include <chrono>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../xtb_b200/csrc/tables/gfn2_tables.h"
#include "../xtb_b200/csrc/tables/slater_tables.h"
#include "../xtb_b200/csrc/tables/eeq_tables.h"
#include "../xtb_b200/csrc/tables/d4_tables.h"
#include "../xtb_b200/csrc/tables/solv_tables.h"

extern "C" {
void scipy_dpotrf_(const char*, const int*, double*, const int*, int*);
void scipy_dsygst_(const int*, const char*, const int*, double*, const int*, const double*, const int*, int*);
void scipy_dsyevd_(const char*, const char*, const int*, double*, const int*, double*, double*, const int*, int*, const int*, int*);
void scipy_dtrsm_(const char*, const char*, const char*, const char*, const int*, const int*, const double*, const double*, const int*, double*, const int*);
void scipy_dgemm_(const char*, const char*, const int*, const int*, const int*, const double*, const double*, const int*, const double*, const int*, const double*, double*, const int*);
void scipy_dsymv_(const char*, const int*, const double*, const double*, const int*, const double*, const int*, const double*, double*, const int*);
void scipy_dsysv_(const char*, const int*, const int*, double*, const int*, int*, double*, const int*, double*, const int*, int*);
void scipy_dgetrf_(const int*, const int*, double*, const int*, int*, int*);
void scipy_dgetri_(const int*, double*, const int*, const int*, double*, const int*, int*);
void scipy_openblas_set_num_threads(int);
}

namespace {

// src/mctc/convert.f90:25-29, src/mctc/mctc_constants.f90:22,34
const double PI = 3.1415926535897932384626433832795029;
const double AUTOAA = 0.52917726;
const double AATOAU = 1.0 / AUTOAA;
const double AUTOEV = 27.21138505;
const double EVTOAU = 1.0 / AUTOEV;
const double KB = 3.166808578545117e-06;

typedef std::vector<double> dvec;
typedef std::vector<int> ivec;

double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ------------------------------------------------------------------------------------
// element level data (src/xtb/gfn2.f90:731-967 initData/initHamiltonian/initCoulomb)
// ------------------------------------------------------------------------------------
inline int el_nshell(int Z) { return gfn2_nshell[Z - 1]; }
inline int el_ang(int Z, int ish) { return gfn2_ang_shell[Z - 1][ish]; }
inline double el_atomic_rad(int Z) { return AATOAU * param_atomic_rad_aa[Z - 1]; }       // atomicrad.f90:42
inline double el_cov_rad_d3(int Z) { return param_covalent_rad_d3_aa[Z - 1] * AATOAU * 4.0 / 3.0; } // covalentradd3.f90:38-62
inline double el_sqrt_zr4r2(int Z) { return std::sqrt(0.5 * (param_r4_over_r2[Z - 1] * std::sqrt((double)Z))); } // sqrtzr4r2.f90:68
inline double el_ncoord_rcov(int Z) { return 4.0 / 3.0 * ncoord_rad_aa[Z - 1] / 0.52917726; } // ncoord.f90:87

// number of primitives, src/xtb/gfn2.f90:893-937 (setGFN2NumberOfPrimitives)
int el_nprim(int Z, int ish) {
    int l = el_ang(Z, ish), n = gfn2_pqn[Z - 1][ish];
    if (Z <= 2) return l == 0 ? 3 : 4;
    if (l == 0 || l == 1) return n > 5 ? 6 : 4;
    if (l == 2) return 3;
    return 4;
}
// shell hardness, src/xtb/gfn1.f90:845-871 (setGFN1ShellHardness called at gfn2.f90:791)
inline double el_shell_hardness(int Z, int ish) {
    return gfn2_chemical_hardness[Z - 1] * (1.0 + gfn2_shell_hardness_l[Z - 1][el_ang(Z, ish)]);
}
// third order shell gam, src/xtb/gfn2.f90:940-967; gam3shell(kind, l) gfn2.f90:44-46
inline double el_third_order_shell(int Z, int ish) {
    static const double gam3shell[2][4] = {{1.0, 0.5, 0.25, 0.25}, {1.0, 0.5, 0.25, 0.25}};
    return gfn2_third_order_atom[Z - 1] * gam3shell[gfn2_kinds[Z - 1] - 1][el_ang(Z, ish)];
}
// valence electrons, src/type/molecule.f90:560-597
double el_valence_z(int Z) {
    int nc = 0;
    if (Z <= 2) nc = 0; else if (Z <= 10) nc = 2; else if (Z <= 18) nc = 10; else if (Z <= 29) nc = 18;
    else if (Z <= 36) nc = 28; else if (Z <= 47) nc = 36; else if (Z <= 54) nc = 46; else if (Z <= 71) nc = 54;
    else if (Z <= 79) nc = 68; else if (Z <= 86) nc = 78; else nc = 86;
    double z = (double)Z - (double)nc;
    if (Z > 57 && Z < 72) z = 3.0;
    return z;
}

// ------------------------------------------------------------------------------------
// STO-nG expansion, src/slater.f90:373-483 (slaterToGauss, norm=.true.)
// ------------------------------------------------------------------------------------
void slater_to_gauss(int ng, int n, int l, double zeta, double* alpha, double* coeff) {
    static const double dfact[8] = {1., 1., 3., 15., 105., 945., 10395., 135135.};
    int ityp = (l == 0) ? n : (l == 1) ? 4 + n : (l == 2) ? 7 + n : 9 + n;  // 1-based
    ityp -= 1;
    const double z2 = zeta * zeta;
    for (int p = 0; p < ng; ++p) {
        double a = 0, c = 0;
        switch (ng) {
            case 1: a = sto_alpha1[ityp]; c = 1.0; break;
            case 2: a = sto_alpha2[ityp][p]; c = sto_coeff2[ityp][p]; break;
            case 3: a = sto_alpha3[ityp][p]; c = sto_coeff3[ityp][p]; break;
            case 4: a = sto_alpha4[ityp][p]; c = sto_coeff4[ityp][p]; break;
            case 5: a = sto_alpha5[ityp][p]; c = sto_coeff5[ityp][p]; break;
            case 6:
                if (n == 6) {
                    if (l == 0) { a = sto_alpha6s[p]; c = sto_coeff6s[p]; }
                    else { a = sto_alpha6p[p]; c = sto_coeff6p[p]; }
                } else { a = sto_alpha6[ityp][p]; c = sto_coeff6[ityp][p]; }
                break;
        }
        alpha[p] = a * z2;
        coeff[p] = c;
    }
    const double top = 2.0 / PI;
    for (int p = 0; p < ng; ++p)
        coeff[p] = coeff[p] * std::pow(top * alpha[p], 0.75) * std::pow(std::sqrt(4 * alpha[p]), (double)l) /
                   std::sqrt(dfact[l]);
}

// ------------------------------------------------------------------------------------
// basis set, src/xtb/basisset.f90:33-308 (newBasisset) -- only the valence branches are
// reachable for GFN2 (every shell is a valence shell, see src/xtb/data.f90:469-495)
// ------------------------------------------------------------------------------------
struct Shell {
    int atom, l, nprim, ao0, nao;  // ao0: first spherical AO, nao: 1/3/5
    double alp[6], cont[6];        // radial part; Cartesian d factors applied per component
    double zeta, level;
    int ish_el;                    // shell index within the element
};
struct Basis {
    int nat = 0, nshell = 0, nao = 0;
    std::vector<Shell> sh;
    ivec sh0, nsh;      // first shell / number of shells per atom
    ivec ao2sh, ao2at;  // per spherical AO
    ivec ao0_at, nao_at;  // fila2
};
const int NCART[3] = {1, 3, 6};
const int NSPH[3] = {1, 3, 5};
// Cartesian exponents src/intgrad.f90:30-55 in order s | px py pz | dxx dyy dzz dxy dxz dyz
const int LX[10] = {0, 1, 0, 0, 2, 0, 0, 1, 1, 0};
const int LY[10] = {0, 0, 1, 0, 0, 2, 0, 1, 0, 1};
const int LZ[10] = {0, 0, 0, 1, 0, 0, 2, 0, 1, 1};
const int ITT[3] = {0, 1, 4};
// src/xtb/basisset.f90:512-513 : contraction prefactors of the six Cartesian d functions
const double DCART[6] = {1.0, 1.0, 1.0, 1.7320508075688772935, 1.7320508075688772935, 1.7320508075688772935};

void build_basis(int nat, const int* Z, Basis& b) {
    b = Basis();
    b.nat = nat;
    b.sh0.resize(nat); b.nsh.resize(nat); b.ao0_at.resize(nat); b.nao_at.resize(nat);
    for (int i = 0; i < nat; ++i) {
        int z = Z[i];
        b.sh0[i] = b.nshell; b.nsh[i] = el_nshell(z); b.ao0_at[i] = b.nao;
        for (int m = 0; m < el_nshell(z); ++m) {
            Shell s;
            s.atom = i; s.l = el_ang(z, m); s.nprim = el_nprim(z, m); s.ish_el = m;
            s.zeta = gfn2_slater_exponent[z - 1][m];
            s.level = gfn2_self_energy[z - 1][m];
            slater_to_gauss(s.nprim, gfn2_pqn[z - 1][m], s.l, s.zeta, s.alp, s.cont);
            s.ao0 = b.nao; s.nao = NSPH[s.l];
            for (int k = 0; k < s.nao; ++k) { b.ao2sh.push_back(b.nshell); b.ao2at.push_back(i); }
            b.nao += s.nao;
            b.sh.push_back(s);
            b.nshell++;
        }
        b.nao_at[i] = b.nao - b.ao0_at[i];
    }
}

// ------------------------------------------------------------------------------------
// primitive integrals, src/intgrad.f90
// ------------------------------------------------------------------------------------
// olapp, intgrad.f90:95-111
inline double olapp(int l, double gama) {
    static const double dftr[8] = {1., 1., 3., 15., 105., 945., 10395., 135135.};
    if (l % 2) return 0.0;
    int lh = l / 2;
    return std::pow(0.5 / gama, lh) * dftr[lh];
}
// horizontal_shift, intgrad.f90:357-380 (0-based coefficient arrays)
inline void horizontal_shift(double ae, int l, double* c) {
    switch (l) {
        case 1: c[0] += ae * c[1]; break;
        case 2: c[0] += ae * ae * c[2]; c[1] += 2 * ae * c[2]; break;
        case 3: c[0] += ae * ae * ae * c[3]; c[1] += 3 * ae * ae * c[3]; c[2] += 3 * ae * c[3]; break;
        case 4: c[0] += ae * ae * ae * ae * c[4]; c[1] += 4 * ae * ae * ae * c[4]; c[2] += 6 * ae * ae * c[4]; c[3] += 4 * ae * c[4]; break;
        default: break;
    }
}
// form_product, intgrad.f90:382-471: polynomial product for la,lb <= 3
inline void form_product(const double* a, const double* b, int la, int lb, double* d) {
    for (int k = 0; k <= la + lb; ++k) {
        double s = 0.0;
        for (int i = 0; i <= k; ++i) if (i <= la && k - i <= lb) s += a[i] * b[k - i];
        d[k] = s;
    }
}
// multipole_3d, intgrad.f90:509-555: s3d[0]=S, [1..3]=x,y,z, [4..9]=xx,yy,zz,xy,xz,yz
inline void multipole_3d(const double* ri, const double* rj, const double* rc, const double* rp, const int* li,
                         const int* lj, const double* s1d, double* s3d) {
    double val[3][3];
    for (int k = 0; k < 3; ++k) {
        double vi[8] = {0}, vj[8] = {0}, vv[16] = {0};
        vi[li[k]] = 1.0; vj[lj[k]] = 1.0;
        double rpc = rp[k] - rc[k];
        horizontal_shift(rp[k] - ri[k], li[k], vi);
        horizontal_shift(rp[k] - rj[k], lj[k], vj);
        form_product(vi, vj, li[k], lj[k], vv);
        val[k][0] = val[k][1] = val[k][2] = 0.0;
        for (int l = 0; l <= li[k] + lj[k]; ++l) {
            val[k][0] += s1d[l] * vv[l];
            val[k][1] += (s1d[l + 1] + rpc * s1d[l]) * vv[l];
            val[k][2] += (s1d[l + 2] + 2 * rpc * s1d[l + 1] + rpc * rpc * s1d[l]) * vv[l];
        }
    }
    s3d[0] = val[0][0] * val[1][0] * val[2][0];
    s3d[1] = val[0][1] * val[1][0] * val[2][0];
    s3d[2] = val[0][0] * val[1][1] * val[2][0];
    s3d[3] = val[0][0] * val[1][0] * val[2][1];
    s3d[4] = val[0][2] * val[1][0] * val[2][0];
    s3d[5] = val[0][0] * val[1][2] * val[2][0];
    s3d[6] = val[0][0] * val[1][0] * val[2][2];
    s3d[7] = val[0][1] * val[1][1] * val[2][0];
    s3d[8] = val[0][1] * val[1][0] * val[2][1];
    s3d[9] = val[0][0] * val[1][1] * val[2][1];
}
// multipole_grad_3d, intgrad.f90:607-697: derivative w.r.t. centre i, ds3d[c][comp]
inline void multipole_grad_3d(const double* ri, const double* rj, const double* rc, const double* rp, double ai,
                              const int* li, const int* lj, const double* s1d, double* s3d, double ds3d[3][10]) {
    double val[3][3], gra[3][3];
    for (int k = 0; k < 3; ++k) {
        double vi[8] = {0}, vj[8] = {0}, gi[8] = {0}, vv[16] = {0}, gg[16] = {0};
        double rpc = rp[k] - rc[k];
        vi[li[k]] = 1.0; vj[lj[k]] = 1.0;
        gi[li[k] + 1] = 2 * ai;
        if (li[k] > 0) gi[li[k] - 1] = -li[k];
        horizontal_shift(rp[k] - ri[k], li[k] - 1, gi);
        horizontal_shift(rp[k] - ri[k], li[k] + 1, gi);
        horizontal_shift(rp[k] - ri[k], li[k], vi);
        horizontal_shift(rp[k] - rj[k], lj[k], vj);
        form_product(vi, vj, li[k], lj[k], vv);
        form_product(gi, vj, li[k] + 1, lj[k], gg);
        for (int c = 0; c < 3; ++c) val[k][c] = gra[k][c] = 0.0;
        for (int l = 0; l <= li[k] + lj[k] + 1; ++l) {
            double m0 = s1d[l], m1 = s1d[l + 1] + rpc * s1d[l], m2 = s1d[l + 2] + 2 * rpc * s1d[l + 1] + rpc * rpc * s1d[l];
            val[k][0] += m0 * vv[l]; val[k][1] += m1 * vv[l]; val[k][2] += m2 * vv[l];
            gra[k][0] += m0 * gg[l]; gra[k][1] += m1 * gg[l]; gra[k][2] += m2 * gg[l];
        }
    }
    // component -> (order in x, order in y, order in z)
    static const int ord[10][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {2, 0, 0}, {0, 2, 0}, {0, 0, 2}, {1, 1, 0}, {1, 0, 1}, {0, 1, 1}};
    for (int c = 0; c < 10; ++c) {
        const int* o = ord[c];
        s3d[c] = val[0][o[0]] * val[1][o[1]] * val[2][o[2]];
        ds3d[0][c] = gra[0][o[0]] * val[1][o[1]] * val[2][o[2]];
        ds3d[1][c] = val[0][o[0]] * gra[1][o[1]] * val[2][o[2]];
        ds3d[2][c] = val[0][o[0]] * val[1][o[1]] * gra[2][o[2]];
    }
}

// CAO -> SAO transformation of one (6 x 6) block, src/intgrad.f90:144-231 (dtrf2).
// blk[mi][mj] (i = first index here); d indices are contracted with trafo columns 2..6.
const double S3H = 0.86602540378443864676;  // 0.5*sqrt(3)
const double DTRAFO[6][5] = {  // [cart m][sph a]  (columns 2..6 of trafo, intgrad.f90:70-86)
    {S3H, 0.5, 0, 0, 0}, {-S3H, 0.5, 0, 0, 0}, {0, -1.0, 0, 0, 0}, {0, 0, 1, 0, 0}, {0, 0, 0, 1, 0}, {0, 0, 0, 0, 1}};
inline void cao2sao(double blk[6][6], int li, int lj) {
    if (li < 2 && lj < 2) return;
    double t[6][6];
    if (lj == 2) {
        int ni = NCART[li];
        for (int i = 0; i < ni; ++i)
            for (int a = 0; a < 5; ++a) {
                double s = 0;
                for (int m = 0; m < 6; ++m) s += DTRAFO[m][a] * blk[i][m];
                t[i][a] = s;
            }
        for (int i = 0; i < ni; ++i) for (int a = 0; a < 5; ++a) blk[i][a] = t[i][a];
    }
    if (li == 2) {
        int nj = NSPH[lj];
        for (int a = 0; a < 5; ++a)
            for (int j = 0; j < nj; ++j) {
                double s = 0;
                for (int m = 0; m < 6; ++m) s += DTRAFO[m][a] * blk[m][j];
                t[a][j] = s;
            }
        for (int a = 0; a < 5; ++a) for (int j = 0; j < nj; ++j) blk[a][j] = t[a][j];
    }
}

// Counters of the discrete decisions of one single point (same numbering as
// include/xtb_b200.h xtbb200_getDecisionCounts), incremented where the decisions are taken; unordered pairs / triples.
long long g_cnt[12];
inline void cnt_add(int k, long long v) {
#pragma omp atomic
    g_cnt[k] += v;
}

// get_multiints, intgrad.f90:879-955.  out[c][mi][mj], c = 0..9 (S, D(3), Q(6)), Cartesian.
void get_multiints(const Shell& si, const Shell& sj, const double* ri, const double* rj, const double* point,
                   double intcut, double out[10][6][6]) {
    for (int c = 0; c < 10; ++c) for (int a = 0; a < 6; ++a) for (int b = 0; b < 6; ++b) out[c][a][b] = 0.0;
    double rij2 = 0;
    for (int k = 0; k < 3; ++k) rij2 += (ri[k] - rj[k]) * (ri[k] - rj[k]);
    if (rij2 > 2000.0) return;
    cnt_add(0, 1);
    const double sqrtpi = std::sqrt(PI);
    const int naoi = NCART[si.l], naoj = NCART[sj.l];
    for (int ip = 0; ip < si.nprim; ++ip) {
        double alpi = si.alp[ip];
        for (int jp = 0; jp < sj.nprim; ++jp) {
            double alpj = sj.alp[jp];
            double est = alpi * alpj * rij2;
            if (est > intcut * (alpi + alpj)) continue;
            cnt_add(1, 1);
            double ab = 1.0 / (alpi + alpj);
            est = est * ab;
            double sq = sqrtpi * std::sqrt(ab);
            double kab = std::exp(-est) * (sq * sq * sq);
            double rp[3];
            for (int k = 0; k < 3; ++k) rp[k] = (alpi * ri[k] + alpj * rj[k]) * ab;
            double t[12] = {0};
            for (int k = 0; k <= si.l + sj.l + 2; ++k) t[k] = olapp(k, alpi + alpj);
            for (int mli = 0; mli < naoi; ++mli) {
                double ci = si.cont[ip] * (si.l == 2 ? DCART[mli] : 1.0);
                int li[3] = {LX[ITT[si.l] + mli], LY[ITT[si.l] + mli], LZ[ITT[si.l] + mli]};
                for (int mlj = 0; mlj < naoj; ++mlj) {
                    double cj = sj.cont[jp] * (sj.l == 2 ? DCART[mlj] : 1.0);
                    double cc = kab * cj * ci;
                    int lj[3] = {LX[ITT[sj.l] + mlj], LY[ITT[sj.l] + mlj], LZ[ITT[sj.l] + mlj]};
                    double saw[10];
                    multipole_3d(ri, rj, point, rp, li, lj, t, saw);
                    for (int c = 0; c < 10; ++c) out[c][mli][mlj] += saw[c] * cc;
                }
            }
        }
    }
}

// shiftintg, intgrad.f90:703-728: g[c][k] c = 0..18, k = xyz
inline void shiftintg(double g[19][3], const double* s, const double* r) {
    for (int k = 0; k < 3; ++k) {
        g[10][k] = g[1][k] - r[0] * g[0][k];
        g[11][k] = g[2][k] - r[1] * g[0][k];
        g[12][k] = g[3][k] - r[2] * g[0][k];
        g[13][k] = g[4][k] - 2 * r[0] * g[1][k] + r[0] * r[0] * g[0][k];
        g[14][k] = g[5][k] - 2 * r[1] * g[2][k] + r[1] * r[1] * g[0][k];
        g[15][k] = g[6][k] - 2 * r[2] * g[3][k] + r[2] * r[2] * g[0][k];
        g[16][k] = g[7][k] - r[0] * g[2][k] - r[1] * g[1][k] + r[0] * r[1] * g[0][k];
        g[17][k] = g[8][k] - r[0] * g[3][k] - r[2] * g[1][k] + r[0] * r[2] * g[0][k];
        g[18][k] = g[9][k] - r[1] * g[3][k] - r[2] * g[2][k] + r[1] * r[2] * g[0][k];
    }
    g[10][0] -= s[0]; g[11][1] -= s[0]; g[12][2] -= s[0];
    g[13][0] += -2 * s[1] + 2 * r[0] * s[0];
    g[14][1] += -2 * s[2] + 2 * r[1] * s[0];
    g[15][2] += -2 * s[3] + 2 * r[2] * s[0];
    g[16][0] += -s[2] + r[1] * s[0];
    g[16][1] += -s[1] + r[0] * s[0];
    g[17][0] += -s[3] + r[2] * s[0];
    g[17][2] += -s[1] + r[0] * s[0];
    g[18][1] += -s[3] + r[2] * s[0];
    g[18][2] += -s[2] + r[1] * s[0];
}

// get_grad_multiint, intgrad.f90:958-1035.  sdqg[c][k][mi][mj] (c = 0..18), Cartesian
void get_grad_multiint(const Shell& si, const Shell& sj, const double* ri, const double* rj, double intcut,
                       double sdqg[19][3][6][6]) {
    double sdq[10][6][6];
    for (int c = 0; c < 19; ++c) for (int k = 0; k < 3; ++k) for (int a = 0; a < 6; ++a) for (int b = 0; b < 6; ++b) sdqg[c][k][a][b] = 0;
    for (int c = 0; c < 10; ++c) for (int a = 0; a < 6; ++a) for (int b = 0; b < 6; ++b) sdq[c][a][b] = 0;
    double rij[3], rij2 = 0;
    for (int k = 0; k < 3; ++k) { rij[k] = ri[k] - rj[k]; rij2 += rij[k] * rij[k]; }
    if (rij2 > 2000.0) return;
    const double sqrtpi = std::sqrt(PI);
    const int naoi = NCART[si.l], naoj = NCART[sj.l];
    for (int ip = 0; ip < si.nprim; ++ip) {
        double alpi = si.alp[ip];
        for (int jp = 0; jp < sj.nprim; ++jp) {
            double alpj = sj.alp[jp];
            double est = alpi * alpj * rij2;
            if (est > intcut * (alpi + alpj)) continue;
            cnt_add(3, 1);
            double ab = 1.0 / (alpi + alpj);
            est = est * ab;
            double sq = sqrtpi * std::sqrt(ab);
            double kab = std::exp(-est) * (sq * sq * sq);
            double rp[3];
            for (int k = 0; k < 3; ++k) rp[k] = (alpi * ri[k] + alpj * rj[k]) * ab;
            double t[12] = {0};
            for (int k = 0; k <= si.l + sj.l + 3; ++k) t[k] = olapp(k, alpi + alpj);
            for (int mli = 0; mli < naoi; ++mli) {
                double ci = si.cont[ip] * (si.l == 2 ? DCART[mli] : 1.0);
                int li[3] = {LX[ITT[si.l] + mli], LY[ITT[si.l] + mli], LZ[ITT[si.l] + mli]};
                for (int mlj = 0; mlj < naoj; ++mlj) {
                    double cj = sj.cont[jp] * (sj.l == 2 ? DCART[mlj] : 1.0);
                    double cc = kab * cj * ci;
                    int lj[3] = {LX[ITT[sj.l] + mlj], LY[ITT[sj.l] + mlj], LZ[ITT[sj.l] + mlj]};
                    double saw[10], sawg[3][10];
                    multipole_grad_3d(ri, rj, rj, rp, alpi, li, lj, t, saw, sawg);
                    for (int c = 0; c < 10; ++c) {
                        sdq[c][mli][mlj] += saw[c] * cc;
                        for (int k = 0; k < 3; ++k) sdqg[c][k][mli][mlj] += sawg[k][c] * cc;
                    }
                }
            }
        }
    }
    for (int mli = 0; mli < naoi; ++mli)
        for (int mlj = 0; mlj < naoj; ++mlj) {
            double g[19][3], s[10];
            for (int c = 0; c < 10; ++c) { s[c] = sdq[c][mli][mlj]; for (int k = 0; k < 3; ++k) g[c][k] = sdqg[c][k][mli][mlj]; }
            shiftintg(g, s, rij);
            for (int c = 10; c < 19; ++c) for (int k = 0; k < 3; ++k) sdqg[c][k][mli][mlj] = g[c][k];
        }
}

// h0scal, src/scc_core.f90:644-680 (valence-valence branch; GFN2 has only valence shells)
// kScale / enScale from src/xtb/gfn2.f90:834-853
inline double h0scal(int li, int lj, int zi, int zj) {
    static const double kshell[4] = {1.85, 2.23, 2.23, 2.23};
    double ks = 0.5 * (kshell[li] + kshell[lj]);
    if ((li == 0 && lj == 2) || (li == 2 && lj == 0)) ks = 2.0;  // ksd
    if ((li == 1 && lj == 2) || (li == 2 && lj == 1)) ks = 2.0;  // kpd
    double enscale = 0.005 * (2.0 + 2.0);
    double den = (param_pauling_en[zi - 1] - param_pauling_en[zj - 1]);
    den = den * den;
    double enpoly = (1.0 + enscale * den * (1.0 + 0.0 * den));
    return ks * enpoly * 1.0;
}
// shellPoly, scc_core.f90:728-754
inline double shell_poly(double ipoly, double jpoly, double irad, double jrad, const double* a, const double* b) {
    double dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
    double rab = std::sqrt(dx * dx + dy * dy + dz * dz);
    double rr = jrad + irad;
    double r = rab / rr;
    double rf1 = 1.0 + 0.01 * ipoly * std::sqrt(r);
    double rf2 = 1.0 + 0.01 * jpoly * std::sqrt(r);
    return rf1 * rf2;
}

// ------------------------------------------------------------------------------------
// state of one calculation
// ------------------------------------------------------------------------------------
struct Oracle {
    int nat = 0;
    ivec Z;
    dvec xyz;  // 3*nat, [atom][k]
    double charge = 0;
    int uhf = 0;
    double acc = 1.0, etemp = 300.0;
    int maxiter = 250;
    int guess = 0;  // 0 EEQ, 1 zero charges, 2 restart arrays
    double broydamp = 0.4;  // src/setparam.f90:226
    // external point charges (src/type/pcem.f90; set through xtb_setExternalCharges, src/api/calculator.f90:470-540):
    // positions [npc][3] in bohr, charges, hardness gam = chemical hardness of the element given for the charge
    dvec pc_xyz, pc_q, pc_gam;
    // implicit solvation (xtb_setSolvent -> TSolvInput, src/solv/input.F90:30-66): solv_on, normalised solvent name,
    // ALPB flag, kernel (1 Still, 2 P16), reference state (1 gsolv, 2 reference, 3 mol1bar), temperature, ion strength / radius
    bool solv_on = false;
    std::string solv_name;
    bool solv_alpb = false;
    int solv_kernel = 2, solv_state = 1;
    double solv_temp = 298.15, solv_ionst = 0.0, solv_ionrad = 0.0;

    Basis bas;
    // results / intermediates, by name
    std::map<std::string, dvec> d;
    std::map<std::string, ivec> iv;
    double timers[8] = {0};
    int nthreads = 1;
};

inline const double* R(const Oracle& o, int i) { return &o.xyz[3 * i]; }

// ------------------------------------------------------------------------------------
// coordination numbers
// ------------------------------------------------------------------------------------
struct CNPair { int i, j; double d[3]; };  // d = dcount/dR_i (i > j)
// src/disp/coordinationnumber.F90:351-479 (ncoordLatP) with the gfn (513-...) or erf+EN counting
void cn_lp(const Oracle& o, bool gfn, dvec& cn, std::vector<CNPair>& pairs) {
    const double k4 = 4.10451, k5 = 19.08857, k6 = 2 * 11.28174 * 11.28174;
    const double kcn = gfn ? 10.0 : 7.5, cutoff2 = 1600.0;
    const double sqrtpi = std::sqrt(PI);
    int n = o.nat;
    cn.assign(n, 0.0);
    pairs.clear();
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < i; ++j) {
            double rij[3], r2 = 0;
            for (int k = 0; k < 3; ++k) { rij[k] = R(o, i)[k] - R(o, j)[k]; r2 += rij[k] * rij[k]; }
            if (r2 > cutoff2 || r2 < 1.0e-12) continue;
            if (gfn) cnt_add(4, 1);
            double r1 = std::sqrt(r2);
            int zi = o.Z[i], zj = o.Z[j];
            double rc = el_cov_rad_d3(zi) + el_cov_rad_d3(zj);
            double den = 1.0, cf, df;
            if (gfn) {
                // gfnCount = expCount(k,r,r0)*expCount(2k,r,r0+2)
                auto expc = [](double k, double r, double r0) { return 1.0 / (1.0 + std::exp(-k * (r0 / r - 1.0))); };
                auto dexpc = [](double k, double r, double r0) {
                    double e = std::exp(-k * (r0 / r - 1.0));
                    return (-k * r0 * e) / (r * r * ((e + 1.0) * (e + 1.0)));
                };
                cf = expc(kcn, r1, rc) * expc(2 * kcn, r1, rc + 2);
                df = dexpc(kcn, r1, rc) * expc(2 * kcn, r1, rc + 2) + expc(kcn, r1, rc) * dexpc(2 * kcn, r1, rc + 2);
            } else {
                double a = std::fabs(param_pauling_en[zi - 1] - param_pauling_en[zj - 1]) + k5;
                den = k4 * std::exp(-(a * a) / k6);
                cf = 0.5 * (1.0 + std::erf(-kcn * (r1 - rc) / rc));
                df = -kcn / sqrtpi / rc * std::exp(-kcn * kcn * (r1 - rc) * (r1 - rc) / (rc * rc));
            }
            double countf = den * cf;
            cn[i] += countf; cn[j] += countf;
            CNPair p; p.i = i; p.j = j;
            for (int k = 0; k < 3; ++k) p.d[k] = den * df * rij[k] / r1;
            pairs.push_back(p);
        }
}
// g(:,i) += sum_j dcndr(:,i,j) v(j), the dense gemv of src/scf_module.F90:726 written over pairs
// (dcndr(:,i,i)+=d, dcndr(:,j,j)-=d, dcndr(:,i,j)+=d, dcndr(:,j,i)-=d; coordinationnumber.F90:437-443)
void cn_chain(const std::vector<CNPair>& pairs, const dvec& v, dvec& g) {
    for (const CNPair& p : pairs) {
        double w = v[p.i] + v[p.j];
        for (int k = 0; k < 3; ++k) { g[3 * p.i + k] += p.d[k] * w; g[3 * p.j + k] -= p.d[k] * w; }
    }
}
// src/disp/ncoord.f90:469-506 (ncoord_erf) and 575-617 (ncoord_d4); single precision radii/EN tables
void ncoord_erf_d4(const Oracle& o, bool d4, dvec& cn) {
    const double k4 = 4.10451, k5 = 19.08857, k6 = 2 * 11.28174 * 11.28174, kn = 7.5;
    int n = o.nat;
    cn.assign(n, 0.0);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < i; ++j) {
            double r2 = 0;
            for (int k = 0; k < 3; ++k) { double t = R(o, j)[k] - R(o, i)[k]; r2 += t * t; }
            if (r2 > 1600.0) continue;
            if (d4) cnt_add(5, 1);
            double r = std::sqrt(r2);
            double rco = el_ncoord_rcov(o.Z[j]) + el_ncoord_rcov(o.Z[i]);
            double den = 1.0;
            if (d4) {
                double a = std::fabs(ncoord_en[o.Z[i] - 1] - ncoord_en[o.Z[j] - 1]) + k5;
                den = k4 * std::exp(-(a * a) / k6);
            }
            double tmp = den * 0.5 * (1.0 + std::erf(-kn * (r - rco) / rco));
            cn[i] += tmp; cn[j] += tmp;
        }
}

// ------------------------------------------------------------------------------------
// EEQ guess charges: src/xtb/calculator.f90:432-496, src/eeq_model.f90:58-80,782-870,
// src/charge_model.f90:147-245, src/scc_core.f90:1475-1506 (iniqshell)
// ------------------------------------------------------------------------------------
bool eeq_charges(const Oracle& o, dvec& q) {
    int n = o.nat, m = n + 1;
    dvec cn;
    ncoord_erf_d4(o, false, cn);
    dvec A((size_t)m * m, 0.0), X(m, 0.0);
    const double sqrt2pi = std::sqrt(2.0 / PI);
    for (int i = 0; i < n; ++i) {
        int zi = o.Z[i] - 1;
        double tmp = eeq2019_kappa[zi] / (std::sqrt(cn[i]) + 1e-14);
        X[i] = -eeq2019_en[zi] + tmp * cn[i];
        for (int j = 0; j < i; ++j) {
            int zj = o.Z[j] - 1;
            double r2 = 0;
            for (int k = 0; k < 3; ++k) { double t = R(o, j)[k] - R(o, i)[k]; r2 += t * t; }
            double r1 = std::sqrt(r2);
            double gamij = 1.0 / std::sqrt(eeq2019_alpha[zi] * eeq2019_alpha[zi] + eeq2019_alpha[zj] * eeq2019_alpha[zj]);
            double v = std::erf(gamij * r1) / r1;
            A[j + (size_t)m * i] = v; A[i + (size_t)m * j] = v;
        }
        A[i + (size_t)m * i] = eeq2019_gam[zi] + sqrt2pi / eeq2019_alpha[zi];
    }
    for (int i = 0; i < m; ++i) { A[n + (size_t)m * i] = 1.0; A[i + (size_t)m * n] = 1.0; }
    A[n + (size_t)m * n] = 0.0;
    X[n] = o.charge;
    ivec ipiv(m);
    int one = 1, info = 0, lwork = -1;
    double wq;
    scipy_dsysv_("U", &m, &one, A.data(), &m, ipiv.data(), X.data(), &m, &wq, &lwork, &info);
    lwork = (int)wq;
    dvec work(std::max(1, lwork));
    scipy_dsysv_("U", &m, &one, A.data(), &m, ipiv.data(), X.data(), &m, work.data(), &lwork, &info);
    if (info != 0) return false;
    q.assign(X.begin(), X.begin() + n);
    return true;
}
void iniqshell(const Oracle& o, const dvec& q, dvec& qsh) {
    qsh.assign(o.bas.nshell, 0.0);
    int k = 0;
    for (int i = 0; i < o.nat; ++i) {
        int z = o.Z[i];
        double zv = el_valence_z(z), ntot = -1.0e-6;
        for (int m = 0; m < el_nshell(z); ++m) {
            double zshell = gfn2_reference_occ_l[z - 1][el_ang(z, m)];
            ntot += zshell;
            if (ntot > zv) zshell = 0.0;
            qsh[k++] = zshell / zv * q[i];
        }
    }
}

// ------------------------------------------------------------------------------------
// D4 model: src/disp/dftd4.F90:106-203 (newD4Model, mode p_refq_gfn2xtb => secq = 0,
// q = refq, tmp_hq = refh), 348-379 (zeta/dzeta), 381-449 (trapzd), 451-466 (cngw)
// ------------------------------------------------------------------------------------
inline double d4_zeta(double a, double c, double qref, double qmod) {
    if (qmod < 0.0) return std::exp(a);
    return std::exp(a * (1.0 - std::exp(c * (1.0 - qref / qmod))));
}
inline double d4_dzeta(double a, double c, double qref, double qmod) {
    if (qmod < 0.0) return 0.0;
    return -a * c * std::exp(c * (1.0 - qref / qmod)) * d4_zeta(a, c, qref, qmod) * qref / (qmod * qmod);
}
inline double cngw(double wf, double cn, double cnref) {
    double val = -wf * (cn - cnref) * (cn - cnref);
    return val < -200.0 ? 0.0 : std::exp(val);
}
struct D4Model {
    double alpha[118][7][23];
    int ncount[118][7];
    bool have[118];
    std::map<std::pair<int, int>, std::vector<double>> c6cache;  // (zi,zj) -> [ii*7+jj]
    D4Model() { std::memset(have, 0, sizeof(have)); }
    void ensure(int Z) {
        int ia = Z - 1;
        if (have[ia]) return;
        const double g_a = 3.0, g_c = 2.0;
        int cncount[19] = {0};
        cncount[0] = 1;
        for (int j = 0; j < d4_refn[ia]; ++j) {
            int is = d4_refsys[ia][j];  // 1-based reference system == element number of the partner
            double iz = d4_zeff[is - 1];
            double zf = d4_zeta(g_a, d4_gam[is - 1] * g_c, 0.0 + iz, d4_refh[ia][j] + iz);
            int icn = (int)std::lround(d4_refcn[ia][j]);
            cncount[icn]++;
            for (int w = 0; w < 23; ++w) {
                double sec_al = d4_sscale[is - 1] * d4_secaiw[is - 1][w] * zf;
                alpha[ia][j][w] = std::max(d4_ascale[ia][j] * (d4_alphaiw[ia][j][w] - d4_hcount[ia][j] * sec_al), 0.0);
            }
        }
        for (int j = 0; j < d4_refn[ia]; ++j) {
            int icn = cncount[(int)std::lround(d4_refcn[ia][j])];
            ncount[ia][j] = icn * (icn + 1) / 2;
        }
        have[ia] = true;
    }
    const std::vector<double>& c6(int zi, int zj) {
        auto key = std::make_pair(zi, zj);
        auto it = c6cache.find(key);
        if (it != c6cache.end()) return it->second;
        ensure(zi); ensure(zj);
        static const double freq[23] = {0.000001, 0.050000, 0.100000, 0.200000, 0.300000, 0.400000, 0.500000, 0.600000,
                                        0.700000, 0.800000, 0.900000, 1.000000, 1.200000, 1.400000, 1.600000, 1.800000,
                                        2.000000, 2.500000, 3.000000, 4.000000, 5.000000, 7.500000, 10.00000};
        double weights[23];
        weights[0] = 0.5 * (freq[1] - freq[0]);
        for (int w = 1; w < 22; ++w) weights[w] = 0.5 * ((freq[w] - freq[w - 1]) + (freq[w + 1] - freq[w]));
        weights[22] = 0.5 * (freq[22] - freq[21]);
        const double thopi = 3.0 / PI;
        std::vector<double> v(49, 0.0);
        for (int ii = 0; ii < d4_refn[zi - 1]; ++ii)
            for (int jj = 0; jj < d4_refn[zj - 1]; ++jj) {
                double s = 0;
                for (int w = 0; w < 23; ++w) s += alpha[zi - 1][ii][w] * alpha[zj - 1][jj][w] * weights[w];
                v[ii * 7 + jj] = thopi * s;
            }
        return c6cache[key] = v;
    }
};
D4Model& d4model() { static D4Model m; return m; }

// ------------------------------------------------------------------------------------
// modified Broyden mixing, src/broyden.f90:37-150
// ------------------------------------------------------------------------------------
struct Broyden {
    int n = 0, maxiter = 0;
    dvec df, u, a, omega, qlast, dqlast;  // df,u: [it][n]
    void init(int n_, int maxiter_) {
        n = n_; maxiter = maxiter_;
        int keep = std::min(maxiter, 300);  // storage grown lazily below
        df.assign((size_t)keep * n, 0.0); u.assign((size_t)keep * n, 0.0);
        a.assign((size_t)maxiter * maxiter, 0.0); omega.assign(maxiter, 0.0);
        qlast.assign(n, 0.0); dqlast.assign(n, 0.0);
    }
    void step(dvec& q, const dvec& dq, int iter, double alpha) {
        const double omega0 = 0.01, minw = 1.0, maxw = 100000.0, wfac = 0.01;
        if (iter == 1) {
            dqlast = dq; qlast = q;
            for (int i = 0; i < n; ++i) q[i] += alpha * dq[i];
            return;
        }
        int it1 = iter - 1, c1 = it1 - 1;  // c1: 0-based slot
        if ((size_t)it1 * n > df.size()) { df.resize((size_t)it1 * 2 * n, 0.0); u.resize((size_t)it1 * 2 * n, 0.0); }
        double dd = 0;
        for (int i = 0; i < n; ++i) dd += dq[i] * dq[i];
        double om = std::sqrt(dd);
        if (om > (wfac / maxw)) om = wfac / om; else om = maxw;
        if (om < minw) om = minw;
        omega[c1] = om;
        double* dfc = &df[(size_t)c1 * n];
        double nn = 0;
        for (int i = 0; i < n; ++i) { dfc[i] = dq[i] - dqlast[i]; nn += dfc[i] * dfc[i]; }
        double inv = 1.0 / std::sqrt(nn);
        for (int i = 0; i < n; ++i) dfc[i] *= inv;
        dvec c(it1), beta((size_t)it1 * it1), gamma(it1);
        for (int i = 0; i < it1; ++i) {
            const double* dfi = &df[(size_t)i * n];
            double s = 0, t = 0;
            for (int k = 0; k < n; ++k) { s += dfi[k] * dfc[k]; t += dfi[k] * dq[k]; }
            a[i + (size_t)maxiter * c1] = s; a[c1 + (size_t)maxiter * i] = s;
            c[i] = omega[i] * t;
        }
        for (int i = 0; i < it1; ++i) {
            for (int j = 0; j < it1; ++j) beta[j + (size_t)it1 * i] = omega[j] * omega[i] * a[j + (size_t)maxiter * i];
            beta[i + (size_t)it1 * i] += omega0 * omega0;
        }
        // matinv: dgetrf + dgetri (broyden.f90:156-184)
        ivec ipiv(it1); int info = 0; dvec work(it1);
        scipy_dgetrf_(&it1, &it1, beta.data(), &it1, ipiv.data(), &info);
        if (info == 0) scipy_dgetri_(&it1, beta.data(), &it1, ipiv.data(), work.data(), &it1, &info);
        for (int j = 0; j < it1; ++j) { double s = 0; for (int i = 0; i < it1; ++i) s += c[i] * beta[i + (size_t)it1 * j]; gamma[j] = s; }
        double* uc = &u[(size_t)c1 * n];
        for (int i = 0; i < n; ++i) uc[i] = alpha * dfc[i] + inv * (q[i] - qlast[i]);
        dqlast = dq; qlast = q;
        for (int i = 0; i < n; ++i) q[i] += alpha * dq[i];
        for (int k = 0; k < it1; ++k) {
            const double* uk = &u[(size_t)k * n];
            double f = omega[k] * gamma[k];
            for (int i = 0; i < n; ++i) q[i] -= f * uk[i];
        }
    }
};

// occu, src/scc_core.f90:1112-1170
void occu(int ndim, int nel, int& nopen, int& ihomoa, int& ihomob, dvec& focca, dvec& foccb) {
    ivec focc(ndim, 0);
    focca.assign(ndim, 0.0); foccb.assign(ndim, 0.0);
    if (nel % 2 == 0) {
        int ihomo = nel / 2;
        for (int i = 0; i < ihomo; ++i) focc[i] = 2;
        if (nopen > 1)
            for (int i = 1; i <= nopen / 2; ++i) { focc[ihomo - i] -= 1; focc[ihomo + i - 1] += 1; }
    } else {
        int na = nel / 2 + (nopen - 1) / 2 + 1, nb = nel / 2 - (nopen - 1) / 2;
        for (int i = 0; i < na; ++i) focc[i] += 1;
        for (int i = 0; i < nb; ++i) focc[i] += 1;
    }
    ihomoa = ihomob = 0;
    for (int i = 0; i < ndim; ++i) {
        if (focc[i] == 2) { focca[i] = 1.0; foccb[i] = 1.0; }
        if (focc[i] == 1) focca[i] = 1.0;
    }
    for (int i = 0; i < ndim; ++i) {
        if (focca[i] > 0.99) ihomoa = i + 1;
        if (foccb[i] > 0.99) ihomob = i + 1;
    }
}
// fermismear, src/scc_core.f90:983-1064 (nel = number of electrons of this spin = HOMO index)
void fermismear(int norbs, int nel, double t, const double* eig, double* occ, double& e_fermi, double& s) {
    const double boltz = KB * AUTOEV, thr = 1e-9;
    double bkt = boltz * t;
    if (nel + 1 > norbs) {
        e_fermi = eig[nel - 1];
    } else if (nel == 0) {
        e_fermi = eig[0];
    } else {
        e_fermi = 0.5 * (eig[nel - 1] + eig[nel]);
        double occt = nel;
        for (int ncycle = 1; ncycle <= 200; ++ncycle) {
            double total_number = 0, total_dfermi = 0;
            for (int i = 0; i < norbs; ++i) {
                double fermifunct = 0, dfermifunct = 0;
                double x = (eig[i] - e_fermi) / bkt;
                if (x < 50) {
                    double ex = std::exp(x);
                    fermifunct = 1.0 / (ex + 1.0);
                    dfermifunct = ex / (bkt * (ex + 1.0) * (ex + 1.0));
                }
                occ[i] = fermifunct;
                total_number += fermifunct;
                total_dfermi += dfermifunct;
            }
            double change_fermi = (occt - total_number) / total_dfermi;  // unconditional, scc_core.f90:1041
            e_fermi += change_fermi;
            if (std::fabs(occt - total_number) <= thr) break;
        }
    }
    s = 0;
    for (int i = 0; i < norbs; ++i)
        if (occ[i] > thr && 1.0 - occ[i] > thr) s += occ[i] * std::log(occ[i]) + (1.0 - occ[i]) * std::log(1.0 - occ[i]);
    s = s * KB * t;
}
// dmat, src/scc_core.f90:1177-1205 :  P = C * diag(focc) * C^T
void dmat(int n, const double* focc, const double* C, double* P) {
    dvec T((size_t)n * n);
    for (int m = 0; m < n; ++m) for (int i = 0; i < n; ++i) T[i + (size_t)n * m] = C[i + (size_t)n * m] * focc[m];
    double one = 1.0, zero = 0.0;
    scipy_dgemm_("N", "T", &n, &n, &n, &one, C, &n, T.data(), &n, &zero, P, &n);
}

// ------------------------------------------------------------------------------------
// multipole potentials and energies, src/aespot.F90
// ------------------------------------------------------------------------------------
inline int lin0(int a, int b) { int mx = std::max(a, b), mn = std::min(a, b); return mn + mx * (mx + 1) / 2; }  // 0-based lin
// setvsdq (shift=true, aespot.F90:727-848) / setdvsdq (shift=false, aespot.F90:861-952)
void setvsdq(const Oracle& o, bool shift, const dvec& q, const dvec& dipm, const dvec& qp, const dvec& gab3,
             const dvec& gab5, dvec& vs, dvec& vd, dvec& vq) {
    int n = o.nat;
    vs.assign(n, 0.0); vd.assign(3 * n, 0.0); vq.assign(6 * n, 0.0);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) {
        const double* ra = R(o, i);
        double stmp = 0, dtmp[3] = {0, 0, 0}, qtmp[6] = {0, 0, 0, 0, 0, 0};
        for (int j = 0; j < n; ++j) {
            double g3 = gab3[j + (size_t)n * i], g5 = gab5[j + (size_t)n * i];
            const double* rb = R(o, j);
            double dra[3] = {ra[0] - rb[0], ra[1] - rb[1], ra[2] - rb[2]};
            double dum3a = 0, dum5a = 0, r2a = 0, r2ab = 0, t1a = 0, t2a = 0, t3a = 0;
            for (int l1 = 0; l1 < 3; ++l1) {
                if (shift) { r2a += ra[l1] * ra[l1]; t1a += ra[l1] * dra[l1]; t3a += ra[l1] * dipm[3 * j + l1]; }
                r2ab += dra[l1] * dra[l1];
                t2a += dipm[3 * j + l1] * dra[l1];
                for (int l2 = 0; l2 < 3; ++l2) {
                    int ll = lin0(l1, l2);
                    if (shift) dum5a = dum5a - qp[6 * j + ll] * dra[l1] * dra[l2] - 1.50 * q[j] * dra[l1] * dra[l2] * ra[l1] * ra[l2];
                    else dum5a = dum5a - qp[6 * j + ll] * dra[l1] * dra[l2];
                    if (l2 >= l1) continue;
                    int ki = l1 + l2 + 2;  // 0-based index of xy,xz,yz in (xx,yy,zz,xy,xz,yz)
                    qtmp[ki] -= 3.0 * q[j] * g5 * dra[l2] * dra[l1];
                }
                qtmp[l1] -= 1.50 * q[j] * g5 * dra[l1] * dra[l1];
            }
            if (shift) {
                dum3a = -t1a * q[j] - t2a;
                dum5a = dum5a + t3a * r2ab - 3.0 * t1a * t2a + 0.50 * q[j] * r2a * r2ab;
                stmp = stmp + dum5a * g5 + dum3a * g3;
            } else {
                dum3a = -t2a;
                stmp = stmp + dum3a * g3 + dum5a * g5;
            }
            for (int l1 = 0; l1 < 3; ++l1) {
                double d3 = dra[l1] * q[j], d5;
                if (shift) d5 = 3.0 * dra[l1] * t2a - r2ab * dipm[3 * j + l1] - q[j] * r2ab * ra[l1] + 3.0 * q[j] * dra[l1] * t1a;
                else d5 = 3.0 * dra[l1] * t2a - r2ab * dipm[3 * j + l1];
                dtmp[l1] += d3 * g3 + d5 * g5;
                qtmp[l1] += 0.50 * r2ab * q[j] * g5;
            }
        }
        vs[i] = stmp;
        for (int k = 0; k < 3; ++k) vd[3 * i + k] = dtmp[k];
        for (int k = 0; k < 6; ++k) vq[6 * i + k] = qtmp[k];
        double qs1 = gfn2_dip_kernel[o.Z[i] - 1] * 2.0, qs2 = gfn2_quad_kernel[o.Z[i] - 1] * 6.0;
        double t3a = 0, t2a = 0;
        for (int l1 = 0; l1 < 3; ++l1) {
            if (shift) t3a += ra[l1] * dipm[3 * i + l1] * qs1;
            vd[3 * i + l1] -= qs1 * dipm[3 * i + l1];
            for (int l2 = 0; l2 < l1; ++l2) {
                int ll = lin0(l1, l2), ki = l1 + l2 + 2;
                vq[6 * i + ki] -= qp[6 * i + ll] * qs2;
                if (shift) {
                    t3a -= ra[l1] * ra[l2] * qp[6 * i + ll] * qs2;
                    vd[3 * i + l1] += ra[l2] * qp[6 * i + ll] * qs2;
                    vd[3 * i + l2] += ra[l1] * qp[6 * i + ll] * qs2;
                }
            }
            int ll = lin0(l1, l1);
            vq[6 * i + l1] -= qp[6 * i + ll] * qs2 * 0.50;
            if (shift) {
                t3a -= ra[l1] * ra[l1] * qp[6 * i + ll] * qs2 * 0.50;
                vd[3 * i + l1] += ra[l1] * qp[6 * i + ll] * qs2;
            }
            t2a += qp[6 * i + ll];
        }
        if (shift) vs[i] += t3a;
        t2a *= gfn2_quad_kernel[o.Z[i] - 1];
        for (int l1 = 0; l1 < 3; ++l1) {
            vq[6 * i + l1] += t2a;
            if (shift) { vd[3 * i + l1] -= 2.0 * ra[l1] * t2a; vs[i] += t2a * ra[l1] * ra[l1]; }
        }
    }
}
// aniso_electro_cpu, aespot.F90:569-662
void aniso_electro(const Oracle& o, const dvec& q, const dvec& dipm, const dvec& qp, const dvec& gab3, const dvec& gab5,
                   double& e, double& epol) {
    static const int idx[3][3] = {{0, 1, 3}, {1, 2, 4}, {3, 4, 5}};
    int n = o.nat;
    epol = 0;
    for (int i = 0; i < n; ++i) {
        const double* dp1 = &dipm[3 * i];
        const double* qp1 = &qp[6 * i];
        double tt = dp1[0] * dp1[0] + dp1[1] * dp1[1] + dp1[2] * dp1[2], tt3 = 0;
        for (int k = 0; k < 3; ++k) for (int l = 0; l < 3; ++l) tt3 += qp1[idx[l][k]] * qp1[idx[l][k]];
        epol += gfn2_dip_kernel[o.Z[i] - 1] * tt + tt3 * gfn2_quad_kernel[o.Z[i] - 1];
    }
    double e01 = 0, e02 = 0, e11 = 0;
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : e01, e02, e11)
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < i; ++j) {
            double q1 = q[i];
            const double* dp1 = &dipm[3 * i];
            const double* qp1 = &qp[6 * i];
            double rij[3], r2 = 0;
            for (int k = 0; k < 3; ++k) { rij[k] = R(o, j)[k] - R(o, i)[k]; r2 += rij[k] * rij[k]; }
            double ed = 0, eq = 0, edd = 0;
            for (int k = 0; k < 3; ++k) {
                ed += q[j] * dp1[k] * rij[k];
                ed -= dipm[3 * j + k] * q1 * rij[k];
                for (int l = 0; l < 3; ++l) {
                    int kl = idx[l][k];
                    double tt = rij[l] * rij[k], tt3 = 3.0 * tt;
                    eq += q[j] * qp1[kl] * tt;
                    eq += qp[6 * j + kl] * q1 * tt;
                    edd -= dipm[3 * j + k] * dp1[l] * tt3;
                }
                edd += dipm[3 * j + k] * dp1[k] * r2;
            }
            e01 += ed * gab3[j + (size_t)n * i];
            e02 += eq * gab5[j + (size_t)n * i];
            e11 += edd * gab5[j + (size_t)n * i];
        }
    e = e01 + e02 + e11;
}
// dzero / gab / dgab, aespot.F90:1317-1357
inline void dzero(double dex, double rabinv, double radi, double radj, double& damp, double& ddamp) {
    double rco = 0.5 * (radi + radj);
    damp = 1.0 / (1.0 + 6.0 * std::pow(rco * rabinv, dex));
    ddamp = -dex * rabinv * (damp * damp - damp);
}

// mmompop_cpu, aespot.F90:289-427
void mmompop(const Oracle& o, const double* P, const double* S, const double* D, const double* Q, dvec& dipm, dvec& qp) {
    int n = o.bas.nao, nat = o.nat;
    dipm.assign(3 * nat, 0.0); qp.assign(6 * nat, 0.0);
    int nth = o.nthreads;
    std::vector<dvec> dl(nth, dvec(3 * nat, 0.0)), ql(nth, dvec(6 * nat, 0.0));
#pragma omp parallel num_threads(nth)
    {
#ifdef _OPENMP
        int tid = omp_get_thread_num();
#else
        int tid = 0;
#endif
        double* dm = dl[tid].data();
        double* qm = ql[tid].data();
#pragma omp for schedule(dynamic, 8)
        for (int i = 0; i < n; ++i) {
            int ii = o.bas.ao2at[i];
            const double* ra = R(o, ii);
            for (int j = 0; j <= i; ++j) {
                int jj = o.bas.ao2at[j];
                const double* rb = R(o, jj);
                size_t ji = j + (size_t)n * i;
                double pij = P[ji], ps = pij * S[ji];
                bool diag = (i == j);
                for (int k = 0; k < 3; ++k) {
                    double xk1 = ra[k], xk2 = rb[k];
                    double pdmk = pij * D[k + 3 * ji];
                    dm[3 * ii + k] += xk1 * ps - pdmk;
                    if (!diag) dm[3 * jj + k] += xk2 * ps - pdmk;
                    for (int l = 0; l < k; ++l) {
                        int kl = k * (k + 1) / 2 + l;  // 0-based packed (xx,xy,yy,xz,yz,zz)
                        int kj = k + l + 2;            // 0-based in (xx,yy,zz,xy,xz,yz)
                        double xl1 = ra[l], xl2 = rb[l];
                        double pdml = pij * D[l + 3 * ji];
                        double pqm = pij * Q[kj + 6 * ji];
                        qm[6 * ii + kl] += pdmk * xl1 + pdml * xk1 - xl1 * xk1 * ps - pqm;
                        if (!diag) qm[6 * jj + kl] += pdmk * xl2 + pdml * xk2 - xl2 * xk2 * ps - pqm;
                    }
                    int kl = k * (k + 1) / 2 + k;
                    double pqm = pij * Q[k + 6 * ji];
                    qm[6 * ii + kl] += 2.0 * pdmk * xk1 - xk1 * xk1 * ps - pqm;
                    if (!diag) qm[6 * jj + kl] += 2.0 * pdmk * xk2 - xk2 * xk2 * ps - pqm;
                }
            }
        }
    }
    for (int t = 0; t < nth; ++t) {
        for (int k = 0; k < 3 * nat; ++k) dipm[k] += dl[t][k];
        for (int k = 0; k < 6 * nat; ++k) qp[k] += ql[t][k];
    }
    for (int i = 0; i < nat; ++i) {
        double tii = 0.50 * (qp[6 * i + 0] + qp[6 * i + 2] + qp[6 * i + 5]);
        for (int k = 0; k < 6; ++k) qp[6 * i + k] *= 1.50;
        qp[6 * i + 0] -= tii; qp[6 * i + 2] -= tii; qp[6 * i + 5] -= tii;
    }
}

// ------------------------------------------------------------------------------------
// the single point, src/scf_module.F90:74-920 (scf) + src/scc_core.f90:255-640 (scc)
// ------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------
// implicit solvation: GBSA / ALPB (src/solv/{model,gbsa,born,sasa,kernel,state}.f90)
// ------------------------------------------------------------------------------------
struct Solv {
    int n = 0, kernel = 2;
    bool lsalt = false, lhb = false;
    double keps = 0, alpbet = 0, gshift = 0, born_scale = 1, eps = 1, kappa = 0, ion_rad = 0, lrcut = 0, srcut = 0;
    dvec vdwr, rho, svdw, vdwsa, wrp, trj2, gamsasa, hbmag;
    // geometry dependent
    dvec brad, brdr /* [i][j][3]: d brad(i) / d R_j */, sasa, dsdrt /* [i][j][3]: d sasa(i) / d R_j */, dsdr, hbw, dhbdw;
    dvec ionscr, discr, bornmat;
    double gsasa = 0, adet = 0;
};

const double KCALTOAU = 1.0 / 627.50947428;   // src/mctc/convert.f90:31-32

// initSolvModel + loadInternalParam + paramToModel + initBorn: src/solv/model.f90:185-207,250-474,498-528,
// src/solv/gbsa.f90:197-388.  Returns false when the solvent has no internal GFN2 parameter set for the kernel.
bool solv_init(const Oracle& o, Solv& sv) {
    const SolvParamSet* ps = nullptr;
    for (int i = 0; i < solv_nsets; ++i)
        if (solv_param_sets[i].kernel == o.solv_kernel && o.solv_name == solv_param_sets[i].name) ps = &solv_param_sets[i];
    if (!ps) return false;
    const int n = o.nat;
    sv.n = n; sv.kernel = o.solv_kernel;
    const double molcm3toau = 8.92388e-2, surfaceTension = 1.0e-5;
    const double eps = ps->epsv, density = ps->rhos / ps->smass * molcm3toau;
    double gshift = ps->gshift * KCALTOAU;
    const double born_offset = ps->soset * 0.1 * AATOAU, probe = ps->rprobe * AATOAU;
    // getStateShift, src/solv/state.f90:55-85
    const double ideal = 24.79, ambient = 298.15;
    if (o.solv_state == 2)
        gshift += o.solv_temp * KB * (std::log(ideal * o.solv_temp / ambient) + std::log(density / 1.0e-3 * 1.0 / ps->smass));
    else if (o.solv_state == 3)
        gshift += o.solv_temp * KB * std::log(ideal * o.solv_temp / ambient);
    bool anyhb = false;
    for (int k = 0; k < 94; ++k) anyhb = anyhb || std::fabs(ps->tmp[k]) > 1.e-3;
    sv.gshift = gshift; sv.eps = eps; sv.born_scale = ps->c1;
    sv.alpbet = o.solv_alpb ? 0.571412 / eps : 0.0;
    sv.keps = (1.0 / eps - 1.0) / (1.0 + sv.alpbet);
    sv.lsalt = o.solv_ionst > 0.0;
    if (sv.lsalt) {
        const double kappaConst = 0.7897e-3;
        sv.ion_rad = o.solv_ionrad;
        sv.kappa = 1.0 / (std::sqrt(eps * o.solv_temp * kappaConst / o.solv_ionst) * AATOAU);
    }
    sv.lrcut = 35.0 * AATOAU;
    const double w = 0.3 * AATOAU, w3 = w * (w * w), ah3 = -1.0 / (4.0 * w3);
    sv.vdwr.resize(n); sv.rho.resize(n); sv.svdw.resize(n); sv.vdwsa.resize(n); sv.wrp.resize(n); sv.trj2.resize(2 * n);
    sv.gamsasa.resize(n); sv.hbmag.resize(n);
    double maxsa = 0;
    for (int i = 0; i < n; ++i) {
        const int z = o.Z[i] - 1;
        sv.vdwr[i] = AATOAU * solv_vdw_rad_d3_aa[z];
        sv.rho[i] = sv.vdwr[i] * ps->sx[z];
        sv.svdw[i] = sv.vdwr[i] - born_offset;
        sv.vdwsa[i] = sv.vdwr[i] + probe;
        maxsa = std::max(maxsa, sv.vdwsa[i]);
        sv.trj2[2 * i] = (sv.vdwsa[i] - w) * (sv.vdwsa[i] - w);
        sv.trj2[2 * i + 1] = (sv.vdwsa[i] + w) * (sv.vdwsa[i] + w);
        const double v = sv.vdwsa[i];
        double r = v + w;
        double wr = (0.25 / w + 3.0 * ah3 * (0.2 * r * r - 0.5 * r * v + v * v / 3.0)) * r * r * r;
        r = v - w;
        wr -= (0.25 / w + 3.0 * ah3 * (0.2 * r * r - 0.5 * r * v + v * v / 3.0)) * r * r * r;
        sv.wrp[i] = wr;
        sv.gamsasa[i] = ps->gamscale[z] * (4.0 * PI) * surfaceTension;
        sv.hbmag[i] = anyhb ? -(ps->tmp[z] * ps->tmp[z]) * KCALTOAU : 0.0;
        sv.lhb = sv.lhb || sv.hbmag[i] < 0.0;
    }
    sv.srcut = 2.0 * (w + maxsa) + 2.0 * AATOAU;
    return true;
}

// one descreening term psi_i <- j and its radial derivative factor (d psi / d r) / r: src/solv/born.f90:130-330.
// r1: the value the reference's variable r1 holds when the non-overlap formula is evaluated (1/r of this pair except in
// the case described at solv_update)
inline bool born_psi_term(double r, double r1_nonoverlap, double rho_j, double vdw_i, double& g, double& dg) {
    double r1 = 1.0 / r;
    if (r >= vdw_i + rho_j) {   // no overlap
        r1 = r1_nonoverlap;
        const double ap = r + rho_j, am = r - rho_j, ab = ap * am, rhab = rho_j / ab, lnab = 0.5 * std::log(am / ap) * r1;
        g = rhab + lnab;
        dg = -2.0 * rhab / ab + (rhab - lnab) * r1 * r1;
        return true;
    }
    if (r + rho_j > vdw_i) {    // partial overlap
        const double r12 = 0.5 * r1, r24 = r12 * r12, ap = r + rho_j, am = r - rho_j, rh1 = 1.0 / vdw_i, rhr1 = 1.0 / ap,
                     aprh1 = ap * rh1, lnab = std::log(aprh1);
        g = rh1 - rhr1 + r12 * (0.5 * am * (rhr1 - rh1 * aprh1) - lnab);
        dg = rhr1 * rhr1 * (1.0 - 0.25 * am * r1 * (1.0 + aprh1 * aprh1)) + rho_j * r24 * (rhr1 - rh1 * aprh1) + r12 * (r1 * lnab - rhr1);
        dg = dg * r1;
        return true;
    }
    return false;               // sphere j inside the van der Waals sphere of i
}

// update: src/solv/gbsa.f90:391-432 (neighbour lists 826-1010 are implied by the cutoffs lrcut / srcut)
void solv_update(const Oracle& o, Solv& sv) {
    const int n = sv.n;
    const size_t nn = (size_t)n * n;
    // ---- compute_psi / compute_bornr (GBOBC II): src/solv/born.f90:36-84,87-345
    // The reference keeps 1/r in a variable r1 that every branch of its pair loop sets before use, except one: when the
    // descreening sphere of j lies completely inside the van der Waals sphere of i (r + rho_j <= vdw_i, nothing to add to
    // psi_i) while i does not overlap j, the j <- i term (born.f90:292-309, `case(1)`) is evaluated with the r1 left behind
    // by the preceding pair of the list.  Reproduced here: r1_state is that variable, the pair order is the list's
    // (i ascending, j < i ascending, gbsa.f90:883-906); before the first assignment the reference's value is undefined, 0 here.
    sv.brad.assign(n, 0.0); sv.brdr.assign(3 * nn, 0.0);
    const double lr2 = sv.lrcut * sv.lrcut;
    double r1_state = 0.0;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < i; ++j) {
            double dr[3], r2 = 0;
            for (int k = 0; k < 3; ++k) { dr[k] = R(o, i)[k] - R(o, j)[k]; r2 += dr[k] * dr[k]; }
            if (!(r2 < lr2)) continue;
            const double r = std::sqrt(r2);
            const bool ovij = r < sv.vdwr[i] + sv.rho[j], ovji = r < sv.rho[i] + sv.vdwr[j];
            const bool ij_term = !ovij || r + sv.rho[j] > sv.vdwr[i], ji_term = !ovji || r + sv.rho[i] > sv.vdwr[j];
            // which branches assign r1: every one that evaluates a term, except the j <- i term of case(1)
            const bool stale = ovij && !ovji && !ij_term;
            if (stale) cnt_add(11, 1);
            else if (ij_term || ji_term) r1_state = 1.0 / r;
            double g, dg;
            if (born_psi_term(r, 1.0 / r, sv.rho[j], sv.vdwr[i], g, dg)) {
                sv.brad[i] += g;
                for (int k = 0; k < 3; ++k) { sv.brdr[3 * ((size_t)i * n + i) + k] += dg * dr[k]; sv.brdr[3 * ((size_t)i * n + j) + k] -= dg * dr[k]; }
            }
            if (born_psi_term(r, stale ? r1_state : 1.0 / r, sv.rho[i], sv.vdwr[j], g, dg)) {
                sv.brad[j] += g;
                for (int k = 0; k < 3; ++k) { sv.brdr[3 * ((size_t)j * n + j) + k] -= dg * dr[k]; sv.brdr[3 * ((size_t)j * n + i) + k] += dg * dr[k]; }
            }
        }
    const double alp = 1.0, bet = 0.8, gam = 4.85;   // src/solv/born.f90:28-30
    for (int i = 0; i < n; ++i) {
        double br = sv.brad[i];
        const double s1 = 1.0 / sv.svdw[i], v1 = 1.0 / sv.vdwr[i], s2 = 0.5 * sv.svdw[i];
        br = br * s2;
        double arg2 = br * (gam * br - bet);
        const double arg = br * (alp + arg2);
        arg2 = 2.0 * arg2 + alp + gam * br * br;
        const double th = std::tanh(arg), ch = std::cosh(arg);
        br = sv.born_scale / (s1 - v1 * th);
        double dpsi = ch * (s1 - v1 * th);
        dpsi = sv.born_scale * s2 * v1 * arg2 / (dpsi * dpsi);
        sv.brad[i] = br;
        for (size_t t = 3 * (size_t)i * n; t < 3 * (size_t)(i + 1) * n; ++t) sv.brdr[t] *= dpsi;
    }
    // ---- compute_numsa / compute_w_sp: src/solv/sasa.f90:41-125,128-204
    const double w = 0.3 * AATOAU, w3 = w * (w * w), ah0 = 0.5, ah1 = 3.0 / (4.0 * w), ah3 = -1.0 / (4.0 * w3), tolsesp = 1.e-6;
    sv.sasa.assign(n, 0.0); sv.dsdrt.assign(3 * nn, 0.0);
    const double sr2 = sv.srcut * sv.srcut;
#pragma omp parallel for schedule(dynamic, 4) num_threads(o.nthreads)
    for (int i = 0; i < n; ++i) {
        ivec nb;
        for (int j = 0; j < n; ++j) {
            if (j == i) continue;
            double r2 = 0;
            for (int k = 0; k < 3; ++k) { const double t = R(o, i)[k] - R(o, j)[k]; r2 += t * t; }
            if (r2 < sr2) nb.push_back(j);
        }
        dvec grds(3 * nb.size());
        ivec grdi(nb.size());
        double sasai = 0;
        double* gi = &sv.dsdrt[3 * (size_t)i * n];
        for (int ip = 0; ip < solv_nang; ++ip) {
            double xyzp[3];
            for (int k = 0; k < 3; ++k) xyzp[k] = R(o, i)[k] + sv.vdwsa[i] * solv_leb230[ip][k];
            int nni = 0;
            double sasap = 1.0;
            for (size_t t = 0; t < nb.size(); ++t) {
                const int ia = nb[t];
                double tj[3], tj2 = 0;
                for (int k = 0; k < 3; ++k) { tj[k] = xyzp[k] - R(o, ia)[k]; tj2 += tj[k] * tj[k]; }
                if (tj2 < sv.trj2[2 * ia + 1]) {
                    if (tj2 <= sv.trj2[2 * ia]) { sasap = 0.0; break; }
                    const double sqtj = std::sqrt(tj2), uj = sqtj - sv.vdwsa[ia], ah3uj2 = ah3 * uj * uj;
                    double dsasaij = ah1 + 3.0 * ah3uj2, sasaij = ah0 + (ah1 + ah3uj2) * uj;
                    sasaij = std::max(0.0, std::min(sasaij, 1.0));
                    if (sasaij < 2.2250738585072014e-308) { sasap = 0.0; break; }
                    sasap *= sasaij;
                    dsasaij = dsasaij / (sasaij * sqtj);
                    grdi[nni] = ia;
                    for (int k = 0; k < 3; ++k) grds[3 * nni + k] = dsasaij * tj[k];
                    nni++;
                }
            }
            if (sasap > tolsesp) {
                const double wsa = solv_leb230[ip][3] * sv.wrp[i] * sasap;
                sasai += wsa;
                for (int jj = 0; jj < nni; ++jj)
                    for (int k = 0; k < 3; ++k) {
                        const double dd = wsa * grds[3 * jj + k];
                        gi[3 * i + k] += dd;
                        gi[3 * grdi[jj] + k] -= dd;
                    }
            }
        }
        sv.sasa[i] = sasai;
    }
    // dsdr = dsdrt . gamsasa, gsasa (gbsa.f90:419-420)
    sv.dsdr.assign(3 * n, 0.0);
    sv.gsasa = 0;
    for (int i = 0; i < n; ++i) {
        sv.gsasa += sv.sasa[i] * sv.gamsasa[i];
        for (int j = 0; j < n; ++j) for (int k = 0; k < 3; ++k) sv.dsdr[3 * j + k] += sv.dsdrt[3 * ((size_t)i * n + j) + k] * sv.gamsasa[i];
    }
    // getDebyeHueckel, gbsa.f90:794-821
    if (sv.lsalt) {
        sv.ionscr.resize(n); sv.discr.resize(n);
        const double aa = 0.5 / sv.eps;
        for (int i = 0; i < n; ++i) {
            const double gg = sv.kappa * (sv.brad[i] + sv.ion_rad);
            sv.ionscr[i] = aa * std::exp(gg) / (1.0 + gg);
            sv.discr[i] = sv.ionscr[i] * sv.kappa * gg / (1.0 + gg);
        }
    }
    // compute_fhb, gbsa.f90:1013-1046
    sv.hbw.assign(n, 0.0); sv.dhbdw.assign(n, 0.0);
    if (sv.lhb)
        for (int i = 0; i < n; ++i) {
            const double smaxd = 1.0 / (sv.vdwsa[i] * sv.vdwsa[i]);
            sv.hbw[i] = sv.hbmag[i] * sv.sasa[i] * smaxd;
            sv.dhbdw[i] = sv.hbmag[i] * smaxd;
        }
}

// inertia tensor of the van der Waals spheres and the shape descriptor of ALPB: getADet, gbsa.f90:643-687
void solv_inertia(const Oracle& o, const Solv& sv, double center[3], double inertia[3][3], double& totRad3, double& adet) {
    const double tof = 2.0 / 5.0;
    totRad3 = 0;
    for (int k = 0; k < 3; ++k) center[k] = 0;
    for (int i = 0; i < sv.n; ++i) {
        const double rad3 = sv.vdwr[i] * sv.vdwr[i] * sv.vdwr[i];
        totRad3 += rad3;
        for (int k = 0; k < 3; ++k) center[k] += R(o, i)[k] * rad3;
    }
    for (int k = 0; k < 3; ++k) center[k] /= totRad3;
    for (int a = 0; a < 3; ++a) for (int c = 0; c < 3; ++c) inertia[a][c] = 0;
    for (int i = 0; i < sv.n; ++i) {
        const double rad2 = sv.vdwr[i] * sv.vdwr[i], rad3 = rad2 * sv.vdwr[i];
        double vec[3], r2 = 0;
        for (int k = 0; k < 3; ++k) { vec[k] = R(o, i)[k] - center[k]; r2 += vec[k] * vec[k]; }
        for (int a = 0; a < 3; ++a) {
            inertia[a][a] += rad3 * (r2 + tof * rad2);
            for (int c = 0; c < 3; ++c) inertia[a][c] -= rad3 * vec[a] * vec[c];
        }
    }
    const double det = inertia[0][0] * (inertia[1][1] * inertia[2][2] - inertia[1][2] * inertia[2][1]) -
                       inertia[0][1] * (inertia[1][0] * inertia[2][2] - inertia[1][2] * inertia[2][0]) +
                       inertia[0][2] * (inertia[1][0] * inertia[2][1] - inertia[1][1] * inertia[2][0]);
    adet = std::sqrt(std::pow(det, 1.0 / 3.0) / (tof * totRad3));
}

// addBornMatrix: gbsa.f90:435-479 with the kernels of src/solv/kernel.f90:620-760
void solv_born_matrix(const Oracle& o, Solv& sv) {
    const int n = sv.n;
    sv.bornmat.assign((size_t)n * n, 0.0);
    const double zetaP16o16 = 1.028 / 16.0;
    for (int i = 0; i < n; ++i) {
        for (int j = 0; j < i; ++j) {
            double r2 = 0;
            for (int k = 0; k < 3; ++k) { const double t = R(o, i)[k] - R(o, j)[k]; r2 += t * t; }
            double v;
            if (sv.kernel == 1) {
                const double aa = sv.brad[i] * sv.brad[j], dd = 0.25 * r2 / aa, expd = std::exp(-dd), fgb2 = r2 + aa * expd;
                if (sv.lsalt) {
                    const double fgb = std::sqrt(fgb2), gg = sv.ionscr[i] + sv.ionscr[j];
                    v = (std::exp(-sv.kappa * fgb) * gg - 1.0) / fgb;
                } else v = sv.keps / std::sqrt(fgb2);
            } else {
                const double r1 = std::sqrt(r2), ab = std::sqrt(sv.brad[i] * sv.brad[j]);
                double arg = ab / (ab + zetaP16o16 * r1);
                arg = arg * arg; arg = arg * arg; arg = arg * arg; arg = arg * arg;
                v = sv.keps / (r1 + ab * arg);
            }
            sv.bornmat[i + (size_t)n * j] += v; sv.bornmat[j + (size_t)n * i] += v;
        }
        if (sv.kernel == 1 && sv.lsalt) sv.bornmat[i + (size_t)n * i] += (std::exp(-sv.kappa * sv.brad[i]) * sv.ionscr[i] * 2.0 - 1.0) / sv.brad[i];
        else sv.bornmat[i + (size_t)n * i] += sv.keps / sv.brad[i];
        if (sv.lhb) sv.bornmat[i + (size_t)n * i] += 2.0 * sv.hbw[i];
    }
    if (sv.alpbet > 0.0) {
        double c[3], in[3][3], tr;
        solv_inertia(o, sv, c, in, tr, sv.adet);
        for (auto& v : sv.bornmat) v += sv.keps * sv.alpbet / sv.adet;
    }
}

// addGradient: gbsa.f90:573-640 (kernels: src/solv/kernel.f90:47-330; hydrogen bonds gbsa.f90:1049-1090; ALPB shape
// term gbsa.f90:690-791).  Returns gborn (incl. the ALPB term) and ghb.
void solv_gradient(const Oracle& o, const Solv& sv, const dvec& q, dvec& grad, double& gborn, double& ghb) {
    const int n = sv.n;
    dvec grddb(n, 0.0);
    double egb = 0;
    const double zetaP16 = 1.028, zetaP16o16 = zetaP16 / 16.0;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < i; ++j) {
            double dr[3], r2 = 0;
            for (int k = 0; k < 3; ++k) { dr[k] = R(o, i)[k] - R(o, j)[k]; r2 += dr[k] * dr[k]; }
            const double qq = q[i] * q[j];
            if (sv.kernel == 1 && sv.lsalt) {
                double aa = sv.brad[i] * sv.brad[j];
                const double dd = 0.25 * r2 / aa, expd = std::exp(-dd), dfgb = r2 + aa * expd, fgb = std::sqrt(dfgb);
                aa = sv.kappa * fgb;
                const double expa = std::exp(-aa), gg = (sv.ionscr[i] + sv.ionscr[j]) * expa;
                double qfg = qq / fgb;
                egb += qfg * (gg - 1.0);
                const double dfgb3 = qfg * (gg * (1.0 + aa) - 1.0) / dfgb, ap = (1.0 - 0.25 * expd) * dfgb3;
                for (int k = 0; k < 3; ++k) { grad[3 * i + k] -= ap * dr[k]; grad[3 * j + k] += ap * dr[k]; }
                qfg = qfg * expa;
                const double bp = -0.5 * expd * (1.0 + dd) * dfgb3;
                grddb[i] += sv.brad[j] * bp + qfg * sv.discr[i];
                grddb[j] += sv.brad[i] * bp + qfg * sv.discr[j];
            } else if (sv.kernel == 1) {
                const double aa = sv.brad[i] * sv.brad[j], dd = 0.25 * r2 / aa, expd = std::exp(-dd), fgb2 = r2 + aa * expd,
                             dfgb2 = 1.0 / fgb2, dfgb = std::sqrt(dfgb2), dfgb3 = dfgb2 * dfgb * sv.keps;
                egb += qq * sv.keps * dfgb;
                const double ap = (1.0 - 0.25 * expd) * dfgb3;
                for (int k = 0; k < 3; ++k) { grad[3 * i + k] -= ap * dr[k] * qq; grad[3 * j + k] += ap * dr[k] * qq; }
                const double bp = -0.5 * expd * (1.0 + dd) * dfgb3;
                grddb[i] += sv.brad[j] * bp * qq;
                grddb[j] += sv.brad[i] * bp * qq;
            } else {
                const double r1 = std::sqrt(r2), ab = std::sqrt(sv.brad[i] * sv.brad[j]), arg1 = ab / (ab + zetaP16o16 * r1);
                double arg16 = arg1 * arg1;
                arg16 = arg16 * arg16; arg16 = arg16 * arg16; arg16 = arg16 * arg16;
                const double fgb = r1 + ab * arg16, dfgb = 1.0 / fgb, dfgb2 = dfgb * dfgb;
                egb += qq * sv.keps * dfgb;
                const double ap = (1.0 - zetaP16 * arg1 * arg16) * dfgb2;
                for (int k = 0; k < 3; ++k) {
                    const double dG = ap * dr[k] * sv.keps / r1 * qq;
                    grad[3 * i + k] -= dG; grad[3 * j + k] += dG;
                }
                const double bp = -0.5 * (r1 * zetaP16 / ab * arg1 + 1.0) / ab * arg16 * dfgb2;
                grddb[i] += sv.brad[j] * bp * sv.keps * qq;
                grddb[j] += sv.brad[i] * bp * sv.keps * qq;
            }
        }
    for (int i = 0; i < n; ++i) {   // self-energy part
        if (sv.kernel == 1 && sv.lsalt) {
            const double gg = std::exp(-sv.kappa * sv.brad[i]), aa = 2.0 * sv.ionscr[i] * gg - 1.0, qq = q[i] / sv.brad[i];
            egb += 0.5 * qq * q[i] * aa;
            const double ap = aa - sv.brad[i] * 2.0 * (sv.discr[i] + sv.ionscr[i] * sv.kappa) * gg;
            grddb[i] += -0.5 * qq * qq * ap;
        } else {
            const double bp = 1.0 / sv.brad[i], qq = q[i] * bp;
            egb += 0.5 * q[i] * qq * sv.keps;
            grddb[i] += -0.5 * sv.keps * qq * bp * q[i];
        }
    }
    // contract with the Born radii derivatives: gradient += brdr . grddb
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j)
            for (int k = 0; k < 3; ++k) grad[3 * j + k] += sv.brdr[3 * ((size_t)i * n + j) + k] * grddb[i];
    gborn = egb;
    for (int k = 0; k < 3 * n; ++k) grad[k] += sv.dsdr[k];
    ghb = 0;
    if (sv.lhb) {
        for (int i = 0; i < n; ++i) ghb += sv.hbw[i] * q[i] * q[i];
        for (int i = 0; i < n; ++i) {
            if (sv.dhbdw[i] == 0.0) continue;
            const double dhbed = sv.dhbdw[i] * q[i] * q[i];
            for (int j = 0; j < n; ++j) for (int k = 0; k < 3; ++k) grad[3 * j + k] += sv.dsdrt[3 * ((size_t)i * n + j) + k] * dhbed;
        }
    }
    if (sv.alpbet > 0.0) {
        double qtot = 0;
        for (int i = 0; i < n; ++i) qtot += q[i];
        gborn += qtot * qtot * sv.alpbet / sv.adet * sv.keps;
        // addADetDeriv, gbsa.f90:690-791
        double c[3], in[3][3], tr, adet;
        solv_inertia(o, sv, c, in, tr, adet);
        const double kE = sv.keps * sv.alpbet;
        const double f = (250.0 / (48.0 * tr * tr * tr * std::pow(adet, 5))) * (-0.5 * kE * qtot * qtot / (adet * adet));
        double ad[3][3];
        ad[0][0] = in[0][0] * (in[1][1] + in[2][2]) - in[0][1] * in[0][1] - in[0][2] * in[0][2];
        ad[0][1] = ad[1][0] = in[0][1] * in[2][2] - in[0][2] * in[1][2];
        ad[0][2] = ad[2][0] = in[0][2] * in[1][1] - in[0][1] * in[2][1];
        ad[1][1] = in[1][1] * (in[0][0] + in[2][2]) - in[0][1] * in[0][1] - in[1][2] * in[1][2];
        ad[1][2] = ad[2][1] = in[0][0] * in[1][2] - in[0][1] * in[0][2];
        ad[2][2] = in[2][2] * (in[0][0] + in[1][1]) - in[0][2] * in[0][2] - in[1][2] * in[1][2];
        for (int i = 0; i < n; ++i) {
            const double rad3 = sv.vdwr[i] * sv.vdwr[i] * sv.vdwr[i];
            double vec[3];
            for (int k = 0; k < 3; ++k) vec[k] = R(o, i)[k] - c[k];
            for (int a = 0; a < 3; ++a) {
                double s = 0;
                for (int k = 0; k < 3; ++k) s += ad[a][k] * f * vec[k];
                grad[3 * i + a] += rad3 * s;
            }
        }
    }
}

int singlepoint(Oracle& o) {
    const int n = o.nat;
    build_basis(n, o.Z.data(), o.bas);
    const Basis& b = o.bas;
    const int nao = b.nao, nsh = b.nshell;
    auto& D = o.d;
    std::memset(o.timers, 0, sizeof(o.timers));
    std::memset(g_cnt, 0, sizeof(g_cnt));
    double t0 = now_s();

    // ---- guess (src/api/interface.f90:82-101, src/xtb/calculator.f90:432-496)
    dvec q(n, 0.0), qsh, dipm(3 * n, 0.0), qp(6 * n, 0.0);
    if (o.guess == 2 && (int)D["restart_qsh"].size() == nsh) {
        qsh = D["restart_qsh"]; dipm = D["restart_dipm"]; qp = D["restart_qp"];
    } else {
        if (o.guess == 0) {
            if (!eeq_charges(o, q)) q.assign(n, o.charge / n);
        } else {
            q.assign(n, o.charge / n);
        }
        D["q_guess"] = q;
        iniqshell(o, q, qsh);
    }
    o.timers[0] = now_s() - t0;  // guess (not a reference slot)
    t0 = now_s();

    // ---- setup (scf_module.F90:237-491)
    double intcut = std::max(20.0, 25.0 - 10.0 * std::log10(o.acc));
    double neglect = 10.0e-9 * o.acc;
    double scfconv = 1.e-6 * o.acc, qconv = 1.e-4 * o.acc;
    bool broy = n > 1;
    // qsh2qat
    for (int i = 0; i < n; ++i) q[i] = 0;
    for (int s = 0; s < nsh; ++s) q[b.sh[s].atom] += qsh[s];
    double sumz = 0;
    for (int i = 0; i < n; ++i) sumz += el_valence_z(o.Z[i]);
    int nel = (int)std::lround(sumz - o.charge);
    int nopen = o.uhf;
    if (nopen == 0 && nel % 2 != 0) nopen = 1;
    // setzshell, scc_core.f90:1509-1542
    dvec zsh(nsh);
    {
        int k = 0;
        for (int i = 0; i < n; ++i) {
            int z = o.Z[i];
            double ntot = -1.0e-6;
            for (int m = 0; m < el_nshell(z); ++m) {
                zsh[k] = gfn2_reference_occ_l[z - 1][el_ang(z, m)];
                ntot += zsh[k];
                if (ntot > el_valence_z(z)) zsh[k] = 0.0;
                k++;
            }
        }
    }
    if (nel > 2 * nao) return 3;
    dvec focca, foccb, focc(nao, 0.0);
    int ihomoa = 0, ihomob = 0, ihomo = 0;
    if (nel != 0) {
        occu(nao, nel, nopen, ihomoa, ihomob, focca, foccb);
        for (int i = 0; i < nao; ++i) focc[i] = focca[i] + foccb[i];
        ihomo = ihomoa;
    } else { focca.assign(nao, 0.0); foccb.assign(nao, 0.0); nopen = 0; }

    // Coulomb matrix, src/coulomb/klopmanohno.F90:257-312 with arithmetic average, gExp = 2
    dvec jmat((size_t)nsh * nsh, 0.0), gam3(nsh), hard(nsh);
    for (int s = 0; s < nsh; ++s) {
        int z = o.Z[b.sh[s].atom];
        hard[s] = el_shell_hardness(z, b.sh[s].ish_el);
        gam3[s] = el_third_order_shell(z, b.sh[s].ish_el);
    }
    const double gExp = 2.0;
#pragma omp parallel for schedule(dynamic, 8)
    for (int iat = 0; iat < n; ++iat) {
        for (int jat = 0; jat < iat; ++jat) {
            double r2 = 0;
            for (int k = 0; k < 3; ++k) { double t = R(o, jat)[k] - R(o, iat)[k]; r2 += t * t; }
            double r1 = std::sqrt(r2);
            for (int is = b.sh0[iat]; is < b.sh0[iat] + b.nsh[iat]; ++is)
                for (int js = b.sh0[jat]; js < b.sh0[jat] + b.nsh[jat]; ++js) {
                    double gij = 0.5 * (hard[is] + hard[js]);
                    double rterm = 1.0 / std::pow(std::pow(r1, gExp) + std::pow(gij, -gExp), 1.0 / gExp);
                    jmat[js + (size_t)nsh * is] = rterm; jmat[is + (size_t)nsh * js] = rterm;
                }
        }
        for (int is = b.sh0[iat]; is < b.sh0[iat] + b.nsh[iat]; ++is) {
            for (int js = b.sh0[iat]; js < is; ++js) {
                double gij = 0.5 * (hard[is] + hard[js]);
                jmat[js + (size_t)nsh * is] = gij; jmat[is + (size_t)nsh * js] = gij;
            }
            jmat[is + (size_t)nsh * is] = hard[is];
        }
    }
    o.timers[1] = now_s() - t0;
    t0 = now_s();

    // ---- classical repulsion, src/xtb/repulsion.F90:41-158
    dvec grad(3 * n, 0.0);
    double ep = 0;
    {
        dvec en(n, 0.0);
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < i; ++j) {
                int zi = o.Z[i], zj = o.Z[j];
                double alpha = std::sqrt(gfn2_rep_alpha[zi - 1] * gfn2_rep_alpha[zj - 1]);
                double zeff = gfn2_rep_zeff[zi - 1] * gfn2_rep_zeff[zj - 1];
                double kExp = (zi > 2 || zj > 2) ? 1.5 : 1.0;
                double rij[3], r2 = 0;
                for (int k = 0; k < 3; ++k) { rij[k] = R(o, i)[k] - R(o, j)[k]; r2 += rij[k] * rij[k]; }
                if (r2 > 1600.0 || r2 < 1.0e-8) continue;
                cnt_add(6, 1);
                double r1 = std::sqrt(r2);
                double t16 = std::pow(r1, kExp), t26 = std::exp(-alpha * t16), t27 = std::pow(r1, 1.0);
                double dE = zeff * t26 / t27;
                for (int k = 0; k < 3; ++k) {
                    double dG = -(alpha * t16 * kExp + 1.0) * dE * rij[k] / r2;
                    grad[3 * i + k] += dG; grad[3 * j + k] -= dG;
                }
                en[i] += 0.5 * dE; en[j] += 0.5 * dE;
            }
        for (int i = 0; i < n; ++i) ep += en[i];
    }
    D["grad_rep"] = grad;
    o.timers[3] = now_s() - t0;
    t0 = now_s();

    // ---- CN (gfn) and self-consistent D4 setup (scf_module.F90:518-537, src/xtb/dispersion.f90:69-107)
    dvec cn;
    std::vector<CNPair> cnpairs;
    cn_lp(o, true, cn, cnpairs);
    D["cn"] = cn;
    D4Model& dm = d4model();
    ivec itbl(n);
    int ndim = 0;
    for (int i = 0; i < n; ++i) { itbl[i] = ndim; ndim += d4_refn[o.Z[i] - 1]; dm.ensure(o.Z[i]); }
    dvec gw(ndim, 0.0), dispmat((size_t)ndim * ndim, 0.0);
    {
        dvec covcn;
        ncoord_erf_d4(o, true, covcn);
        D["cn_d4"] = covcn;
        const double wf = 6.0;
        // d4(), src/disp/dftd4.F90:588-657
        for (int i = 0; i < n; ++i) {
            int ia = o.Z[i] - 1;
            double norm = 0;
            for (int ii = 0; ii < d4_refn[ia]; ++ii)
                for (int iii = 1; iii <= dm.ncount[ia][ii]; ++iii) norm += cngw(iii * wf, covcn[i], d4_refcovcn[ia][ii]);
            norm = norm > 1e-80 ? 1.0 / norm : 0.0;
            double maxcn = -1e300;
            for (int ii = 0; ii < d4_refn[ia]; ++ii) maxcn = std::max(maxcn, d4_refcovcn[ia][ii]);
            for (int ii = 0; ii < d4_refn[ia]; ++ii) {
                int k = itbl[i] + ii;
                for (int iii = 1; iii <= dm.ncount[ia][ii]; ++iii) gw[k] += cngw(iii * wf, covcn[i], d4_refcovcn[ia][ii]) * norm;
                if (norm == 0.0) gw[k] = (std::fabs(maxcn - d4_refcovcn[ia][ii]) < 1e-12) ? 1.0 : 0.0;
            }
        }
        // build_wdispmat, src/disp/dftd4.F90:660-732 ; s6=1, s8=2.7, a1=0.52, a2=5 (gfn2.f90:65-66), s10 = 0
        const double s6 = 1.0, s8 = 2.7, a1 = 0.52, a2 = 5.0;
        // pre-touch the c6 cache serially (not thread safe)
        for (int i = 0; i < n; ++i) for (int j = 0; j <= i; ++j) dm.c6(o.Z[i], o.Z[j]);
#pragma omp parallel for schedule(dynamic, 8)
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < i; ++j) {
                int ia = o.Z[i], ja = o.Z[j];
                int k = itbl[i], l = itbl[j];
                double r4r2ij = 3.0 * el_sqrt_zr4r2(ia) * el_sqrt_zr4r2(ja);
                double cutoff = a1 * std::sqrt(r4r2ij) + a2;
                double r2 = 0;
                for (int c = 0; c < 3; ++c) { double t = R(o, j)[c] - R(o, i)[c]; r2 += t * t; }
                if (r2 > 72.0 * 72.0) continue;
                double oor6 = 1.0 / (r2 * r2 * r2 + std::pow(cutoff, 6));
                double oor8 = 1.0 / (r2 * r2 * r2 * r2 + std::pow(cutoff, 8));
                const std::vector<double>& c6 = dm.c6cache.find(std::make_pair(ia, ja))->second;
                for (int ii = 0; ii < d4_refn[ia - 1]; ++ii)
                    for (int jj = 0; jj < d4_refn[ja - 1]; ++jj) {
                        double gwgw = gw[k + ii] * gw[l + jj];
                        if (gwgw < 1.0e-7) continue;
                        cnt_add(7, 1);
                        double c6ab = c6[ii * 7 + jj];
                        double c8 = r4r2ij * c6ab;
                        double v = gw[k + ii] * gw[l + jj] * (-s6 * (c6ab * oor6) - s8 * (c8 * oor8) - 0.0);
                        dispmat[(k + ii) + (size_t)ndim * (l + jj)] = v;
                        dispmat[(l + jj) + (size_t)ndim * (k + ii)] = v;
                    }
            }
    }
    o.timers[2] = now_s() - t0;
    t0 = now_s();

    // ---- integrals (scf_module.F90:544-604)
    // getSelfEnergy, hamiltonian.F90:98-141
    dvec selfE(nsh), dSEdcn(nsh);
    for (int s = 0; s < nsh; ++s) {
        int z = o.Z[b.sh[s].atom];
        double kcn = gfn2_kcn_l[z - 1][b.sh[s].l];
        selfE[s] = b.sh[s].level - kcn * cn[b.sh[s].atom];
        dSEdcn[s] = -kcn;
    }
    size_t nn = (size_t)nao * nao, ntri = (size_t)nao * (nao + 1) / 2;
    dvec S(nn, 0.0), Dp(3 * nn, 0.0), Qp(6 * nn, 0.0), H0(ntri, 0.0), H0n(ntri, 0.0);
    const double point[3] = {0, 0, 0};
    // build_SDQH0, hamiltonian.F90:146-371
#pragma omp parallel for schedule(dynamic, 4)
    for (int iat = 0; iat < n; ++iat)
        for (int jat = 0; jat < iat; ++jat) {
            const double* ra = R(o, iat);
            const double* rb = R(o, jat);
            int izp = o.Z[iat], jzp = o.Z[jat];
            for (int is = b.sh0[iat]; is < b.sh0[iat] + b.nsh[iat]; ++is)
                for (int js = b.sh0[jat]; js < b.sh0[jat] + b.nsh[jat]; ++js) {
                    const Shell& si = b.sh[is];
                    const Shell& sj = b.sh[js];
                    double hii = selfE[is], hjj = selfE[js];
                    double zi = si.zeta, zj = sj.zeta;
                    double zetaij = std::pow(2 * std::sqrt(zi * zj) / (zi + zj), 0.5);
                    double km = h0scal(si.l, sj.l, izp, jzp);
                    double hav = 0.5 * km * (hii + hjj) * zetaij;
                    double shpoly = shell_poly(gfn2_shell_poly_l[izp - 1][si.l], gfn2_shell_poly_l[jzp - 1][sj.l],
                                               el_atomic_rad(izp), el_atomic_rad(jzp), ra, rb);
                    double blk[10][6][6];
                    get_multiints(si, sj, ra, rb, point, intcut, blk);
                    for (int c = 0; c < 10; ++c) cao2sao(blk[c], si.l, sj.l);
                    for (int ii = 0; ii < si.nao; ++ii) {
                        int iao = si.ao0 + ii;
                        for (int jj = 0; jj < sj.nao; ++jj) {
                            int jao = sj.ao0 + jj;
                            size_t ij = (size_t)iao * (iao + 1) / 2 + jao;
                            H0[ij] += hav * shpoly * blk[0][ii][jj];
                            H0n[ij] += hav * shpoly;
                            size_t a = jao + (size_t)nao * iao, a2 = iao + (size_t)nao * jao;
                            S[a] = S[a2] = blk[0][ii][jj];
                            for (int c = 0; c < 3; ++c) Dp[c + 3 * a] = Dp[c + 3 * a2] = blk[1 + c][ii][jj];
                            for (int c = 0; c < 6; ++c) Qp[c + 6 * a] = Qp[c + 6 * a2] = blk[4 + c][ii][jj];
                        }
                    }
                }
        }
    // diagonal blocks, hamiltonian.F90:311-369
#pragma omp parallel for schedule(dynamic, 8)
    for (int iat = 0; iat < n; ++iat) {
        const double* ra = R(o, iat);
        for (int is = b.sh0[iat]; is < b.sh0[iat] + b.nsh[iat]; ++is) {
            const Shell& si = b.sh[is];
            for (int ii = 0; ii < si.nao; ++ii) {
                int i = si.ao0 + ii;
                size_t d = (size_t)i * (i + 1) / 2 + i;
                S[i + (size_t)nao * i] = 1.0 + S[i + (size_t)nao * i];
                H0[d] += selfE[is];
                H0n[d] += selfE[is];
            }
            for (int js = b.sh0[iat]; js <= is; ++js) {
                const Shell& sj = b.sh[js];
                double blk[10][6][6];
                get_multiints(si, sj, ra, ra, point, intcut, blk);
                for (int c = 1; c < 10; ++c) cao2sao(blk[c], si.l, sj.l);
                for (int ii = 0; ii < si.nao; ++ii) {
                    int iao = si.ao0 + ii;
                    for (int jj = 0; jj < sj.nao; ++jj) {
                        int jao = sj.ao0 + jj;
                        if (jao > iao && is == js) continue;
                        size_t a = iao + (size_t)nao * jao, a2 = jao + (size_t)nao * iao;
                        for (int c = 0; c < 3; ++c) {
                            Dp[c + 3 * a] += blk[1 + c][ii][jj];
                            if (iao != jao) Dp[c + 3 * a2] += blk[1 + c][ii][jj];
                        }
                        for (int c = 0; c < 6; ++c) {
                            Qp[c + 6 * a] += blk[4 + c][ii][jj];
                            if (iao != jao) Qp[c + 6 * a2] += blk[4 + c][ii][jj];
                        }
                    }
                }
            }
        }
    }
    // setdqlist, aespot.F90:34-85 (lists of (i,j), j<=i, 0-based here)
    ivec mdlst, mqlst;
    {
        double thr2 = neglect * 1.0e-2;
        for (int i = 0; i < nao; ++i)
            for (int j = 0; j <= i; ++j) {
                size_t a = i + (size_t)nao * j;
                double tmp1 = 0, tmp2 = 0;
                for (int k = 0; k < 3; ++k) { tmp1 += Dp[k + 3 * a] * Dp[k + 3 * a]; tmp2 -= Qp[k + 6 * a] * Qp[k + 6 * a]; }
                for (int k = 0; k < 6; ++k) tmp2 += 2.0 * Qp[k + 6 * a] * Qp[k + 6 * a];
                if (tmp1 > thr2) { mdlst.push_back(i); mdlst.push_back(j); }
                if (tmp2 > thr2) { mqlst.push_back(i); mqlst.push_back(j); }
            }
    }
    int ndp = (int)mdlst.size() / 2, nqp = (int)mqlst.size() / 2;
    // get_radcn, aespot.F90:1267-1281 ; mmomgabzero, aespot.F90:1219-1255
    dvec radcn(n), gab3((size_t)n * n, 0.0), gab5((size_t)n * n, 0.0);
    const double cnShift = 1.2, cnExp = 4.0, cnRMax = 5.0, dipDamp = 3.0, quadDamp = 4.0;
    for (int i = 0; i < n; ++i) {
        double rco = xtb_multi_rad[o.Z[i] - 1];
        double t1 = cn[i] - xtb_valence_cn[o.Z[i] - 1] - cnShift;
        radcn[i] = rco + (cnRMax - rco) / (1.0 + std::exp(-cnExp * t1));
    }
#pragma omp parallel for schedule(dynamic, 16)
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < i; ++j) {
            double tmp2 = 0;
            for (int k = 0; k < 3; ++k) { double t = R(o, j)[k] - R(o, i)[k]; tmp2 += t * t; }
            double tmp1 = 1.0 / std::sqrt(tmp2), damp, ddamp;
            dzero(dipDamp, tmp1, radcn[i], radcn[j], damp, ddamp);
            double g3 = damp * std::pow(tmp1, 3.0);
            dzero(quadDamp, tmp1, radcn[i], radcn[j], damp, ddamp);
            double g5 = damp * std::pow(tmp1, 5.0);
            gab3[j + (size_t)n * i] = gab3[i + (size_t)n * j] = g3;
            gab5[j + (size_t)n * i] = gab5[i + (size_t)n * j] = g5;
        }
    o.timers[4] = now_s() - t0;
    t0 = now_s();

    // ---- SCC (scc_core.f90:255-640)
    dvec Sf = S;
    int info = 0;
    scipy_dpotrf_("U", &nao, Sf.data(), &nao, &info);
    if (info != 0) return 4;
    int nbr = nsh + 9 * n;
    Broyden br;
    br.init(nbr, o.maxiter);
    dvec H(nn), P(nn, 0.0), emo(nao, 0.0), shellShift(nsh), atomicShift(n), vs, vd, vq, q_in(nbr), dq(nbr);
    dvec zetavec(ndim), zerovec(ndim), dumvec(ndim);
    double eel = 0, ees = 0, eaes = 0, epol = 0, ed = 0, egap = 0, ga = 0, gb = 0, efa = 0, efb = 0;
    bool converged = false, lastdiag = false;
    int iter = 0, lwork = -1, liwork = -1;
    dvec work(1); ivec iwork(1);
    {
        double wq; int iq;
        scipy_dsyevd_("V", "U", &nao, H.data(), &nao, emo.data(), &wq, &lwork, &iq, &liwork, &info);
        lwork = (int)wq; liwork = iq;
        work.resize(lwork); iwork.resize(liwork);
    }
    const double one = 1.0, zero = 0.0, half = 0.5;
    const int ione = 1;
    const double g_a = 3.0, g_c = 2.0;
    double t_eig = 0;
    // potential of the external point charges on the shells: jpot_pcem_gfn2, src/embedding.f90:225-256
    const int npc = (int)o.pc_q.size();
    dvec Vpc(nsh, 0.0);
    double epcem = 0.0;
    for (int s = 0; s < nsh && npc > 0; ++s) {
        const double* ri = R(o, b.sh[s].atom);
        double eh1 = 0.0;
        for (int kk = 0; kk < npc; ++kk) {
            double r2 = 0;
            for (int k = 0; k < 3; ++k) { double dx = o.pc_xyz[3 * kk + k] - ri[k]; r2 += dx * dx; }
            const double xj = 0.5 * (hard[s] + o.pc_gam[kk]);
            eh1 += o.pc_q[kk] / std::sqrt(r2 + 1.0 / (xj * xj));
        }
        Vpc[s] = eh1;
    }
    // implicit solvation: solvation%update, scf_module.F90:299-320 (GFN2: cm5a = 0, the model sees the Mulliken charges)
    Solv sv;
    double gsolv_scc = 0.0;
    if (o.solv_on) {
        if (!solv_init(o, sv)) return 7;
        solv_update(o, sv);
        solv_born_matrix(o, sv);
    }
    for (iter = 1; iter <= o.maxiter; ++iter) {
        // potentials: thirdorder.f90:100-131, coulomb.f90:89-114
        for (int i = 0; i < n; ++i) atomicShift[i] = 0.0;
        // shellShift(:) = externShift, scc_core.f90:438
        for (int s = 0; s < nsh; ++s) shellShift[s] = Vpc[s] + qsh[s] * qsh[s] * gam3[s];
        scipy_dsymv_("L", &nsh, &one, jmat.data(), &nsh, qsh.data(), &ione, &one, shellShift.data(), &ione);
        setvsdq(o, true, q, dipm, qp, gab3, gab5, vs, vd, vq);
        // solvation%addShift, scc_core.f90:444-448 -> gbsa.f90:482-507: atomicShift += bornMat q
        if (o.solv_on)
            for (int i = 0; i < n; ++i) {
                double t = 0;
                for (int j = 0; j < n; ++j) t += sv.bornmat[i + (size_t)n * j] * q[j];
                atomicShift[i] += t;
            }
        // disppot, dftd4.F90:735-818
        for (int i = 0; i < n; ++i) {
            int ati = o.Z[i] - 1;
            double iz = d4_zeff[ati];
            for (int ii = 0; ii < d4_refn[ati]; ++ii) {
                int k = itbl[i] + ii;
                zetavec[k] = zerovec[k] = 0.0;
                if (gw[k] < 1.0e-7) continue;
                zerovec[k] = d4_dzeta(g_a, d4_gam[ati] * g_c, d4_refq[ati][ii] + iz, q[i] + iz);
                zetavec[k] = d4_zeta(g_a, d4_gam[ati] * g_c, d4_refq[ati][ii] + iz, q[i] + iz);
            }
        }
        scipy_dsymv_("U", &ndim, &one, dispmat.data(), &ndim, zetavec.data(), &ione, &zero, dumvec.data(), &ione);
        for (int i = 0; i < n; ++i)
            for (int ii = 0; ii < d4_refn[o.Z[i] - 1]; ++ii) atomicShift[i] += dumvec[itbl[i] + ii] * zerovec[itbl[i] + ii];
        for (int s = 0; s < nsh; ++s) shellShift[s] += atomicShift[b.sh[s].atom];

        // buildIsoAnisotropicH1, scc_core.f90:154-251
#pragma omp parallel for schedule(static)
        for (int i = 0; i < nao; ++i) {
            int ish = b.ao2sh[i], iat = b.ao2at[i];
            for (int j = 0; j <= i; ++j) {
                size_t k = (size_t)i * (i + 1) / 2 + j;
                size_t ji = j + (size_t)nao * i;
                double dum = S[ji] * AUTOEV * 0.5;
                int jsh = b.ao2sh[j], jat = b.ao2at[j];
                double eh1 = -dum * (shellShift[ish] + shellShift[jsh]);
                double h = H0[k] + eh1;
                eh1 = dum * (vs[iat] + vs[jat]);
                h = h + eh1;
                H[ji] = h; H[i + (size_t)nao * j] = h;
            }
        }
#pragma omp parallel for schedule(static)
        for (int m = 0; m < ndp; ++m) {
            int i = mdlst[2 * m], j = mdlst[2 * m + 1];
            int ii = b.ao2at[i], jj = b.ao2at[j];
            size_t ij = i + (size_t)nao * j;
            double eh1 = 0;
            for (int l = 0; l < 3; ++l) eh1 += Dp[l + 3 * ij] * (vd[3 * ii + l] + vd[3 * jj + l]);
            eh1 = 0.50 * eh1 * AUTOEV;
            H[ij] += eh1; H[j + (size_t)nao * i] = H[ij];
        }
#pragma omp parallel for schedule(static)
        for (int m = 0; m < nqp; ++m) {
            int i = mqlst[2 * m], j = mqlst[2 * m + 1];
            int ii = b.ao2at[i], jj = b.ao2at[j];
            size_t ij = i + (size_t)nao * j;
            double eh1 = 0;
            for (int l = 0; l < 6; ++l) eh1 += Qp[l + 6 * ij] * (vq[6 * ii + l] + vq[6 * jj + l]);
            eh1 = 0.50 * eh1 * AUTOEV;
            H[ij] += eh1; H[j + (size_t)nao * i] = H[ij];
        }
        if (iter == 1) D["fock_first"] = H;

        // fact_solve, src/mctc/lapack/eigensolve.F90:218-246
        double te = now_s();
        scipy_dsygst_(&ione, "U", &nao, H.data(), &nao, Sf.data(), &nao, &info);
        if (info != 0) return 5;
        scipy_dsyevd_("V", "U", &nao, H.data(), &nao, emo.data(), work.data(), &lwork, iwork.data(), &liwork, &info);
        if (info != 0) return 5;
        scipy_dtrsm_("L", "U", "N", "N", &nao, &nao, &one, Sf.data(), &nao, H.data(), &nao);
        t_eig += now_s() - te;

        if (ihomo + 1 <= nao && ihomo >= 1) egap = emo[ihomo] - emo[ihomo - 1];
        if (o.etemp > 0.1) {
            if (nel > 0) occu(nao, nel, nopen, ihomoa, ihomob, focca, foccb);
            else { focca.assign(nao, 0.0); foccb.assign(nao, 0.0); ihomoa = ihomob = 0; }
            fermismear(nao, ihomoa, o.etemp, emo.data(), focca.data(), efa, ga);
            fermismear(nao, ihomob, o.etemp, emo.data(), foccb.data(), efb, gb);
            for (int i = 0; i < nao; ++i) focc[i] = focca[i] + foccb[i];
        } else { ga = gb = 0; }

        for (int s = 0; s < nsh; ++s) q_in[s] = qsh[s];
        for (int i = 0; i < n; ++i) {
            for (int k = 0; k < 3; ++k) q_in[nsh + 9 * i + k] = dipm[3 * i + k];
            for (int k = 0; k < 6; ++k) q_in[nsh + 9 * i + 3 + k] = qp[6 * i + k];
        }
        dmat(nao, focc.data(), H.data(), P.data());

        // mpopsh, scc_core.f90:1408-1426
        {
            dvec pop(nsh, 0.0);
            for (int i = 0; i < nao; ++i) {
                int ii = b.ao2sh[i];
                for (int j = 0; j < i; ++j) {
                    double ps = P[j + (size_t)nao * i] * S[j + (size_t)nao * i];
                    pop[ii] += ps; pop[b.ao2sh[j]] += ps;
                }
                pop[ii] += P[i + (size_t)nao * i] * S[i + (size_t)nao * i];
            }
            for (int s = 0; s < nsh; ++s) qsh[s] = zsh[s] - pop[s];
        }
        for (int i = 0; i < n; ++i) q[i] = 0;
        for (int s = 0; s < nsh; ++s) q[b.sh[s].atom] += qsh[s];

        double eold = eel;
        // electro, scc_core.f90:686-722 + coulomb.f90:117-142 + thirdorder.f90:134-164
        {
            double eThird = 0;
            for (int s = 0; s < nsh; ++s) eThird += qsh[s] * qsh[s] * qsh[s] * gam3[s] / 3.0;
            dvec shift(nsh);
            scipy_dsymv_("L", &nsh, &one, jmat.data(), &nsh, qsh.data(), &ione, &zero, shift.data(), &ione);
            double dot = 0;
            for (int s = 0; s < nsh; ++s) dot += shift[s] * qsh[s];
            ees = 0.5 * dot + eThird;
            double h = 0;
            size_t k = 0;
            for (int i = 0; i < nao; ++i) {
                for (int j = 0; j < i; ++j) { h += P[j + (size_t)nao * i] * H0[k]; k++; }
                h += P[i + (size_t)nao * i] * H0[k] * 0.5; k++;
            }
            eel = ees + 2.0 * h * EVTOAU;
        }
        mmompop(o, P.data(), S.data(), Dp.data(), Qp.data(), dipm, qp);
        aniso_electro(o, q, dipm, qp, gab3, gab5, eaes, epol);
        eel = eel + eaes + epol;
        // edisp_scc, dftd4.F90:821-864
        for (int i = 0; i < n; ++i) {
            int ati = o.Z[i] - 1;
            double iz = d4_zeff[ati];
            for (int ii = 0; ii < d4_refn[ati]; ++ii) {
                int k = itbl[i] + ii;
                zetavec[k] = 0.0;
                if (gw[k] < 1.0e-7) continue;
                zetavec[k] = d4_zeta(g_a, d4_gam[ati] * g_c, d4_refq[ati][ii] + iz, q[i] + iz);
            }
        }
        scipy_dsymv_("U", &ndim, &half, dispmat.data(), &ndim, zetavec.data(), &ione, &zero, dumvec.data(), &ione);
        ed = 0;
        for (int k = 0; k < ndim; ++k) ed += dumvec[k] * zetavec[k];
        eel = eel + ed;
        // electro_pcem, src/embedding.f90:261-280, called at scc_core.f90:539-541
        if (npc > 0) {
            epcem = 0.0;
            for (int s = 0; s < nsh; ++s) epcem += qsh[s] * Vpc[s];
            eel += epcem;
        }
        // solvation%getEnergy, scc_core.f90:543-548 -> gbsa.f90:510-535 (output charges of this iteration)
        if (o.solv_on) {
            double e = 0;
            for (int i = 0; i < n; ++i) {
                double t = 0;
                for (int j = 0; j < n; ++j) t += 0.5 * sv.bornmat[i + (size_t)n * j] * q[j];
                e += q[i] * t;
            }
            gsolv_scc = e + sv.gsasa + sv.gshift;
            eel += gsolv_scc;
        }
        eel = eel + ga + gb;

        bool econverged = std::fabs(eel - eold) < scfconv;
        for (int s = 0; s < nsh; ++s) dq[s] = qsh[s] - q_in[s];
        for (int i = 0; i < n; ++i) {
            for (int k = 0; k < 3; ++k) dq[nsh + 9 * i + k] = dipm[3 * i + k] - q_in[nsh + 9 * i + k];
            for (int k = 0; k < 6; ++k) dq[nsh + 9 * i + 3 + k] = qp[6 * i + k] - q_in[nsh + 9 * i + 3 + k];
        }
        double rmsq = 0;
        for (int k = 0; k < nbr; ++k) rmsq += dq[k] * dq[k];
        rmsq = std::sqrt(rmsq / (double)n);
        bool qconverged = rmsq < qconv;
        if (iter == 1) { D["q_out_first"] = q; D["eel_first"] = dvec(1, eel); }

        if (broy) {
            br.step(q_in, dq, iter, o.broydamp);
            for (int s = 0; s < nsh; ++s) qsh[s] = q_in[s];
            for (int i = 0; i < n; ++i) {
                for (int k = 0; k < 3; ++k) dipm[3 * i + k] = q_in[nsh + 9 * i + k];
                for (int k = 0; k < 6; ++k) qp[6 * i + k] = q_in[nsh + 9 * i + 3 + k];
            }
        }
        // (single-atom systems keep the undamped output: scc_core.f90:575 `iter.gt.0` damping with damp0
        //  is only reachable for n == 1; not needed for the parity configs)
        for (int i = 0; i < n; ++i) q[i] = 0;
        for (int s = 0; s < nsh; ++s) q[b.sh[s].atom] += qsh[s];
        D["iter_energy"].push_back(eel);
        D["iter_rmsq"].push_back(rmsq);

        if (econverged && qconverged) {
            converged = true;
            if (lastdiag) break;
            lastdiag = true;
        }
    }
    int jter = std::min(iter, o.maxiter);
    o.timers[5] = now_s() - t0;
    o.timers[7] = t_eig;
    t0 = now_s();

    // ---- gradient (scf_module.F90:665-798)
    dvec Pew(nn);
    {
        dvec tmp(nao);
        for (int i = 0; i < nao; ++i) tmp[i] = focc[i] * emo[i] * EVTOAU;
        dmat(nao, tmp.data(), H.data(), Pew.data());
    }
    dvec dvs, dvd, dvq;
    setvsdq(o, false, q, dipm, qp, gab3, gab5, dvs, dvd, dvq);
    dvec dhdcn(n, 0.0);
    {
        // build_dSDQH0_noreset, hamiltonian.F90:575-792
        int nth = o.nthreads;
        std::vector<dvec> gl(nth, dvec(3 * n, 0.0)), dl(nth, dvec(n, 0.0));
#pragma omp parallel num_threads(nth)
        {
#ifdef _OPENMP
            int tid = omp_get_thread_num();
#else
            int tid = 0;
#endif
            double* g = gl[tid].data();
            double* dh = dl[tid].data();
#pragma omp for schedule(dynamic, 2)
            for (int iat = 0; iat < n; ++iat)
                for (int jat = 0; jat < iat; ++jat) {
                    const double* ri = R(o, iat);
                    const double* rj = R(o, jat);
                    int izp = o.Z[iat], jzp = o.Z[jat];
                    double rij[3], rij2 = 0;
                    for (int k = 0; k < 3; ++k) { rij[k] = ri[k] - rj[k]; rij2 += rij[k] * rij[k]; }
                    if (rij2 > 1600.0) continue;
                    for (int is = b.sh0[iat]; is < b.sh0[iat] + b.nsh[iat]; ++is)
                        for (int js = b.sh0[jat]; js < b.sh0[jat] + b.nsh[jat]; ++js) {
                            cnt_add(2, 1);
                            const Shell& si = b.sh[is];
                            const Shell& sj = b.sh[js];
                            double hii = selfE[is], hjj = selfE[js];
                            // dshellPoly, src/grad_core.f90:40-72
                            double shpoly, dshpoly[3];
                            {
                                double iPoly = gfn2_shell_poly_l[izp - 1][si.l], jPoly = gfn2_shell_poly_l[jzp - 1][sj.l];
                                double rab = std::sqrt(rij2);
                                double rr = el_atomic_rad(jzp) + el_atomic_rad(izp);
                                double r = rab / rr;
                                double rf1 = 1.0 + 0.01 * iPoly * std::sqrt(r);
                                double rf2 = 1.0 + 0.01 * jPoly * std::sqrt(r);
                                double dum = 0.5 * (0.01 * iPoly * rf2 + 0.01 * jPoly * rf1) * std::sqrt(r) / rij2;
                                shpoly = rf1 * rf2;
                                for (int k = 0; k < 3; ++k) dshpoly[k] = dum * rij[k];
                            }
                            double hav = 0.5 * (hii + hjj);
                            double sdqg[19][3][6][6];
                            get_grad_multiint(si, sj, ri, rj, intcut, sdqg);
                            for (int c = 0; c < 19; ++c) for (int k = 0; k < 3; ++k) cao2sao(sdqg[c][k], si.l, sj.l);
                            double gx[3] = {0, 0, 0}, dCN = 0;
                            for (int ii = 0; ii < si.nao; ++ii) {
                                int iao = si.ao0 + ii;
                                for (int jj = 0; jj < sj.nao; ++jj) {
                                    int jao = sj.ao0 + jj;
                                    size_t ji = jao + (size_t)nao * iao;
                                    double Pij = P[ji];
                                    double Hij = H0n[(size_t)iao * (iao + 1) / 2 + jao] * EVTOAU;
                                    double HPij = Hij * Pij;
                                    double f = 2 * HPij - 2 * Pew[ji] - Pij * (shellShift[is] + shellShift[js]) + Pij * (dvs[iat] + dvs[jat]);
                                    for (int k = 0; k < 3; ++k) {
                                        gx[k] += 2 * HPij * S[ji] * dshpoly[k] / shpoly + sdqg[0][k][ii][jj] * f;
                                        double dtmp = 0, qtmp = 0;
                                        for (int c = 0; c < 3; ++c)
                                            dtmp += sdqg[10 + c][k][ii][jj] * dvd[3 * iat + c] + sdqg[1 + c][k][ii][jj] * dvd[3 * jat + c];
                                        for (int c = 0; c < 6; ++c)
                                            qtmp += sdqg[13 + c][k][ii][jj] * dvq[6 * iat + c] + sdqg[4 + c][k][ii][jj] * dvq[6 * jat + c];
                                        gx[k] += Pij * dtmp + Pij * qtmp;
                                    }
                                    dCN += HPij / hav * S[ji];
                                }
                            }
                            dh[iat] += dCN * dSEdcn[is];
                            dh[jat] += dCN * dSEdcn[js];
                            for (int k = 0; k < 3; ++k) { g[3 * iat + k] += gx[k]; g[3 * jat + k] -= gx[k]; }
                        }
                }
        }
        for (int t = 0; t < nth; ++t) {
            for (int k = 0; k < 3 * n; ++k) grad[k] += gl[t][k];
            for (int k = 0; k < n; ++k) dhdcn[k] += dl[t][k];
        }
        for (int s = 0; s < nsh; ++s)
            for (int ii = 0; ii < b.sh[s].nao; ++ii) {
                int i = b.sh[s].ao0 + ii;
                dhdcn[b.sh[s].atom] += P[i + (size_t)nao * i] * dSEdcn[s] * EVTOAU;
            }
    }
    cn_chain(cnpairs, dhdcn, grad);
    D["grad_h0"] = grad;
    {
        // dradcn (aespot.F90:1292-1307) + aniso_grad (aespot.F90:1070-1165)
        dvec w(n, 0.0);
        dvec gloc(3 * n, 0.0);
#pragma omp parallel for schedule(dynamic, 8)
        for (int i = 0; i < n; ++i) {
            double q1 = q[i];
            const double* rr = R(o, i);
            const double* dip = &dipm[3 * i];
            const double* qp1 = &qp[6 * i];
            double tmp2 = 0, gi[3] = {0, 0, 0};
            for (int j = 0; j < n; ++j) {
                if (i == j) continue;
                double rij[3], r2 = 0;
                for (int k = 0; k < 3; ++k) { rij[k] = R(o, j)[k] - rr[k]; r2 += rij[k] * rij[k]; }
                double rabi = 1.0 / std::sqrt(r2), damp1, damp2, ddamp;
                dzero(dipDamp, rabi, radcn[i], radcn[j], damp1, ddamp);
                double dgab3 = (-3.0 * rabi * std::pow(rabi, 3.0)) * damp1 + ddamp * std::pow(rabi, 3.0);
                dzero(quadDamp, rabi, radcn[i], radcn[j], damp2, ddamp);
                double dgab5 = (-5.0 * rabi * std::pow(rabi, 5.0)) * damp2 + ddamp * std::pow(rabi, 5.0);
                double g3 = gab3[j + (size_t)n * i], g5 = gab5[j + (size_t)n * i];
                double ed = 0, edd = 0, eq = 0;
                for (int k = 0; k < 3; ++k) {
                    ed += q[j] * dip[k] * rij[k];
                    ed -= dipm[3 * j + k] * q1 * rij[k];
                    double tt = q1 * dipm[3 * j + k] - q[j] * dip[k];
                    gi[k] += g3 * tt;
                    double ddm2 = 0, ddm3a = 0, ddm3b = 0, qqa = 0, qqb = 0;
                    for (int l = 0; l < 3; ++l) {
                        int kl = lin0(l, k);
                        double t2 = rij[l] * rij[k], tt3 = 3.0 * t2;
                        eq += q[j] * qp1[kl] * t2;
                        eq += qp[6 * j + kl] * q1 * t2;
                        edd -= dipm[3 * j + k] * dip[l] * tt3;
                        ddm2 += dipm[3 * j + l] * dip[l];
                        ddm3a += dip[l] * rij[l];
                        ddm3b += dipm[3 * j + l] * rij[l];
                        qqa += rij[l] * qp[6 * j + kl];
                        qqb += rij[l] * qp1[kl];
                    }
                    edd += dipm[3 * j + k] * dip[k] * r2;
                    gi[k] -= 2.0 * g5 * ddm2 * rij[k];
                    gi[k] += 3.0 * g5 * ddm3a * dipm[3 * j + k];
                    gi[k] += 3.0 * g5 * ddm3b * dip[k];
                    gi[k] -= 2.0 * g5 * qqa * q1;
                    gi[k] -= 2.0 * g5 * qqb * q[j];
                }
                for (int k = 0; k < 3; ++k) {
                    double dxi = rij[k] * rabi;
                    gi[k] -= ed * dgab3 * dxi;
                    gi[k] -= (eq + edd) * dgab5 * dxi;
                }
                double rab = 0.50 * (radcn[i] + radcn[j]);
                tmp2 += ed * dipDamp * g3 * (damp1 / rab) * std::pow(rab * rabi, dipDamp);
                tmp2 += (eq + edd) * quadDamp * g5 * (damp2 / rab) * std::pow(rab * rabi, quadDamp);
            }
            tmp2 = 3.0 * tmp2;
            // dradcn factor
            double rco = xtb_multi_rad[o.Z[i] - 1];
            double t1 = std::exp(-cnExp * (cn[i] - xtb_valence_cn[o.Z[i] - 1] - cnShift));
            double t2 = (cnRMax - rco) / (1.0 + 2.0 * t1 + t1 * t1) * cnExp * t1;
            w[i] = -tmp2 * t2;
            for (int k = 0; k < 3; ++k) gloc[3 * i + k] = gi[k];
        }
        for (int k = 0; k < 3 * n; ++k) grad[k] += gloc[k];
        cn_chain(cnpairs, w, grad);
    }
    D["grad_aes"] = grad;
    double embd = 0;
    {
        // D4 gradient: CN(cov) + d4_full_gradient_latp, dftd4.F90:1624-1734
        dvec cnc;
        std::vector<CNPair> covpairs;
        cn_lp(o, false, cnc, covpairs);
        D["cn_cov"] = cnc;
        const double wf = 6.0, s6 = 1.0, s8 = 2.7, a1 = 0.52, a2 = 5.0, s9 = 5.0;
        const int alp = 16;
        // weight_references, dftd4.F90:948-1065
        dvec zv(7 * n, 0.0), z0(7 * n, 0.0), zdcn(7 * n, 0.0), z0dcn(7 * n, 0.0), zdq(7 * n, 0.0);
        for (int i = 0; i < n; ++i) {
            int ati = o.Z[i] - 1;
            double zi = d4_zeff[ati], gi = g_c * d4_gam[ati];
            double norm = 0, dnorm = 0;
            for (int ir = 0; ir < d4_refn[ati]; ++ir)
                for (int ic = 1; ic <= dm.ncount[ati][ir]; ++ic) {
                    double twf = ic * wf, g = cngw(twf, cnc[i], d4_refcovcn[ati][ir]);
                    norm += g;
                    dnorm += 2 * twf * (d4_refcovcn[ati][ir] - cnc[i]) * g;
                }
            norm = norm > 1e-80 ? 1.0 / norm : 0.0;
            double maxcn = -1e300;
            for (int ir = 0; ir < d4_refn[ati]; ++ir) maxcn = std::max(maxcn, d4_refcovcn[ati][ir]);
            for (int ir = 0; ir < d4_refn[ati]; ++ir) {
                double expw = 0, expd = 0;
                for (int ic = 1; ic <= dm.ncount[ati][ir]; ++ic) {
                    double twf = ic * wf, g = cngw(twf, cnc[i], d4_refcovcn[ati][ir]);
                    expw += g;
                    expd += 2 * twf * (d4_refcovcn[ati][ir] - cnc[i]) * g;
                }
                double gwk = expw * norm;
                if (norm == 0.0) gwk = (std::fabs(maxcn - d4_refcovcn[ati][ir]) < 1e-12) ? 1.0 : 0.0;
                double zq = d4_zeta(g_a, gi, d4_refq[ati][ir] + zi, q[i] + zi);
                double z00 = d4_zeta(g_a, gi, d4_refq[ati][ir] + zi, zi);
                double dgwk = expd * norm - expw * dnorm * norm * norm;
                zv[7 * i + ir] = zq * gwk;
                z0[7 * i + ir] = z00 * gwk;
                zdcn[7 * i + ir] = zq * dgwk;
                zdq[7 * i + ir] = d4_dzeta(g_a, gi, d4_refq[ati][ir] + zi, q[i] + zi) * gwk;
                z0dcn[7 * i + ir] = z00 * dgwk;
            }
        }
        // get_atomic_c6, dftd4.F90:1069-1147 ; c6[i*n+j], dc6dcn[i*n+j] = dC6ij/dCNi
        dvec c6a((size_t)n * n), dc6a((size_t)n * n);
        auto atomic_c6 = [&](const dvec& zvv, const dvec& zdd) {
#pragma omp parallel for schedule(dynamic, 16)
            for (int i = 0; i < n; ++i)
                for (int j = 0; j <= i; ++j) {
                    int ati = o.Z[i], atj = o.Z[j];
                    const std::vector<double>& rc6 = dm.c6cache.find(std::make_pair(ati, atj))->second;
                    double dc6 = 0, di = 0, dj = 0;
                    for (int ir = 0; ir < d4_refn[ati - 1]; ++ir)
                        for (int jr = 0; jr < d4_refn[atj - 1]; ++jr) {
                            double r = rc6[ir * 7 + jr];
                            dc6 += zvv[7 * i + ir] * zvv[7 * j + jr] * r;
                            di += zdd[7 * i + ir] * zvv[7 * j + jr] * r;
                            dj += zvv[7 * i + ir] * zdd[7 * j + jr] * r;
                        }
                    c6a[(size_t)i * n + j] = c6a[(size_t)j * n + i] = dc6;
                    dc6a[(size_t)i * n + j] = di;
                    dc6a[(size_t)j * n + i] = dj;
                }
        };
        atomic_c6(zv, zdcn);
        dvec dEdcn(n, 0.0);
        // disp_gradient_latp, dftd4.F90:1832-1933 (cutoff 60)
        {
            int nth = o.nthreads;
            std::vector<dvec> gl(nth, dvec(3 * n, 0.0)), dl(nth, dvec(n, 0.0));
#pragma omp parallel num_threads(nth)
            {
#ifdef _OPENMP
                int tid = omp_get_thread_num();
#else
                int tid = 0;
#endif
#pragma omp for schedule(dynamic, 8)
                for (int i = 0; i < n; ++i)
                    for (int j = 0; j < i; ++j) {
                        int ati = o.Z[i], atj = o.Z[j];
                        double r4r2ij = 3 * el_sqrt_zr4r2(ati) * el_sqrt_zr4r2(atj);
                        double r0 = a1 * std::sqrt(r4r2ij) + a2;
                        double rij[3], r2 = 0;
                        for (int k = 0; k < 3; ++k) { rij[k] = R(o, i)[k] - R(o, j)[k]; r2 += rij[k] * rij[k]; }
                        if (r2 > 3600.0 || r2 < 1.0e-10) continue;
                        cnt_add(8, 1);
                        double t6 = 1. / (r2 * r2 * r2 + std::pow(r0, 6)), t8 = 1. / (r2 * r2 * r2 * r2 + std::pow(r0, 8));
                        double d6 = -6 * r2 * r2 * t6 * t6, d8 = -8 * r2 * r2 * r2 * t8 * t8;
                        double disp = s6 * t6 + s8 * r4r2ij * t8;
                        double ddisp = s6 * d6 + s8 * r4r2ij * d8;
                        double c6ij = c6a[(size_t)i * n + j];
                        for (int k = 0; k < 3; ++k) {
                            double dG = -c6ij * ddisp * rij[k];
                            gl[tid][3 * i + k] += dG; gl[tid][3 * j + k] -= dG;
                        }
                        dl[tid][i] -= dc6a[(size_t)i * n + j] * disp;
                        dl[tid][j] -= dc6a[(size_t)j * n + i] * disp;
                    }
            }
            for (int t = 0; t < nth; ++t) {
                for (int k = 0; k < 3 * n; ++k) grad[k] += gl[t][k];
                for (int k = 0; k < n; ++k) dEdcn[k] += dl[t][k];
            }
        }
        // ATM, dftd4.F90:2023-2177 + 2354-2433, zero-charge reference weights
        atomic_c6(z0, z0dcn);
        {
            int nth = o.nthreads;
            std::vector<dvec> gl(nth, dvec(3 * n, 0.0)), dl(nth, dvec(n, 0.0));
            dvec el(nth, 0.0);
#pragma omp parallel num_threads(nth)
            {
#ifdef _OPENMP
                int tid = omp_get_thread_num();
#else
                int tid = 0;
#endif
                double* g = gl[tid].data();
                double* dc = dl[tid].data();
#pragma omp for schedule(dynamic, 2)
                for (int iat = 0; iat < n; ++iat)
                    for (int jat = 0; jat < iat; ++jat) {
                        double rij[3], r2ij = 0;
                        for (int k = 0; k < 3; ++k) { rij[k] = R(o, jat)[k] - R(o, iat)[k]; r2ij += rij[k] * rij[k]; }
                        if (r2ij > 1600.0) continue;
                        for (int kat = 0; kat < jat; ++kat) {
                            double rik[3], rjk[3], r2ik = 0, r2jk = 0;
                            for (int k = 0; k < 3; ++k) { rik[k] = R(o, kat)[k] - R(o, iat)[k]; r2ik += rik[k] * rik[k]; }
                            if (r2ik > 1600.0) continue;
                            for (int k = 0; k < 3; ++k) { rjk[k] = R(o, kat)[k] - R(o, jat)[k]; r2jk += rjk[k] * rjk[k]; }
                            if (r2jk > 1600.0) continue;
                            cnt_add(9, 1);
                            int ati = o.Z[iat], atj = o.Z[jat], atk = o.Z[kat];
                            double c6ij = c6a[(size_t)jat * n + iat], c6ik = c6a[(size_t)kat * n + iat], c6jk = c6a[(size_t)kat * n + jat];
                            double cij = a1 * std::sqrt(3.0 * el_sqrt_zr4r2(ati) * el_sqrt_zr4r2(atj)) + a2;
                            double cik = a1 * std::sqrt(3.0 * el_sqrt_zr4r2(ati) * el_sqrt_zr4r2(atk)) + a2;
                            double cjk = a1 * std::sqrt(3.0 * el_sqrt_zr4r2(atj) * el_sqrt_zr4r2(atk)) + a2;
                            double dc6ij = dc6a[(size_t)iat * n + jat], dc6ji = dc6a[(size_t)jat * n + iat];
                            double dc6jk = dc6a[(size_t)jat * n + kat], dc6kj = dc6a[(size_t)kat * n + jat];
                            double dc6ik = dc6a[(size_t)iat * n + kat], dc6ki = dc6a[(size_t)kat * n + iat];
                            // deriv_atm_triple
                            double c9 = -std::sqrt(c6ij * c6ik * c6jk);
                            double ccc1 = cij * cjk * cik;
                            double rrr2 = r2ij * r2jk * r2ik, rrr1 = std::sqrt(rrr2), rrr3 = rrr1 * rrr2;
                            double ang_fact = 0.375 / (rrr2 * rrr3);
                            double ang = ang_fact * (r2ij + r2jk - r2ik) * (r2ij - r2jk + r2ik) * (-r2ij + r2jk + r2ik) + 1.0 / rrr3;
                            double cralp = std::pow(ccc1 / rrr1, (double)alp / 3.0);
                            double fdmp = 1.0 / (1.0 + 6.0 * cralp);
                            double dfdmp = -(2.0 * alp * cralp) * fdmp * fdmp;
                            double dE = -fdmp * ang * c9;
                            double dang, dGr[3];
                            dang = -ang_fact * (r2ij * r2ij * r2ij + r2ij * r2ij * (r2jk + r2ik) +
                                                r2ij * (3.0 * r2jk * r2jk + 2.0 * r2jk * r2ik + 3.0 * r2ik * r2ik) -
                                                5.0 * (r2jk - r2ik) * (r2jk - r2ik) * (r2jk + r2ik));
                            dGr[0] = (-dang * c9 * fdmp + dfdmp * c9 * ang) / r2ij;
                            dang = -ang_fact * (r2ik * r2ik * r2ik + r2ik * r2ik * (r2jk + r2ij) +
                                                r2ik * (3.0 * r2jk * r2jk + 2.0 * r2jk * r2ij + 3.0 * r2ij * r2ij) -
                                                5.0 * (r2jk - r2ij) * (r2jk - r2ij) * (r2jk + r2ij));
                            dGr[1] = (-dang * c9 * fdmp + dfdmp * c9 * ang) / r2ik;
                            dang = -ang_fact * (r2jk * r2jk * r2jk + r2jk * r2jk * (r2ik + r2ij) +
                                                r2jk * (3.0 * r2ik * r2ik + 2.0 * r2ik * r2ij + 3.0 * r2ij * r2ij) -
                                                5.0 * (r2ik - r2ij) * (r2ik - r2ij) * (r2ik + r2ij));
                            dGr[2] = (-dang * c9 * fdmp + dfdmp * c9 * ang) / r2jk;
                            double scale = s9 * 1.0;
                            for (int k = 0; k < 3; ++k) {
                                g[3 * iat + k] += (-dGr[0] * rij[k] - dGr[1] * rik[k]) * scale;
                                g[3 * jat + k] += (dGr[0] * rij[k] - dGr[2] * rjk[k]) * scale;
                                g[3 * kat + k] += (dGr[1] * rik[k] + dGr[2] * rjk[k]) * scale;
                            }
                            el[tid] += dE * scale;
                            dc[iat] += -ang * fdmp * (0.5 * c9 * (dc6ij / c6ij + dc6ik / c6ik)) * scale;
                            dc[jat] += -ang * fdmp * (0.5 * c9 * (dc6ji / c6ij + dc6jk / c6jk)) * scale;
                            dc[kat] += -ang * fdmp * (0.5 * c9 * (dc6ki / c6ik + dc6kj / c6jk)) * scale;
                        }
                    }
            }
            for (int t = 0; t < nth; ++t) {
                for (int k = 0; k < 3 * n; ++k) grad[k] += gl[t][k];
                for (int k = 0; k < n; ++k) dEdcn[k] += dl[t][k];
                embd += el[t];
            }
        }
        cn_chain(covpairs, dEdcn, grad);
    }
    D["grad_d4"] = grad;
    if (o.solv_on) {
        // solvation%addGradient + getEnergyParts with the mixed charges wfn%q, scf_module.F90:759-777
        double gborn = 0, ghb = 0;
        solv_gradient(o, sv, q, grad, gborn, ghb);
        // getEnergyParts, gbsa.f90:538-570: gborn = 1/2 q B q - ghb
        double e = 0;
        for (int i = 0; i < n; ++i) {
            double t = 0;
            for (int j = 0; j < n; ++j) t += 0.5 * sv.bornmat[i + (size_t)n * j] * q[j];
            e += q[i] * t;
        }
        D["solvation"] = dvec{(e - ghb) + sv.gsasa + ghb + sv.gshift, e - ghb, sv.gsasa, ghb, sv.gshift, gsolv_scc, gborn, sv.adet};
        D["brad"] = sv.brad; D["sasa"] = sv.sasa; D["bornmat"] = sv.bornmat; D["dsdr"] = sv.dsdr;
    }
    if (npc > 0) {
        // pcem_grad_gfn2, src/embedding.f90:329-365 (scf_module.F90:790-798): wfn%qsh = the mixed shell charges
        dvec gpc(3 * npc, 0.0);
        for (int s = 0; s < nsh; ++s) {
            const int iat = b.sh[s].atom;
            for (int j = 0; j < npc; ++j) {
                double dr[3], r2 = 0;
                for (int k = 0; k < 3; ++k) { dr[k] = R(o, iat)[k] - o.pc_xyz[3 * j + k]; r2 += dr[k] * dr[k]; }
                double rr = 0.5 * (hard[s] + o.pc_gam[j]);
                rr = 1.0 / (rr * rr);
                const double yy = qsh[s] * o.pc_q[j] * std::pow(r2 + rr, -1.5);
                for (int k = 0; k < 3; ++k) { grad[3 * iat + k] -= dr[k] * yy; gpc[3 * j + k] += dr[k] * yy; }
            }
        }
        D["pcgradient"] = gpc;
        D["e_pcem"] = dvec(1, epcem);
    }
    {
        // getCoulombDerivsCluster + gemv, klopmanohno.F90:471-573, scf_module.F90:781-785
        int nth = o.nthreads;
        std::vector<dvec> gl(nth, dvec(3 * n, 0.0));
#pragma omp parallel num_threads(nth)
        {
#ifdef _OPENMP
            int tid = omp_get_thread_num();
#else
            int tid = 0;
#endif
#pragma omp for schedule(dynamic, 8)
            for (int iat = 0; iat < n; ++iat)
                for (int jat = 0; jat < iat; ++jat) {
                    double vec[3], r2 = 0;
                    for (int k = 0; k < 3; ++k) { vec[k] = R(o, jat)[k] - R(o, iat)[k]; r2 += vec[k] * vec[k]; }
                    double r1 = std::sqrt(r2);
                    for (int is = b.sh0[iat]; is < b.sh0[iat] + b.nsh[iat]; ++is)
                        for (int js = b.sh0[jat]; js < b.sh0[jat] + b.nsh[jat]; ++js) {
                            double gij = 0.5 * (hard[is] + hard[js]);
                            double g1 = 1.0 / (std::pow(r1, gExp) + std::pow(gij, -gExp));
                            double f = -std::pow(r1, gExp - 2.0) * g1 * std::pow(g1, 1.0 / gExp);
                            for (int k = 0; k < 3; ++k) {
                                double dG = vec[k] * f;
                                gl[tid][3 * iat + k] -= dG * qsh[is] * qsh[js];
                                gl[tid][3 * jat + k] += dG * qsh[is] * qsh[js];
                            }
                        }
                }
        }
        for (int t = 0; t < nth; ++t) for (int k = 0; k < 3 * n; ++k) grad[k] += gl[t][k];
    }
    o.timers[6] = now_s() - t0;
    t0 = now_s();

    // ---- properties (scf_module.F90:856-913)
    dvec wbo((size_t)n * n, 0.0);
    {
        bool fac_first = true;      // the pair counter of the first call only (open shell: alpha, then beta)
        auto wib = [&](const double* Pm, double fac) {
            dvec PS(nn);
            scipy_dgemm_("N", "N", &nao, &nao, &nao, &one, Pm, &nao, S.data(), &nao, &zero, PS.data(), &nao);
            for (int i = 0; i < n; ++i)
                for (int j = 0; j < i; ++j) {
                    double rab = 0, xsum = 0;
                    for (int k = 0; k < 3; ++k) { double t = R(o, i)[k] - R(o, j)[k]; rab += t * t; }
                    if (rab < 100.0 && fac_first) cnt_add(10, 1);
                    if (rab < 100.0)
                        for (int k = b.ao0_at[i]; k < b.ao0_at[i] + b.nao_at[i]; ++k)
                            for (int m = b.ao0_at[j]; m < b.ao0_at[j] + b.nao_at[j]; ++m)
                                xsum += PS[k + (size_t)nao * m] * PS[m + (size_t)nao * k];
                    wbo[i + (size_t)n * j] += fac * xsum; wbo[j + (size_t)n * i] += fac * xsum;
                }
        };
        if (nopen == 0) wib(P.data(), 1.0);
        else {
            dvec Pa(nn), Pb(nn);
            dmat(nao, focca.data(), H.data(), Pa.data());
            dmat(nao, foccb.data(), H.data(), Pb.data());
            wib(Pa.data(), 2.0); fac_first = false; wib(Pb.data(), 2.0);
        }
    }
    dvec dip(3, 0.0);
    for (int i = 0; i < n; ++i) for (int k = 0; k < 3; ++k) dip[k] += R(o, i)[k] * el_valence_z(o.Z[i]);
    for (int i = 0; i < nao; ++i) {
        for (int j = 0; j < i; ++j)
            for (int k = 0; k < 3; ++k) dip[k] -= 2.0 * P[j + (size_t)nao * i] * Dp[k + 3 * (i + (size_t)nao * j)];
        for (int k = 0; k < 3; ++k) dip[k] -= P[i + (size_t)nao * i] * Dp[k + 3 * (i + (size_t)nao * i)];
    }
    double energy = eel + ep + embd;
    double gnorm = 0;
    for (int k = 0; k < 3 * n; ++k) gnorm += grad[k] * grad[k];
    gnorm = std::sqrt(gnorm);

    D["energy"] = dvec(1, energy);
    D["energies"] = dvec{energy, eel + embd, ep, ees, eaes, epol, ed + embd, egap, gnorm, embd, eel, ed};
    D["gradient"] = grad;
    D["q"] = q; D["qsh"] = qsh; D["dipm"] = dipm; D["qp"] = qp;
    D["wbo"] = wbo; D["dipole"] = dip; D["emo"] = emo; D["focc"] = focc;
    D["S"] = S; D["dpint"] = Dp; D["qpint"] = Qp; D["H0"] = H0; D["H0_noovlp"] = H0n;
    D["P"] = P; D["C"] = H; D["jmat"] = jmat; D["gab3"] = gab3; D["gab5"] = gab5; D["gw"] = gw; D["dispmat"] = dispmat;
    D["selfenergy"] = selfE; D["radcn"] = radcn; D["shellshift"] = shellShift; D["vs"] = vs; D["vd"] = vd; D["vq"] = vq;
    D["zsh"] = zsh;
    D["timers"] = dvec(o.timers, o.timers + 8);
    {
        dvec dc(12);
        for (int k = 0; k < 12; ++k) dc[k] = (double)g_cnt[k];
        D["decisions"] = dc;
    }
    o.iv["mdlst"] = mdlst; o.iv["mqlst"] = mqlst;
    o.iv["info"] = ivec{jter, converged ? 1 : 0, nao, nsh, ndp, nqp, nel, nopen, ihomo, ndim, nbr};
    o.timers[0] += now_s() - t0;
    return converged ? 0 : 1;
}

}  // namespace

// =====================================================================================
// C interface for ctypes (tests/, bench.py cpu_baseline)
// =====================================================================================
extern "C" {
void* orc_create(int nat, const int* Z, const double* xyz, double charge, int uhf) {
    Oracle* o = new Oracle();
    o->nat = nat;
    o->Z.assign(Z, Z + nat);
    o->xyz.assign(xyz, xyz + 3 * nat);
    o->charge = charge;
    o->uhf = uhf;
#ifdef _OPENMP
    o->nthreads = omp_get_max_threads();
#endif
    return o;
}
void orc_destroy(void* h) { delete (Oracle*)h; }
void orc_set_params(void* h, double acc, double etemp, int maxiter) {
    Oracle* o = (Oracle*)h;
    o->acc = acc; o->etemp = etemp; o->maxiter = maxiter;
}
void orc_set_threads(void* h, int nthreads) {
    Oracle* o = (Oracle*)h;
#ifdef _OPENMP
    omp_set_num_threads(nthreads);
    o->nthreads = nthreads;
#else
    o->nthreads = 1;
#endif
    scipy_openblas_set_num_threads(nthreads);
}
void orc_set_guess(void* h, int mode) { ((Oracle*)h)->guess = mode; }
// external point charges: numbers -> hardness as in setExternalCharges_api (src/api/calculator.f90:530-535); npc = 0 releases
void orc_set_pcem(void* h, int npc, const int* numbers, const double* q, const double* xyz) {
    Oracle* o = (Oracle*)h;
    o->pc_q.assign(q, q + npc);
    o->pc_xyz.assign(xyz, xyz + 3 * npc);
    o->pc_gam.resize(npc);
    for (int i = 0; i < npc; ++i) o->pc_gam[i] = gfn2_chemical_hardness[numbers[i] - 1];
}
void orc_set_xyz(void* h, const double* xyz) { Oracle* o = (Oracle*)h; o->xyz.assign(xyz, xyz + 3 * o->nat); }
// implicit solvation (xtb_setSolvent, src/api/calculator.f90:341-445): solvent name as given by the caller (normalised
// here as normalizeSolventName does, src/solv/model.f90:211-246), alpb 0/1, kernel 1 Still / 2 P16; name == NULL releases
void orc_set_solvent(void* h, const char* name, int alpb, int kernel, int state, double temperature, double ion_strength,
                     double ion_rad) {
    Oracle* o = (Oracle*)h;
    if (!name) { o->solv_on = false; return; }
    std::string s(name);
    for (auto& c : s) c = (char)std::tolower((unsigned char)c);
    static const char* alias[][2] = {{"methylcyanide", "acetonitrile"}, {"methylenechloride", "ch2cl2"}, {"dichloromethane", "ch2cl2"},
        {"dcm", "ch2cl2"}, {"chloroform", "chcl3"}, {"trichloromethane", "chcl3"}, {"tcm", "chcl3"}, {"carbondisulfide", "cs2"},
        {"diethylether", "ether"}, {"dimethylformamide", "dmf"}, {"dimethylsulfoxide", "dmso"}, {"furan", "furane"},
        {"nhexan", "hexane"}, {"n-hexan", "hexane"}, {"nhexane", "hexane"}, {"n-hexane", "hexane"}, {"tetrahydrofuran", "thf"},
        {"h2o", "water"}};
    for (auto& a : alias) if (s == a[0]) s = a[1];
    o->solv_on = true; o->solv_name = s; o->solv_alpb = alpb != 0; o->solv_kernel = kernel; o->solv_state = state;
    o->solv_temp = temperature; o->solv_ionst = ion_strength; o->solv_ionrad = ion_rad;
}
void orc_set_restart(void* h, const double* qsh, int nsh, const double* dipm, const double* qp) {
    Oracle* o = (Oracle*)h;
    o->d["restart_qsh"].assign(qsh, qsh + nsh);
    o->d["restart_dipm"].assign(dipm, dipm + 3 * o->nat);
    o->d["restart_qp"].assign(qp, qp + 6 * o->nat);
}
int orc_singlepoint(void* h) {
    Oracle* o = (Oracle*)h;
    o->d.erase("iter_energy"); o->d.erase("iter_rmsq");
    return singlepoint(*o);
}
int orc_len(void* h, const char* name) {
    Oracle* o = (Oracle*)h;
    auto it = o->d.find(name);
    if (it != o->d.end()) return (int)it->second.size();
    auto jt = o->iv.find(name);
    if (jt != o->iv.end()) return (int)jt->second.size();
    return -1;
}
long orc_len64(void* h, const char* name) {
    Oracle* o = (Oracle*)h;
    auto it = o->d.find(name);
    if (it != o->d.end()) return (long)it->second.size();
    auto jt = o->iv.find(name);
    if (jt != o->iv.end()) return (long)jt->second.size();
    return -1;
}
int orc_get(void* h, const char* name, double* buf, long maxlen) {
    Oracle* o = (Oracle*)h;
    auto it = o->d.find(name);
    if (it == o->d.end()) return -1;
    long n = std::min<long>(maxlen, (long)it->second.size());
    std::memcpy(buf, it->second.data(), n * sizeof(double));
    return 0;
}
int orc_geti(void* h, const char* name, int* buf, long maxlen) {
    Oracle* o = (Oracle*)h;
    auto it = o->iv.find(name);
    if (it == o->iv.end()) return -1;
    long n = std::min<long>(maxlen, (long)it->second.size());
    std::memcpy(buf, it->second.data(), n * sizeof(int));
    return 0;
}
// standalone building blocks for component known-answer tests
// (test/unit/test_coulomb.f90 gfn2 cluster, test_repulsion.f90, test_eeq.f90)
void orc_cn_erf(void* h, double* cn) { Oracle* o = (Oracle*)h; dvec c; ncoord_erf_d4(*o, false, c); std::memcpy(cn, c.data(), c.size() * 8); }
int orc_eeq(void* h, double* q) {
    Oracle* o = (Oracle*)h; dvec c;
    if (!eeq_charges(*o, c)) return 1;
    std::memcpy(q, c.data(), c.size() * 8);
    return 0;
}
}
