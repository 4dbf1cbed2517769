// Host harness (test infrastructure): runs the divide-and-conquer tridiagonal eigensolver of kernels_eigh.cuh on the CPU,
// level by level with the same data flow as the CUDA kernels and the same scalar routines (xtb_b200/csrc/eig_dc_core.h),
// so that the tearing tree, deflation, secular solver and eigenvector update can be checked without a GPU.
// Built by tests/test_eig_core.py with g++.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "../xtb_b200/csrc/eig_dc_core.h"

using namespace xtbb200::eigcore;

extern "C" int dc_tridiag_host(int n, const double* din, const double* ein, double* lam_out, double* Qout, int* stats) {
    std::vector<double> Qa((size_t)n * n, 0.0), Qb((size_t)n * n, 0.0), U((size_t)n * n, 0.0);
    std::vector<double> lam(n), lam2(n), z(n), zhat(n), dk(n), zk(n), rc(n), rs(n);
    std::vector<int> perm(n), cols(n), ra(n), rb(n), ctype(n), pos(n);
    std::vector<double> Qg((size_t)n * n, 0.0), Ug((size_t)n * n, 0.0);
    int fails = 0, total_k = 0, total_rot = 0;
    if (n == 1) { lam_out[0] = din[0]; Qout[0] = 1.0; return 0; }
    const int D = tree_depth(n);
    // leaves: torn diagonal, 1x1 or 2x2 blocks
    for (int t = 0; t < (1 << D); ++t) {
        const int lo = tree_bound(n, D, t), hi = tree_bound(n, D, t + 1);
        double a = din[lo] - (lo > 0 ? fabs(ein[lo - 1]) : 0.0);
        if (hi - lo == 1) {
            a -= (hi < n ? fabs(ein[hi - 1]) : 0.0);
            lam[lo] = a;
            Qa[lo + (size_t)n * lo] = 1.0;
        } else {
            const double c = din[lo + 1] - (hi < n ? fabs(ein[hi - 1]) : 0.0), b = ein[lo];
            double rt1, rt2, cs, sn;
            sym2x2(a, b, c, rt1, rt2, cs, sn);
            lam[lo] = rt1; lam[lo + 1] = rt2;
            Qa[lo + (size_t)n * lo] = cs; Qa[lo + 1 + (size_t)n * lo] = sn;
            Qa[lo + (size_t)n * (lo + 1)] = -sn; Qa[lo + 1 + (size_t)n * (lo + 1)] = cs;
        }
    }
    double* Qo = Qa.data();
    double* Qn = Qb.data();
    for (int depth = D - 1; depth >= 0; --depth) {
        for (int t = 0; t < (1 << depth); ++t) {
            const int lo = tree_bound(n, depth, t), hi = tree_bound(n, depth, t + 1), mid = tree_bound(n, depth + 1, 2 * t + 1);
            const int m = hi - lo;
            const double rho_s = ein[mid - 1], rho = 2.0 * fabs(rho_s), sg = rho_s < 0 ? -1.0 : 1.0;
            const double isq2 = 1.0 / sqrt(2.0);
            for (int i = 0; i < m; ++i) {
                const int c = lo + i;
                z[i] = (c < mid ? Qo[(mid - 1) + (size_t)n * c] : sg * Qo[mid + (size_t)n * c]) * isq2;
            }
            // rank sort
            for (int i = 0; i < m; ++i) {
                int r = 0;
                for (int j2 = 0; j2 < m; ++j2) r += (lam[lo + j2] < lam[lo + i] || (lam[lo + j2] == lam[lo + i] && j2 < i)) ? 1 : 0;
                perm[r] = i;
            }
            int nrot = 0;
            const int n1 = mid - lo;
            const int k = deflate_scan(m, n1, perm.data(), &lam[lo], z.data(), rho, cols.data(), ctype.data(), ra.data(),
                                       rb.data(), rc.data(), rs.data(), &nrot);
            int cnt[3];
            group_columns(k, cols.data(), ctype.data(), pos.data(), cnt);
            total_k += k; total_rot += nrot;
            for (int r = 0; r < nrot; ++r)
                for (int row = lo; row < hi; ++row) {
                    double& x = Qo[row + (size_t)n * (lo + ra[r])];
                    double& y = Qo[row + (size_t)n * (lo + rb[r])];
                    const double xn = rc[r] * x + rs[r] * y, yn = rc[r] * y - rs[r] * x;
                    x = xn; y = yn;
                }
            for (int i = 0; i < k; ++i) { dk[i] = lam[lo + cols[i]]; zk[i] = z[cols[i]]; }
            // secular equation: delta(i, j) stored in U[(lo+i) + n (lo+j)]
            for (int j2 = 0; j2 < k; ++j2)
                fails += secular_root(k, j2, dk.data(), zk.data(), rho, &U[lo + (size_t)n * (lo + j2)], 1, &lam2[lo + j2], SeqPar());
            for (int i = 0; i < k; ++i) {
                double w = U[(lo + i) + (size_t)n * (lo + i)];
                for (int j2 = 0; j2 < k; ++j2)
                    if (j2 != i) w *= U[(lo + i) + (size_t)n * (lo + j2)] / (dk[i] - dk[j2]);
                zhat[i] = copysign(sqrt(fabs(w)), zk[i]);
            }
            for (int j2 = 0; j2 < k; ++j2) {
                double s = 0;
                for (int i = 0; i < k; ++i) { double& u = U[(lo + i) + (size_t)n * (lo + j2)]; u = zhat[i] / u; s += u * u; }
                s = 1.0 / sqrt(s);
                for (int i = 0; i < k; ++i) Ug[(lo + pos[i]) + (size_t)n * (lo + j2)] = U[(lo + i) + (size_t)n * (lo + j2)] * s;
            }
            // grouped copies of the non-deflated columns, then two products: top rows use the slots [0, c0 + c1),
            // bottom rows the slots [c0, k); deflated columns copied behind
            for (int i = 0; i < k; ++i)
                for (int row = lo; row < hi; ++row) Qg[row + (size_t)n * (lo + pos[i])] = Qo[row + (size_t)n * (lo + cols[i])];
            for (int j2 = 0; j2 < k; ++j2)
                for (int row = lo; row < hi; ++row) {
                    const int s0 = row < mid ? 0 : cnt[0], s1 = row < mid ? cnt[0] + cnt[1] : k;
                    double s = 0;
                    for (int i = s0; i < s1; ++i) s += Qg[row + (size_t)n * (lo + i)] * Ug[(lo + i) + (size_t)n * (lo + j2)];
                    Qn[row + (size_t)n * (lo + j2)] = s;
                }
            for (int p = k; p < m; ++p) {
                for (int row = lo; row < hi; ++row) Qn[row + (size_t)n * (lo + p)] = Qo[row + (size_t)n * (lo + cols[p])];
                lam2[lo + p] = lam[lo + cols[p]];
            }
        }
        std::swap(Qo, Qn);
        lam.swap(lam2);
    }
    // final ascending order
    for (int i = 0; i < n; ++i) {
        int r = 0;
        for (int j2 = 0; j2 < n; ++j2) r += (lam[j2] < lam[i] || (lam[j2] == lam[i] && j2 < i)) ? 1 : 0;
        lam_out[r] = lam[i];
        std::memcpy(Qout + (size_t)n * r, Qo + (size_t)n * i, sizeof(double) * n);
    }
    if (stats) { stats[0] = fails; stats[1] = total_k; stats[2] = total_rot; }
    return fails;
}
