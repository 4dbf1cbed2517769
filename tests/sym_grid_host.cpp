// Host harness (test infrastructure): walks the tiles of the symmetric trailing-matrix product of k_sytrd_panel
// (xtb_b200/csrc/kernels_eigh.cuh) on the CPU with the kernel's own index arithmetic (xtb_b200/csrc/sytrd_sym_grid.h):
// contiguous tile ranges per CTA, column sums per tile, row sums per CTA segment of a row chunk, and the gather that
// rebuilds y = A v from them.  Built by tests/test_sym_grid.py with g++.
#include <algorithm>
#include <cmath>
#include <vector>

#include "../xtb_b200/csrc/sytrd_sym_grid.h"

using namespace xtbb200;

// A: n x n column-major (only the lower triangle is read), jj: first trailing index, v: n entries (entries < jj ignored).
// y[r], r >= jj.  stats: [0] tiles, [1] largest row-sum slot + 1, [2] largest tile index + 1, [3] CTAs with work,
// [4] largest number of tiles of one CTA, [5] largest number of terms of a row of y while the trailing order is at least 1800 (the resident kernel takes over at about 1860).  Returns 0, or a negative code when a slot would be written twice.
extern "C" int sym_product_host(int n, int jj, int G, const double* A, const double* vin, double* y, long long* stats) {
    const int npad = (n + kSyRH - 1) / kSyRH * kSyRH;
    std::vector<double> v(npad, 0.0);
    for (int r = jj; r < n; ++r) v[r] = vin[r];
    const SymGrid g = sym_grid(jj, n);
    const size_t ncol = (size_t)g.nS * g.nR * kSyCW, nrow = (size_t)(g.nR + G + 1) * kSyRH;
    std::vector<double> colbuf(ncol, NAN), rowbuf(nrow, NAN);
    std::vector<char> colw(ncol / kSyCW, 0), roww(nrow / kSyRH, 0);
    long long maxrow = 0, maxtile = 0, busy = 0, most = 0;
    constexpr int NG = 8, kGC = kSyCW / NG;
    for (int cta = 0; cta < G; ++cta) {
        int tcur = sym_start(g, cta, G);
        const int tend = sym_start(g, cta + 1, G);
        if (tcur >= tend) continue;
        ++busy;
        most = std::max<long long>(most, tend - tcur);
        int cR, cS;
        sym_locate(g, tcur, cR, cS);
        std::vector<double> rowacc(NG * kSyRH, 0.0);
        for (; tcur < tend; ++tcur) {
            const int r0 = cR * kSyRH, c0 = cS * kSyCW;
            if (cS < g.Smin || cS > std::min(4 * cR + 3, g.nS - 1) || cR >= g.nR) return -3;
            if (sym_pre(g, cR) + cS - g.Smin != tcur) return -4;
            if (sym_owner(g, tcur, G) != cta) return -5;
            const bool diag = (cS >> 2) == cR;
            for (int w = 0; w < kSyCW; ++w) {   // column sums
                const int c = c0 + w;
                if (c >= n) continue;
                double cs = 0.0;
                for (int r = r0; r < std::min(r0 + kSyRH, n); ++r) {
                    double x = A[r + (size_t)n * c];
                    if (diag && r <= c) x = 0.0;
                    cs += x * v[r];
                }
                if (colw[tcur]++ >= kSyCW) return -1;
                colbuf[(size_t)tcur * kSyCW + w] = cs;
            }
            maxtile = std::max<long long>(maxtile, tcur + 1);
            for (int trow = 0; trow < kSyRH; ++trow)   // row sums
                for (int tg = 0; tg < NG; ++tg)
                    for (int u = 0; u < kGC; ++u) {
                        const int cl = tg * kGC + u, r = r0 + trow, c = c0 + cl;
                        if (r >= n || c >= n) continue;
                        double x = A[r + (size_t)n * c];
                        if (diag && r < c) x = 0.0;
                        rowacc[tg * kSyRH + trow] += x * v[c];
                    }
            const int Rdone = cR;
            if (++cS > std::min(4 * cR + 3, g.nS - 1)) { ++cR; cS = g.Smin; }
            if (cR != Rdone || tcur + 1 == tend) {
                const size_t slot = (size_t)(Rdone - g.Rmin + cta);
                if (slot >= roww.size()) return -6;
                if (roww[slot]++) return -2;
                for (int t = 0; t < kSyRH; ++t) {
                    double s = 0.0;
                    for (int u = 0; u < NG; ++u) s += rowacc[u * kSyRH + t];
                    rowbuf[slot * kSyRH + t] = s;
                }
                std::fill(rowacc.begin(), rowacc.end(), 0.0);
                maxrow = std::max<long long>(maxrow, (long long)slot + 1);
            }
        }
    }
    for (int r = jj; r < n; ++r) {
        double s = 0.0;
        for (int ks = 0; ks < 16; ++ks) {
            const double* q[kSyGatherTerms];
            if (!sym_gather_ptrs(g, G, r, ks, 16, rowbuf.data(), colbuf.data(), q)) return -7;
            s += sym_gather_sum(q);
        }
        y[r] = s;
    }
    stats[5] = sym_gather_max_terms(n, G, n - 1800);
    stats[0] = g.T; stats[1] = maxrow; stats[2] = maxtile; stats[3] = busy; stats[4] = most;
    return 0;
}
