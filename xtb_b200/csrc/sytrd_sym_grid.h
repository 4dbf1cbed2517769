// Tile grid of the symmetric trailing-matrix product of the Householder tridiagonalisation (k_sytrd_panel in
// kernels_eigh.cuh): pure index arithmetic shared by the kernel, the host launch code and the CPU test
// tests/sym_grid_host.cpp (which walks the tiles exactly as the kernel does).
#pragma once
#include <cstddef>
#if defined(__CUDACC__)
#define XTBB200_HD __host__ __device__ __forceinline__
#else
#define XTBB200_HD inline
#endif
#if defined(__CUDA_ARCH__)
#define XTBB200_LDCG(p) __ldcg(p)
#else
#define XTBB200_LDCG(p) (*(p))
#endif

namespace xtbb200 {

// ---- tile grid of the symmetric trailing-matrix product ------------------------------------------------------------
// The lower triangle of the trailing matrix (first index jj) is cut along ABSOLUTE row chunks of kSyRH rows and column
// strips of kSyCW columns (so that every tile line starts 16-byte aligned for the bulk copy whatever jj is); entries left
// of / above jj are switched off by zeros in v.  Tiles are numbered row chunk by row chunk (strips ascending inside a
// chunk) and dealt to the CTAs in contiguous ranges of equal length.
constexpr int kSyCW = 32, kSyRH = 128, kSyTile = kSyCW * kSyRH, kSyGroups = 4, kSyStagesMax = kSyGroups;
struct SymGrid { int Smin, Rmin, nS, nR, T; };   // T: number of tiles (< 2^31 / grid for every order the solver takes)
XTBB200_HD int sym_pre(const SymGrid& g, int R) {   // tiles before row chunk R (R < nR)
    return 2 * (R * (R - 1) - g.Rmin * (g.Rmin - 1)) + (4 - g.Smin) * (R - g.Rmin);
}
XTBB200_HD int sym_cnt(const SymGrid& g, int R) { return (4 * R + 3 < g.nS - 1 ? 4 * R + 3 : g.nS - 1) - g.Smin + 1; }
XTBB200_HD SymGrid sym_grid(int jj, int n) {
    SymGrid g;
    g.Smin = jj / kSyCW; g.Rmin = jj / kSyRH; g.nS = (n + kSyCW - 1) / kSyCW; g.nR = (n + kSyRH - 1) / kSyRH;
    g.T = sym_pre(g, g.nR - 1) + sym_cnt(g, g.nR - 1);
    return g;
}
XTBB200_HD void sym_next(const SymGrid& g, int& R, int& S) {   // the tile after (R, S)
    if (++S > (4 * R + 3 < g.nS - 1 ? 4 * R + 3 : g.nS - 1)) { ++R; S = g.Smin; }
}
XTBB200_HD int sym_start(const SymGrid& g, int p, int G) { return (int)((unsigned)g.T * (unsigned)p / (unsigned)G); }   // first tile of CTA p
XTBB200_HD int sym_owner(const SymGrid& g, int t, int G) { return (int)(((unsigned)(t + 1) * (unsigned)G - 1u) / (unsigned)g.T); }
XTBB200_HD void sym_locate(const SymGrid& g, int t, int& R, int& S) {
    int lo = g.Rmin, hi = g.nR - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (sym_pre(g, mid) <= t) lo = mid; else hi = mid - 1;
    }
    R = lo; S = g.Smin + (t - sym_pre(g, lo));
}
// y[r] of the product whose tile grid is g = the row sums of the CTA segments of r's row chunk, then the column sums of
// the tiles of r's strip.  Slice ks of nsl takes the terms ks, ks + nsl, ...: their addresses go to q[0 .. kSyGatherTerms)
// (null: no such term), so that the kernel can work them out before the grid barrier and only has to load afterwards; the
// caller adds the slices in order (fixed order of summation).  Returns false when a slice has more terms than q holds.
constexpr int kSyGatherTerms = 6;
XTBB200_HD bool sym_gather_ptrs(const SymGrid& g, int G, int r, int ks, int nsl, const double* rowbuf, const double* colbuf,
                                const double* (&q)[kSyGatherTerms]) {
    const int R = r / kSyRH, S = r / kSyCW;
    const int tb = sym_pre(g, R);
    const int plo = sym_owner(g, tb, G), phi = sym_owner(g, tb + sym_cnt(g, R) - 1, G);
    const int nseg = phi - plo + 1, nit = nseg + g.nR - R;
    const bool sparse = g.T < G;   // fewer tiles than CTAs: CTAs between two owners may have no tile at all
    const double* rowp = rowbuf + (size_t)(R - g.Rmin + plo) * kSyRH + (r - R * kSyRH);
    const double* colp = colbuf + (size_t)(S - g.Smin) * kSyCW + (r - S * kSyCW);
#pragma unroll
    for (int u = 0; u < kSyGatherTerms; ++u) {
        const int i = ks + u * nsl;
        q[u] = nullptr;
        if (i < nseg) {
            if (!sparse || sym_start(g, plo + i, G) < sym_start(g, plo + i + 1, G)) q[u] = rowp + (size_t)i * kSyRH;
        } else if (i < nit) q[u] = colp + (size_t)sym_pre(g, R + (i - nseg)) * kSyCW;
    }
    return nit <= kSyGatherTerms * nsl;
}
XTBB200_HD double sym_gather_sum(const double* const (&q)[kSyGatherTerms]) {
    double x[kSyGatherTerms], s = 0.0;
#pragma unroll
    for (int u = 0; u < kSyGatherTerms; ++u) x[u] = q[u] ? XTBB200_LDCG(q[u]) : 0.0;
#pragma unroll
    for (int u = 0; u < kSyGatherTerms; ++u) s += x[u];
    return s;
}
XTBB200_HD double sym_gather(const SymGrid& g, int G, int r, int ks, int nsl, const double* rowbuf, const double* colbuf) {
    const double* q[kSyGatherTerms];
    sym_gather_ptrs(g, G, r, ks, nsl, rowbuf, colbuf, q);
    return sym_gather_sum(q);
}
// largest number of terms of any row for order n on G CTAs while the first trailing index is at most jjmax (host: the
// symmetric form is used only when every slice of kSySlices threads can hold its share)
inline int sym_gather_max_terms(int n, int G, int jjmax) {
    int worst = 0;
    for (int jj = 1; jj < n && jj <= jjmax; ++jj) {
        const SymGrid g = sym_grid(jj, n);
        for (int R = g.Rmin; R < g.nR; ++R) {
            const int tb = sym_pre(g, R);
            const int nseg = sym_owner(g, tb + sym_cnt(g, R) - 1, G) - sym_owner(g, tb, G) + 1;
            if (nseg + g.nR - R > worst) worst = nseg + g.nR - R;
        }
    }
    return worst;
}

}  // namespace xtbb200
