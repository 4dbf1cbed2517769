// Scalar building blocks of the divide-and-conquer eigensolver for the symmetric tridiagonal matrix (Cuppen's rank-one
// tearing with Gu/Eisenstat's re-computed update vector), shared between the CUDA kernels (kernels_eigh.cuh) and the
// host harness that checks them on the CPU (tests/eig_core_host.cpp).
//
// Replaces the tridiagonal stage of dsyevd (dstedc -> dlaed0/1/2/3/4), called by the reference through
// TEigenSolver%fact_solve (src/mctc/lapack/eigensolve.F90:218-246).  LAPACK itself is not part of /root/reference; the
// routines below restate the published algorithm (Cuppen 1981; Gu & Eisenstat, SIAM J. Matrix Anal. Appl. 16 (1995) 172;
// Li, LAPACK working note 89 for the two-pole rational interpolation of the secular equation).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define EIG_HD __host__ __device__ __forceinline__
#else
#define EIG_HD inline
#endif

namespace xtbb200 {
namespace eigcore {

constexpr double kEps = 1.1102230246251565e-16;   // unit roundoff 2^-53

// Boundaries of the binary tearing tree: node t at depth k covers rows [tree_bound(n,k,t), tree_bound(n,k,t+1)).
EIG_HD int tree_bound(int n, int depth, int t) { return (int)(((long long)t * n) >> depth); }
// depth of the leaves: the largest D with n / 2^D >= 1, so that every leaf has one or two rows
EIG_HD int tree_depth(int n) {
    int D = 0;
    while ((n >> (D + 1)) >= 1) ++D;
    return D;
}

// Eigen-decomposition of [[a, b], [b, c]]: rt1 has the larger absolute value, (cs1, sn1) is its unit eigenvector and
// (-sn1, cs1) the eigenvector of rt2.
EIG_HD void sym2x2(double a, double b, double c, double& rt1, double& rt2, double& cs1, double& sn1) {
    const double sm = a + c, df = a - c, adf = fabs(df), tb = b + b, ab = fabs(tb);
    double acmx = c, acmn = a;
    if (fabs(a) > fabs(c)) { acmx = a; acmn = c; }
    double rt;
    if (adf > ab) { const double q = ab / adf; rt = adf * sqrt(1.0 + q * q); }
    else if (adf < ab) { const double q = adf / ab; rt = ab * sqrt(1.0 + q * q); }
    else rt = ab * sqrt(2.0);
    int sgn1;
    if (sm < 0.0) { rt1 = 0.5 * (sm - rt); sgn1 = -1; rt2 = (acmx / rt1) * acmn - (b / rt1) * b; }
    else if (sm > 0.0) { rt1 = 0.5 * (sm + rt); sgn1 = 1; rt2 = (acmx / rt1) * acmn - (b / rt1) * b; }
    else { rt1 = 0.5 * rt; rt2 = -0.5 * rt; sgn1 = 1; }
    int sgn2;
    double cs;
    if (df >= 0.0) { cs = df + rt; sgn2 = 1; } else { cs = df - rt; sgn2 = -1; }
    if (fabs(cs) > ab) { const double ct = -tb / cs; sn1 = 1.0 / sqrt(1.0 + ct * ct); cs1 = ct * sn1; }
    else if (ab == 0.0) { cs1 = 1.0; sn1 = 0.0; }
    else { const double tn = -cs / tb; cs1 = 1.0 / sqrt(1.0 + tn * tn); sn1 = tn * cs1; }
    if (sgn1 == sgn2) { const double tn = cs1; cs1 = -sn1; sn1 = tn; }
}

// Deflation scan of one merge (sequential by nature: a rotation that removes a close pair changes the candidate the
// next entry is compared with).  d, z: the m eigenvalues / update-vector entries of the node in column order (modified
// in place), perm: column indices in ascending order of d, rho > 0.
// Output: cols[0..k) non-deflated columns in ascending order of d, cols[k..m) deflated columns; rotations
// (ra[t], rb[t], rc[t], rs[t]), t < nrot, to be applied in this order to the eigenvector columns:
//   x = Q[:, ra], y = Q[:, rb]  ->  Q[:, ra] = c x + s y,  Q[:, rb] = c y - s x.
// ctype[c] (work array, m entries): 1 = column of the left child (non-zero in the top n1 rows only), 3 = column of the
// right child (bottom rows only), 2 = dense after a rotation that mixed the two kinds; valid for cols[0..k) on return.
// Returns k.
EIG_HD int deflate_scan(int m, int n1, const int* perm, double* d, double* z, double rho, int* cols, int* ctype, int* ra,
                        int* rb, double* rc, double* rs, int* nrot_out) {
    for (int i = 0; i < m; ++i) ctype[i] = i < n1 ? 1 : 3;
    double dmax = 0.0, zmax = 0.0;
    for (int i = 0; i < m; ++i) { dmax = fmax(dmax, fabs(d[i])); zmax = fmax(zmax, fabs(z[i])); }
    const double tol = 8.0 * kEps * fmax(dmax, zmax);
    int k = 0, ndf = 0, nrot = 0;
    if (rho * zmax <= tol) {
        for (int p = 0; p < m; ++p) cols[p] = perm[p];
        *nrot_out = 0;
        return 0;
    }
    // deflated columns are collected from the end of cols[] backwards and reversed afterwards
    int pj = -1;
    for (int p = 0; p < m; ++p) {
        const int nj = perm[p];
        if (rho * fabs(z[nj]) <= tol) { cols[m - 1 - ndf++] = nj; continue; }
        if (pj < 0) { pj = nj; continue; }
        double s = z[pj], c = z[nj];
        const double t = d[nj] - d[pj];
        // |t c s| / (c^2 + s^2) <= tol is the test of dlaed2 (there with c, s already divided by hypot(c, s)); in this
        // form the square root and the divisions are only paid for the pairs that do deflate
        if (fabs(t * c * s) <= tol * (c * c + s * s)) {
            const double tau = hypot(c, s);
            c /= tau;
            s = -s / tau;
            z[nj] = tau;
            z[pj] = 0.0;
            ra[nrot] = pj; rb[nrot] = nj; rc[nrot] = c; rs[nrot] = s; ++nrot;
            if (ctype[nj] != ctype[pj]) ctype[nj] = 2;
            const double tt = d[pj] * c * c + d[nj] * s * s;
            d[nj] = d[pj] * s * s + d[nj] * c * c;
            d[pj] = tt;
            cols[m - 1 - ndf++] = pj;
        } else {
            cols[k++] = pj;
        }
        pj = nj;
    }
    if (pj >= 0) cols[k++] = pj;
    for (int a = k, b = m - 1; a < b; ++a, --b) { const int t = cols[a]; cols[a] = cols[b]; cols[b] = t; }
    *nrot_out = nrot;
    return k;
}

// Grouping of the k non-deflated columns by kind (left-only, dense, right-only; stable inside a kind): pos[i] = slot
// of the i-th column (ascending order of d) in the grouped order, cnt[0..3) = columns per kind.  With the update matrix
// stored in grouped row order, the top n1 rows of the new eigenvectors need the slots [0, cnt0 + cnt1) only and the
// bottom rows the slots [cnt0, k): two dense products instead of one of twice the size.
EIG_HD void group_columns(int k, const int* cols, const int* ctype, int* pos, int* cnt) {
    int c[3] = {0, 0, 0};
    for (int i = 0; i < k; ++i) c[ctype[cols[i]] - 1]++;
    int o[3] = {0, c[0], c[0] + c[1]};
    for (int i = 0; i < k; ++i) pos[i] = o[ctype[cols[i]] - 1]++;
    cnt[0] = c[0]; cnt[1] = c[1]; cnt[2] = c[2];
}

// how the poles of one secular equation are shared out: one thread (host, SeqPar) or the lanes of a warp (kernels)
struct SeqPar {
    EIG_HD int first() const { return 0; }
    EIG_HD int step() const { return 1; }
    EIG_HD double sum(double x) const { return x; }
};

// Root j (0-based, ascending) of the secular equation 1 + rho sum_i z_i^2 / (d_i - lambda) = 0 with d strictly
// ascending and rho > 0.  The iteration runs in the shifted variable tau = lambda - origin, origin = the closer pole,
// so that delta_i = d_i - lambda_j = (d_i - origin) - tau is obtained without cancellation; these differences are
// what the eigenvector formula needs.  delta is written with the given stride.  Returns 0 when converged (secular
// function below its rounding-error bound or bracket collapsed), 1 when the iteration limit was hit.
template <class Par>
EIG_HD int secular_root(int k, int j, const double* dk, const double* zk, double rho, double* delta, long long stride,
                        double* lam_out, Par par) {
    if (k == 1) {
        const double z0 = zk[0], t = rho * z0 * z0;
        if (par.first() == 0) delta[0] = -t;
        *lam_out = dk[0] + t;
        return 0;
    }
    const bool last = (j == k - 1);
    const double dj = dk[j];
    double origin, lo, hi, tau;
    int isplit, i1, i2;
    if (!last) {
        const double g = dk[j + 1] - dj, half = 0.5 * g;
        double f = 0.0;
        for (int i = par.first(); i < k; i += par.step()) { const double zi = zk[i]; f += zi * (zi / ((dk[i] - dj) - half)); }
        f = 1.0 + rho * par.sum(f);
        if (f >= 0.0) { origin = dj; lo = 0.0; hi = half; tau = half; }
        else { origin = dk[j + 1]; lo = -half; hi = 0.0; tau = -half; }
        isplit = j; i1 = j; i2 = j + 1;
    } else {
        double zz = 0.0;
        for (int i = par.first(); i < k; i += par.step()) { const double zi = zk[i]; zz += zi * zi; }
        zz = par.sum(zz);
        origin = dj; lo = 0.0; hi = rho * zz; tau = 0.5 * hi;
        isplit = k - 2; i1 = k - 2; i2 = k - 1;
    }
    const double o1 = dk[i1] - origin, o2 = dk[i2] - origin;
    int status = 1;
    for (int it = 0; it < 120; ++it) {
        double psi = 0.0, dpsi = 0.0, phi = 0.0, dphi = 0.0;
        for (int i = par.first(); i < k; i += par.step()) {
            const double zi = zk[i], t = zi / ((dk[i] - origin) - tau);
            if (i <= isplit) { psi += zi * t; dpsi += t * t; }
            else { phi += zi * t; dphi += t * t; }
        }
        psi = rho * par.sum(psi); dpsi = rho * par.sum(dpsi); phi = rho * par.sum(phi); dphi = rho * par.sum(dphi);
        const double f = 1.0 + psi + phi;
        const double err = 8.0 * (fabs(psi) + fabs(phi)) + 1.0 + fabs(tau) * (dpsi + dphi);
        if (fabs(f) <= kEps * err) { status = 0; break; }
        if (f < 0.0) lo = tau; else hi = tau;
        if (hi - lo <= 2.0 * kEps * fmax(fabs(lo), fabs(hi))) { status = 0; break; }
        double eta = 0.0;
        bool ok = false;
        if (it < 40 || (it & 1)) {   // two-pole rational interpolation ("middle way"); late iterations alternate with bisection
            const double D1 = o1 - tau, D2 = o2 - tau;
            const double c = f - D1 * dpsi - D2 * dphi;
            const double a = (D1 + D2) * f - D1 * D2 * (dpsi + dphi);
            const double b = D1 * D2 * f;
            double r1, r2;
            bool h1 = false, h2 = false;
            if (c == 0.0) {
                if (a != 0.0) { r1 = b / a; h1 = true; }
            } else {
                double disc = a * a - 4.0 * b * c;
                if (disc < 0.0) disc = 0.0;
                const double sq = sqrt(disc), q = (a >= 0.0) ? 0.5 * (a + sq) : 0.5 * (a - sq);
                r1 = q / c; h1 = true;
                if (q != 0.0) { r2 = b / q; h2 = true; }
            }
            if (h1) { const double t = tau + r1; if (t > lo && t < hi) { eta = r1; ok = true; } }
            if (h2) { const double t = tau + r2; if (t > lo && t < hi && (!ok || fabs(r2) < fabs(eta))) { eta = r2; ok = true; } }
            if (!ok) {
                const double nw = -f / (dpsi + dphi), t = tau + nw;
                if (t > lo && t < hi) { eta = nw; ok = true; }
            }
        }
        if (!ok) eta = 0.5 * (lo + hi) - tau;
        const double tn = tau + eta;
        if (tn == tau) { status = 0; break; }
        tau = tn;
    }
    for (int i = par.first(); i < k; i += par.step()) delta[(long long)i * stride] = (dk[i] - origin) - tau;
    *lam_out = origin + tau;
    return status;
}

}  // namespace eigcore
}  // namespace xtbb200
