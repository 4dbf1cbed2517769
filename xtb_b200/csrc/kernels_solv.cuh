// Implicit solvation, GBSA / ALPB: Born radii (GBOBC II), solvent accessible surface on a 230 point
// Lebedev grid, generalised Born interaction matrix (Still and P16 kernels, Debye-Hueckel salt screening, hydrogen bond
// correction, ALPB shape term), the potential / energy hooks of the SCC and the gradient.
//
// Reference being replaced: src/solv/gbsa.f90 (TBorn: update 391-432, addBornMatrix 435-479, addShift 482-507, getEnergy
// 510-535, getEnergyParts 538-570, addGradient 573-640, getADet 643-687, addADetDeriv 690-791, getDebyeHueckel 794-821,
// compute_fhb 1013-1046, addGradientHBond 1049-1090), src/solv/born.f90:36-345, src/solv/sasa.f90:41-204,
// src/solv/kernel.f90:47-330,620-760.
//
// The reference stores d brad_i / d R_j and d sasa_i / d R_j as dense (3, N, N) arrays and contracts them with dgemv.
// Here the Born radius chain rule is applied pair by pair (no N^2 derivative array); the surface derivative is kept as a
// (3, N, N) array because its weights depend on the converged charges.  All sums have a fixed order: no atomics.
#pragma once
#include "common.cuh"

namespace xtbb200 {

constexpr int kSolvNang = 230;
constexpr int kSolvMaxNb = 640;          // neighbours of one atom inside the surface cutoff (shared-memory list)
__constant__ double c_leb230[kSolvNang * 4];

// smoothing of the surface switching function, src/solv/sasa.f90:30-36 / gbsa.f90:182-186
constexpr double kSolvW = 0.3 * kAAtoAU;
constexpr double kSolvAh0 = 0.5;
constexpr double kSolvAh1 = 3.0 / (4.0 * kSolvW);
constexpr double kSolvAh3 = -1.0 / (4.0 * (kSolvW * (kSolvW * kSolvW)));
constexpr double kSolvLrcut = 35.0 * kAAtoAU;   // src/solv/model.f90:166
constexpr double kZetaP16 = 1.028;              // src/solv/kernel.f90:40-41
constexpr double kZetaP16o16 = kZetaP16 / 16.0;

// per-atom model parameters sv_par[8]: vdwr, rho, svdw, vdwsa, wrp, gamsasa, hbmag, vdwr^3
enum { SVP_VDWR = 0, SVP_RHO, SVP_SVDW, SVP_VDWSA, SVP_WRP, SVP_GAM, SVP_HBMAG, SVP_RAD3 };
// per-molecule scalars sv_mol[24]
enum { SVM_GSASA = 0, SVM_ADET, SVM_CX, SVM_CY, SVM_CZ, SVM_TOTRAD3, SVM_I00, SVM_I01, SVM_I02, SVM_I11, SVM_I12, SVM_I22,
       SVM_GSOLV, SVM_GBORN, SVM_GHB, SVM_GSHIFT, SVM_QTOT, SVM_STRIDE = 24 };

// one descreening term psi_i <- j and (d psi / d r) / r, src/solv/born.f90:130-330.  r1_nonoverlap: the value of the
// reference's variable r1 when the non-overlap formula is evaluated: 1/r, except in the case of born_stale_r1 below.
__device__ __forceinline__ bool born_psi_term(double r, double r1_nonoverlap, double rho_j, double vdw_i, double& g, double& dg) {
    double r1 = 1.0 / r;
    if (r >= vdw_i + rho_j) {
        r1 = r1_nonoverlap;
        const double ap = r + rho_j, am = r - rho_j, ab = ap * am, rhab = rho_j / ab, lnab = 0.5 * log(am / ap) * r1;
        g = rhab + lnab;
        dg = -2.0 * rhab / ab + (rhab - lnab) * r1 * r1;
        return true;
    }
    if (r + rho_j > vdw_i) {
        const double r12 = 0.5 * r1, r24 = r12 * r12, ap = r + rho_j, am = r - rho_j, rh1 = 1.0 / vdw_i, rhr1 = 1.0 / ap,
                     aprh1 = ap * rh1, lnab = log(aprh1);
        g = rh1 - rhr1 + r12 * (0.5 * am * (rhr1 - rh1 * aprh1) - lnab);
        dg = rhr1 * rhr1 * (1.0 - 0.25 * am * r1 * (1.0 + aprh1 * aprh1)) + rho_j * r24 * (rhr1 - rh1 * aprh1) + r12 * (r1 * lnab - rhr1);
        dg = dg * r1;
        return true;
    }
    g = 0.0; dg = 0.0;
    return false;
}

// The reference's pair loop (compute_psi, born.f90:118-337; pairs (hi, lo), hi ascending, lo < hi ascending, inside
// lrcut) keeps 1/r in a variable r1 that every branch sets before use except one: when the descreening sphere of lo lies
// completely inside the van der Waals sphere of hi (r + rho_lo <= vdw_hi) while hi does not overlap lo (r >= rho_hi +
// vdw_lo), the term psi_lo <- hi (`case(1)`, born.f90:292-309) is evaluated with the r1 left by the preceding pair of the
// list that set it.  born_is_stale tells whether (hi, lo) is such a pair; born_stale_r1 walks the list backwards to the
// pair whose 1/r the reference would still hold (0 before the first assignment, where the reference is undefined).  No GFN2
// test system of the reference reaches this; tests/test_gpu_solvation.py constructs one (RbH in GBSA acetone).
__device__ __forceinline__ bool born_is_stale(double r, double vdw_hi, double rho_hi, double vdw_lo, double rho_lo) {
    return r < vdw_hi + rho_lo && !(r < rho_hi + vdw_lo) && !(r + rho_lo > vdw_hi);
}
__device__ __noinline__ double born_stale_r1(const Dev& d, const MolDesc& m, int hi, int lo) {
    int i = hi, j = lo;
    for (;;) {
        if (--j < 0) {
            if (--i < 1) return 0.0;
            j = i - 1;
        }
        const double* xi = d.xyz + 3 * (m.at0 + i);
        const double* xj = d.xyz + 3 * (m.at0 + j);
        const double dx = xi[0] - xj[0], dy = xi[1] - xj[1], dz = xi[2] - xj[2];
        const double r2 = dx * dx + dy * dy + dz * dz;
        if (!(r2 < kSolvLrcut * kSolvLrcut)) continue;
        const double r = sqrt(r2);
        const double* pi = d.sv_par + 8 * (size_t)(m.at0 + i);
        const double* pj = d.sv_par + 8 * (size_t)(m.at0 + j);
        const bool ovij = r < pi[SVP_VDWR] + pj[SVP_RHO], ovji = r < pi[SVP_RHO] + pj[SVP_VDWR];
        const bool ij_term = !ovij || r + pj[SVP_RHO] > pi[SVP_VDWR], ji_term = !ovji || r + pi[SVP_RHO] > pj[SVP_VDWR];
        if (ovij && !ovji && !ij_term) continue;      // a stale pair itself: r1 untouched
        if (ij_term || ji_term) return 1.0 / r;
    }
}

// ---- Born radii: CTA per atom, threads over the partners (compute_psi + compute_bornr) ----
__global__ void __launch_bounds__(128) k_solv_born(Dev d) {
    const MolDesc m = d.mol[blockIdx.y];
    const int i = blockIdx.x;
    if (i >= m.nat) return;
    __shared__ double red[32];
    const double* pi = d.sv_par + 8 * (size_t)(m.at0 + i);
    const double xi = d.xyz[3 * (m.at0 + i)], yi = d.xyz[3 * (m.at0 + i) + 1], zi = d.xyz[3 * (m.at0 + i) + 2];
    const double vdwi = pi[SVP_VDWR];
    double psi = 0.0;
    for (int j = threadIdx.x; j < m.nat; j += blockDim.x) {
        if (j == i) continue;
        const double dx = xi - d.xyz[3 * (m.at0 + j)], dy = yi - d.xyz[3 * (m.at0 + j) + 1], dz = zi - d.xyz[3 * (m.at0 + j) + 2];
        const double r2 = dx * dx + dy * dy + dz * dz;
        if (!(r2 < kSolvLrcut * kSolvLrcut)) continue;
        double g, dg;
        const double r = sqrt(r2);
        const double* pj = d.sv_par + 8 * (size_t)(m.at0 + j);
        double r1 = 1.0 / r;
        if (j > i && born_is_stale(r, pj[SVP_VDWR], pj[SVP_RHO], vdwi, pi[SVP_RHO])) {   // psi_i <- j of the pair (hi = j, lo = i)
            r1 = born_stale_r1(d, m, j, i);
            cnt_add(d, blockIdx.y, 11, 1);
        }
        if (born_psi_term(r, r1, pj[SVP_RHO], vdwi, g, dg)) psi += g;
    }
    psi = block_sum(psi, red);
    if (threadIdx.x == 0) {
        const double alp = 1.0, bet = 0.8, gam = 4.85;   // src/solv/born.f90:28-30
        const double s1 = 1.0 / pi[SVP_SVDW], v1 = 1.0 / vdwi, s2 = 0.5 * pi[SVP_SVDW];
        double br = psi * s2;
        double arg2 = br * (gam * br - bet);
        const double arg = br * (alp + arg2);
        arg2 = 2.0 * arg2 + alp + gam * br * br;
        const double th = tanh(arg), ch = cosh(arg);
        br = d.solv_c1 / (s1 - v1 * th);
        double dpsi = ch * (s1 - v1 * th);
        dpsi = d.solv_c1 * s2 * v1 * arg2 / (dpsi * dpsi);
        d.sv_brad[m.at0 + i] = br;
        d.sv_dbr[m.at0 + i] = dpsi;
        if (d.solv_salt) {   // getDebyeHueckel
            const double aa = 0.5 / d.solv_eps, gg = d.solv_kappa * (br + d.solv_ionrad);
            const double ionscr = aa * exp(gg) / (1.0 + gg);
            d.sv_ion[2 * (m.at0 + i)] = ionscr;
            d.sv_ion[2 * (m.at0 + i) + 1] = ionscr * d.solv_kappa * gg / (1.0 + gg);
        }
    }
}

// value and derivative factor of the switching function of neighbour ia at the grid point; returns 0 outside the outer
// cutoff, 1 inside the switching shell, 2 when the point is buried (compute_w_sp, src/solv/sasa.f90:128-204)
__device__ __forceinline__ int sasa_switch(double tx, double ty, double tz, double vdwsa, double& sasaij, double& dfac) {
    const double tj2 = tx * tx + ty * ty + tz * tz;
    const double lo = (vdwsa - kSolvW) * (vdwsa - kSolvW), hi = (vdwsa + kSolvW) * (vdwsa + kSolvW);
    if (!(tj2 < hi)) return 0;
    if (tj2 <= lo) return 2;
    const double sqtj = sqrt(tj2), uj = sqtj - vdwsa, ah3uj2 = kSolvAh3 * uj * uj;
    double dsasaij = kSolvAh1 + 3.0 * ah3uj2;
    sasaij = kSolvAh0 + (kSolvAh1 + ah3uj2) * uj;
    sasaij = fmax(0.0, fmin(sasaij, 1.0));
    if (sasaij < 2.2250738585072014e-308) return 2;
    dfac = dsasaij / (sasaij * sqtj);
    return 1;
}

// ---- solvent accessible surface and its derivative: CTA per atom, thread per grid point (compute_numsa) ----
// pass 1: thread ip multiplies the switching functions of all neighbours; pass 2: warp per neighbour, lanes over the
// grid points, so that every d sasa_i / d R_j is one warp reduction.  dsdrt must be zero on entry.
__global__ void __launch_bounds__(256) k_solv_sasa(Dev d) {
    const int im = blockIdx.y;
    const MolDesc m = d.mol[im];
    const int i = blockIdx.x;
    if (i >= m.nat) return;
    __shared__ double s_x[kSolvMaxNb], s_y[kSolvMaxNb], s_z[kSolvMaxNb], s_r[kSolvMaxNb];
    __shared__ int s_j[kSolvMaxNb];
    __shared__ double s_wsa[256];
    __shared__ double red[32];
    __shared__ double s_self[8][3];
    __shared__ int s_n;
    const int tid = threadIdx.x;
    const double* pi = d.sv_par + 8 * (size_t)(m.at0 + i);
    const double xi = d.xyz[3 * (m.at0 + i)], yi = d.xyz[3 * (m.at0 + i) + 1], zi = d.xyz[3 * (m.at0 + i) + 2];
    const double rsas = pi[SVP_VDWSA];
    // neighbour list in ascending atom order (ballot compaction by warp 0): an atom can switch a grid point of i only if
    // r_ij < vdwsa_i + vdwsa_j + w, which lies inside the reference's list cutoff srcut for every pair
    if (tid < 32) {
        int n = 0;
        for (int j0 = 0; j0 < m.nat; j0 += 32) {
            const int j = j0 + tid;
            bool in = false;
            double xj = 0, yj = 0, zj = 0, rj = 0;
            if (j < m.nat && j != i) {
                xj = d.xyz[3 * (m.at0 + j)]; yj = d.xyz[3 * (m.at0 + j) + 1]; zj = d.xyz[3 * (m.at0 + j) + 2];
                rj = d.sv_par[8 * (size_t)(m.at0 + j) + SVP_VDWSA];
                const double dx = xi - xj, dy = yi - yj, dz = zi - zj, rc = rsas + rj + kSolvW;
                in = dx * dx + dy * dy + dz * dz < rc * rc;
            }
            const unsigned b = __ballot_sync(0xffffffffu, in);
            if (in) {
                const int p = n + __popc(b & ((1u << tid) - 1u));
                if (p < kSolvMaxNb) { s_x[p] = xj; s_y[p] = yj; s_z[p] = zj; s_r[p] = rj; s_j[p] = j; }
            }
            n += __popc(b);
        }
        if (tid == 0) {
            if (n > kSolvMaxNb) { d.st[im].fail = 30; n = kSolvMaxNb; }   // surface neighbour list overflow
            s_n = n;
        }
    }
    __syncthreads();
    const int nnb = s_n;
    double wsa = 0.0;
    if (tid < kSolvNang) {
        const double px = xi + rsas * c_leb230[4 * tid], py = yi + rsas * c_leb230[4 * tid + 1], pz = zi + rsas * c_leb230[4 * tid + 2];
        double sasap = 1.0;
        for (int t = 0; t < nnb; ++t) {
            double sij, df;
            const int c = sasa_switch(px - s_x[t], py - s_y[t], pz - s_z[t], s_r[t], sij, df);
            if (c == 2) { sasap = 0.0; break; }
            if (c == 1) sasap *= sij;
        }
        if (sasap > 1.e-6) wsa = c_leb230[4 * tid + 3] * pi[SVP_WRP] * sasap;
    }
    s_wsa[tid] = wsa;
    const double sasai = block_sum(wsa, red);   // contains the barriers that publish s_wsa
    const int warp = tid >> 5, lane = tid & 31;
    double sx = 0, sy = 0, sz = 0;   // d sasa_i / d R_i, partial sum of this warp (lane 0)
    double* row = d.sv_dsdrt + 3 * (m.nn0 + (long long)i * m.nat);
    for (int t = warp; t < nnb; t += 8) {
        double gx = 0, gy = 0, gz = 0;
        for (int ip = lane; ip < kSolvNang; ip += 32) {
            const double w = s_wsa[ip];
            if (w == 0.0) continue;
            const double tx = xi + rsas * c_leb230[4 * ip] - s_x[t], ty = yi + rsas * c_leb230[4 * ip + 1] - s_y[t],
                         tz = zi + rsas * c_leb230[4 * ip + 2] - s_z[t];
            double sij, df;
            if (sasa_switch(tx, ty, tz, s_r[t], sij, df) == 1) { gx += w * df * tx; gy += w * df * ty; gz += w * df * tz; }
        }
        gx = warp_sum(gx); gy = warp_sum(gy); gz = warp_sum(gz);
        if (lane == 0) {
            const int j = s_j[t];
            row[3 * j] = -gx; row[3 * j + 1] = -gy; row[3 * j + 2] = -gz;
            sx += gx; sy += gy; sz += gz;
        }
    }
    if (lane == 0) { s_self[warp][0] = sx; s_self[warp][1] = sy; s_self[warp][2] = sz; }
    __syncthreads();
    if (tid < 3) {
        double s = 0;
        for (int w = 0; w < 8; ++w) s += s_self[w][tid];
        row[3 * i + tid] = s;
    }
    if (tid == 0) d.sv_sasa[m.at0 + i] = sasai;
}

// ---- per molecule: surface energy, inertia tensor of the van der Waals spheres and the ALPB shape descriptor ----
__global__ void __launch_bounds__(256) k_solv_shape(Dev d) {
    const int im = blockIdx.x;
    const MolDesc m = d.mol[im];
    __shared__ double red[32];
    const int tid = threadIdx.x, nt = blockDim.x;
    double* sm = d.sv_mol + (size_t)SVM_STRIDE * im;
    double p = 0, t3 = 0, cx = 0, cy = 0, cz = 0;
    for (int i = tid; i < m.nat; i += nt) {
        const double* pa = d.sv_par + 8 * (size_t)(m.at0 + i);
        p += d.sv_sasa[m.at0 + i] * pa[SVP_GAM];
        const double rad3 = pa[SVP_RAD3];
        t3 += rad3;
        cx += d.xyz[3 * (m.at0 + i)] * rad3; cy += d.xyz[3 * (m.at0 + i) + 1] * rad3; cz += d.xyz[3 * (m.at0 + i) + 2] * rad3;
    }
    const double gsasa = block_sum(p, red), tot = block_sum(t3, red);
    cx = block_sum(cx, red) / tot; cy = block_sum(cy, red) / tot; cz = block_sum(cz, red) / tot;
    const double tof = 2.0 / 5.0;
    double in[6] = {0, 0, 0, 0, 0, 0};   // 00 01 02 11 12 22
    for (int i = tid; i < m.nat; i += nt) {
        const double* pa = d.sv_par + 8 * (size_t)(m.at0 + i);
        const double rad3 = pa[SVP_RAD3], rad2 = pa[SVP_VDWR] * pa[SVP_VDWR];
        const double vx = d.xyz[3 * (m.at0 + i)] - cx, vy = d.xyz[3 * (m.at0 + i) + 1] - cy, vz = d.xyz[3 * (m.at0 + i) + 2] - cz;
        const double dg = rad3 * (vx * vx + vy * vy + vz * vz + tof * rad2);
        in[0] += dg - rad3 * vx * vx; in[1] -= rad3 * vx * vy; in[2] -= rad3 * vx * vz;
        in[3] += dg - rad3 * vy * vy; in[4] -= rad3 * vy * vz; in[5] += dg - rad3 * vz * vz;
    }
    for (int k = 0; k < 6; ++k) in[k] = block_sum(in[k], red);
    if (tid == 0) {
        const double det = in[0] * (in[3] * in[5] - in[4] * in[4]) - in[1] * (in[1] * in[5] - in[4] * in[2]) +
                           in[2] * (in[1] * in[4] - in[3] * in[2]);
        sm[SVM_GSASA] = gsasa;
        sm[SVM_ADET] = sqrt(pow(det, 1.0 / 3.0) / (tof * tot));
        sm[SVM_CX] = cx; sm[SVM_CY] = cy; sm[SVM_CZ] = cz; sm[SVM_TOTRAD3] = tot;
        for (int k = 0; k < 6; ++k) sm[SVM_I00 + k] = in[k];
    }
}

// ---- Born interaction matrix, thread per element (addBornMatrix) ----
__global__ void k_solv_bornmat(Dev d) {
    const int im = blockIdx.y;
    const MolDesc m = d.mol[im];
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)m.nat * m.nat) return;
    const int i = (int)(t % m.nat), j = (int)(t / m.nat);
    const double bi = d.sv_brad[m.at0 + i];
    double v;
    if (i == j) {
        if (d.solv_kernel == 1 && d.solv_salt) v = (exp(-d.solv_kappa * bi) * d.sv_ion[2 * (m.at0 + i)] * 2.0 - 1.0) / bi;
        else v = d.solv_keps / bi;
        if (d.solv_hb) {   // compute_fhb: hbw = hbmag * sasa / vdwsa^2
            const double* pa = d.sv_par + 8 * (size_t)(m.at0 + i);
            v += 2.0 * pa[SVP_HBMAG] * d.sv_sasa[m.at0 + i] / (pa[SVP_VDWSA] * pa[SVP_VDWSA]);
        }
    } else {
        const double bj = d.sv_brad[m.at0 + j];
        const double dx = d.xyz[3 * (m.at0 + i)] - d.xyz[3 * (m.at0 + j)], dy = d.xyz[3 * (m.at0 + i) + 1] - d.xyz[3 * (m.at0 + j) + 1],
                     dz = d.xyz[3 * (m.at0 + i) + 2] - d.xyz[3 * (m.at0 + j) + 2];
        const double r2 = dx * dx + dy * dy + dz * dz;
        if (d.solv_kernel == 1) {
            const double aa = bi * bj, dd = 0.25 * r2 / aa, expd = exp(-dd), fgb2 = r2 + aa * expd;
            if (d.solv_salt) {
                const double fgb = sqrt(fgb2), gg = d.sv_ion[2 * (m.at0 + i)] + d.sv_ion[2 * (m.at0 + j)];
                v = (exp(-d.solv_kappa * fgb) * gg - 1.0) / fgb;
            } else v = d.solv_keps / sqrt(fgb2);
        } else {
            const double r1 = sqrt(r2), ab = sqrt(bi * bj);
            double arg = ab / (ab + kZetaP16o16 * r1);
            arg = arg * arg; arg = arg * arg; arg = arg * arg; arg = arg * arg;
            v = d.solv_keps / (r1 + ab * arg);
        }
    }
    if (d.solv_alpbet > 0.0) v += d.solv_keps * d.solv_alpbet / d.sv_mol[(size_t)SVM_STRIDE * im + SVM_ADET];
    d.sv_bmat[m.nn0 + t] = v;
}

// ---- B q: 8 column slices per 32 rows as the other row kernels; which 0: mixed / input charges -> sv_shift (addShift),
//      1: output charges of the iteration -> sv_bq (getEnergy) ----
__global__ void __launch_bounds__(256) k_solv_matvec(Dev d, int which) {
    const int im = blockIdx.y;
    const MolDesc m = d.mol[im];
    if (which >= 0 && d.st[im].done) return;
    if (which < 0) which = 0;
    __shared__ double part[8][33];
    const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + lane;
    if (blockIdx.x * 32 >= m.nat) return;
    const double* q = (which ? d.q_out : d.q) + m.at0;
    double s = 0.0;
    if (i < m.nat) {
        const double* col = d.sv_bmat + m.nn0 + i;   // symmetric: row i = column i, consecutive lanes read consecutive addresses
        for (int j = slice; j < m.nat; j += 8) s += col[(long long)j * m.nat] * q[j];
    }
    part[slice][lane] = s;
    __syncthreads();
    if (slice == 0 && i < m.nat) {
        double t = 0;
        for (int k = 0; k < 8; ++k) t += part[k][lane];
        (which ? d.sv_bq : d.sv_shift)[m.at0 + i] = t;
    }
}

// ---- gradient, part 1: CTA per atom, threads over the partners: direct pair terms of the Born energy and the
//      derivative of the energy with respect to the Born radius of the atom (addGradientStill / SaltStill / P16) ----
__global__ void __launch_bounds__(64) k_solv_grad_pairs(Dev d) {
    const MolDesc m = d.mol[blockIdx.y];
    const int i = blockIdx.x;
    if (i >= m.nat) return;
    __shared__ double red[32];
    const double xi = d.xyz[3 * (m.at0 + i)], yi = d.xyz[3 * (m.at0 + i) + 1], zi = d.xyz[3 * (m.at0 + i) + 2];
    const double bi = d.sv_brad[m.at0 + i], qi = d.q[m.at0 + i];
    const bool salt = d.solv_kernel == 1 && d.solv_salt;
    const double ioni = salt ? d.sv_ion[2 * (m.at0 + i)] : 0.0, disi = salt ? d.sv_ion[2 * (m.at0 + i) + 1] : 0.0;
    double gx = 0, gy = 0, gz = 0, gb = 0;
    for (int j = threadIdx.x; j < m.nat; j += blockDim.x) {
        if (j == i) continue;
        const double dx = xi - d.xyz[3 * (m.at0 + j)], dy = yi - d.xyz[3 * (m.at0 + j) + 1], dz = zi - d.xyz[3 * (m.at0 + j) + 2];
        const double r2 = dx * dx + dy * dy + dz * dz;
        const double bj = d.sv_brad[m.at0 + j], qq = qi * d.q[m.at0 + j];
        double ap, gbi;   // gradient_i -= ap * dr ; grddb_i += gbi
        if (salt) {
            double aa = bi * bj;
            const double dd = 0.25 * r2 / aa, expd = exp(-dd), dfgb = r2 + aa * expd, fgb = sqrt(dfgb);
            aa = d.solv_kappa * fgb;
            const double expa = exp(-aa), gg = (ioni + d.sv_ion[2 * (m.at0 + j)]) * expa;
            double qfg = qq / fgb;
            const double dfgb3 = qfg * (gg * (1.0 + aa) - 1.0) / dfgb;
            ap = (1.0 - 0.25 * expd) * dfgb3;
            qfg = qfg * expa;
            gbi = bj * (-0.5 * expd * (1.0 + dd) * dfgb3) + qfg * disi;
        } else if (d.solv_kernel == 1) {
            const double aa = bi * bj, dd = 0.25 * r2 / aa, expd = exp(-dd), fgb2 = r2 + aa * expd, dfgb2 = 1.0 / fgb2,
                         dfgb = sqrt(dfgb2), dfgb3 = dfgb2 * dfgb * d.solv_keps;
            ap = (1.0 - 0.25 * expd) * dfgb3 * qq;
            gbi = bj * (-0.5 * expd * (1.0 + dd) * dfgb3) * qq;
        } else {
            const double r1 = sqrt(r2), ab = sqrt(bi * bj), arg1 = ab / (ab + kZetaP16o16 * r1);
            double arg16 = arg1 * arg1;
            arg16 = arg16 * arg16; arg16 = arg16 * arg16; arg16 = arg16 * arg16;
            const double fgb = r1 + ab * arg16, dfgb = 1.0 / fgb, dfgb2 = dfgb * dfgb;
            ap = (1.0 - kZetaP16 * arg1 * arg16) * dfgb2 * d.solv_keps / r1 * qq;
            gbi = bj * (-0.5 * (r1 * kZetaP16 / ab * arg1 + 1.0) / ab * arg16 * dfgb2) * d.solv_keps * qq;
        }
        gx -= ap * dx; gy -= ap * dy; gz -= ap * dz;
        gb += gbi;
    }
    gx = block_sum(gx, red); gy = block_sum(gy, red); gz = block_sum(gz, red); gb = block_sum(gb, red);
    if (threadIdx.x == 0) {
        if (salt) {   // self-energy part
            const double gg = exp(-d.solv_kappa * bi), aa = 2.0 * ioni * gg - 1.0, qq = qi / bi;
            const double ap = aa - bi * 2.0 * (disi + ioni * d.solv_kappa) * gg;
            gb += -0.5 * qq * qq * ap;
        } else {
            const double bp = 1.0 / bi, qq = qi * bp;
            gb += -0.5 * d.solv_keps * qq * bp * qi;
        }
        d.sv_grddb[m.at0 + i] = gb;
        // weight of d sasa_i / d R in the gradient: surface tension + hydrogen bond term dhbdw q^2 (addGradientHBond)
        const double* pa = d.sv_par + 8 * (size_t)(m.at0 + i);
        double c = pa[SVP_GAM];
        if (d.solv_hb) c += pa[SVP_HBMAG] / (pa[SVP_VDWSA] * pa[SVP_VDWSA]) * qi * qi;
        d.sv_bq[m.at0 + i] = c;   // sv_bq is free after the SCC
        if (d.srank == 0) { d.grad[3 * (m.at0 + i)] += gx; d.grad[3 * (m.at0 + i) + 1] += gy; d.grad[3 * (m.at0 + i) + 2] += gz; }
    }
}

// ---- gradient, part 2: CTA per atom a: chain rule through the Born radii, surface / hydrogen bond term, ALPB shape ----
__global__ void __launch_bounds__(64) k_solv_grad_chain(Dev d) {
    const int im = blockIdx.y;
    const MolDesc m = d.mol[im];
    const int a = blockIdx.x;
    if (a >= m.nat || d.srank != 0) return;
    __shared__ double red[32];
    const double* pa = d.sv_par + 8 * (size_t)(m.at0 + a);
    const double xa = d.xyz[3 * (m.at0 + a)], ya = d.xyz[3 * (m.at0 + a) + 1], za = d.xyz[3 * (m.at0 + a) + 2];
    const double wa = d.sv_grddb[m.at0 + a] * d.sv_dbr[m.at0 + a];
    double gx = 0, gy = 0, gz = 0;
    for (int b = threadIdx.x; b < m.nat; b += blockDim.x) {
        // surface term: sum_b c_b d sasa_b / d R_a
        const double cb = d.sv_bq[m.at0 + b];
        const double* ds = d.sv_dsdrt + 3 * (m.nn0 + (long long)b * m.nat + a);
        gx += cb * ds[0]; gy += cb * ds[1]; gz += cb * ds[2];
        if (b == a) continue;
        const double dx = xa - d.xyz[3 * (m.at0 + b)], dy = ya - d.xyz[3 * (m.at0 + b) + 1], dz = za - d.xyz[3 * (m.at0 + b) + 2];
        const double r2 = dx * dx + dy * dy + dz * dz;
        if (!(r2 < kSolvLrcut * kSolvLrcut)) continue;
        const double r = sqrt(r2);
        const double* pb = d.sv_par + 8 * (size_t)(m.at0 + b);
        double g, dga, dgb;
        double r1a = 1.0 / r, r1b = r1a;
        if (b > a) { if (born_is_stale(r, pb[SVP_VDWR], pb[SVP_RHO], pa[SVP_VDWR], pa[SVP_RHO])) r1a = born_stale_r1(d, m, b, a); }
        else if (born_is_stale(r, pa[SVP_VDWR], pa[SVP_RHO], pb[SVP_VDWR], pb[SVP_RHO])) r1b = born_stale_r1(d, m, a, b);
        born_psi_term(r, r1a, pb[SVP_RHO], pa[SVP_VDWR], g, dga);   // psi_a <- b
        born_psi_term(r, r1b, pa[SVP_RHO], pb[SVP_VDWR], g, dgb);   // psi_b <- a
        const double f = wa * dga + d.sv_grddb[m.at0 + b] * d.sv_dbr[m.at0 + b] * dgb;
        gx += f * dx; gy += f * dy; gz += f * dz;
    }
    gx = block_sum(gx, red); gy = block_sum(gy, red); gz = block_sum(gz, red);
    if (threadIdx.x == 0) {
        if (d.solv_alpbet > 0.0) {   // addADetDeriv
            const double* sm = d.sv_mol + (size_t)SVM_STRIDE * im;
            const double adet = sm[SVM_ADET], tot = sm[SVM_TOTRAD3], qtot = sm[SVM_QTOT];
            const double i00 = sm[SVM_I00], i01 = sm[SVM_I01], i02 = sm[SVM_I02], i11 = sm[SVM_I11], i12 = sm[SVM_I12], i22 = sm[SVM_I22];
            const double kE = d.solv_keps * d.solv_alpbet;
            const double f = (250.0 / (48.0 * tot * tot * tot * pow(adet, 5))) * (-0.5 * kE * qtot * qtot / (adet * adet));
            const double a00 = i00 * (i11 + i22) - i01 * i01 - i02 * i02, a01 = i01 * i22 - i02 * i12, a02 = i02 * i11 - i01 * i12,
                         a11 = i11 * (i00 + i22) - i01 * i01 - i12 * i12, a12 = i00 * i12 - i01 * i02,
                         a22 = i22 * (i00 + i11) - i02 * i02 - i12 * i12;
            const double vx = xa - sm[SVM_CX], vy = ya - sm[SVM_CY], vz = za - sm[SVM_CZ], rad3 = pa[SVP_RAD3];
            gx += rad3 * f * (a00 * vx + a01 * vy + a02 * vz);
            gy += rad3 * f * (a01 * vx + a11 * vy + a12 * vz);
            gz += rad3 * f * (a02 * vx + a12 * vy + a22 * vz);
        }
        d.grad[3 * (m.at0 + a)] += gx; d.grad[3 * (m.at0 + a) + 1] += gy; d.grad[3 * (m.at0 + a) + 2] += gz;
    }
}

// ---- energy parts with the final (mixed) charges: getEnergyParts; sv_shift must hold B q of these charges ----
__global__ void __launch_bounds__(256) k_solv_energy(Dev d) {
    const int im = blockIdx.x;
    const MolDesc m = d.mol[im];
    __shared__ double red[32];
    const int tid = threadIdx.x, nt = blockDim.x;
    double e = 0, hb = 0, qt = 0;
    for (int i = tid; i < m.nat; i += nt) {
        const double q = d.q[m.at0 + i];
        e += 0.5 * q * d.sv_shift[m.at0 + i];
        qt += q;
        if (d.solv_hb) {
            const double* pa = d.sv_par + 8 * (size_t)(m.at0 + i);
            hb += pa[SVP_HBMAG] * d.sv_sasa[m.at0 + i] / (pa[SVP_VDWSA] * pa[SVP_VDWSA]) * q * q;
        }
    }
    e = block_sum(e, red); hb = block_sum(hb, red); qt = block_sum(qt, red);
    if (tid == 0) {
        double* sm = d.sv_mol + (size_t)SVM_STRIDE * im;
        sm[SVM_GBORN] = e - hb; sm[SVM_GHB] = hb; sm[SVM_GSHIFT] = d.solv_gshift;
        sm[SVM_GSOLV] = (e - hb) + sm[SVM_GSASA] + hb + d.solv_gshift;
        sm[SVM_QTOT] = qt;
    }
}

}  // namespace xtbb200
