// Nuclear gradient kernels.
//
// Reference: src/scf_module.F90:665-798; src/xtb/hamiltonian.F90:575-792 (build_dSDQH0_noreset);
// src/intgrad.f90:607-728, 958-1035 (multipole_grad_3d, shiftintg, get_grad_multiint);
// src/grad_core.f90:40-72 (dshellPoly); src/aespot.F90:1070-1165 (aniso_grad), 1292-1307 (dradcn);
// src/disp/dftd4.F90:948-1147 (weight_references, get_atomic_c6), 1832-1933 (disp_gradient_latp),
// 2023-2177 + 2354-2433 (atm_gradient_latp, deriv_atm_triple); src/coulomb/klopmanohno.F90:471-573;
// src/disp/coordinationnumber.F90:351-479 (dCN/dR, used pairwise instead of the dense dcndr tensor).
#pragma once
#include "common.cuh"
#include "kernels_int.cuh"

namespace xtbb200 {

// moments and derivative (w.r.t. centre i) moments for one Cartesian direction, intgrad.f90:607-660
__device__ __forceinline__ void moments_grad_1d(int li, int lj, double ai, double pa, double pb, double rpc,
                                                const double* s1d, double* val, double* gra) {
    double vi[4] = {0, 0, 0, 0}, vj[3] = {0, 0, 0}, gi[5] = {0, 0, 0, 0, 0};
    vi[li] = 1.0; vj[lj] = 1.0;
    gi[li + 1] = 2 * ai;
    if (li > 0) gi[li - 1] = -li;
    // shift the (li-1) term first, then the (li+1) term (order matters, intgrad.f90:640-641)
    if (li - 1 >= 1) shift_poly(pa, li - 1, gi);
    {
        const int l = li + 1;   // 1..3
        if (l == 1) gi[0] += pa * gi[1];
        else if (l == 2) { gi[0] += pa * pa * gi[2]; gi[1] += 2 * pa * gi[2]; }
        else { gi[0] += pa * pa * pa * gi[3]; gi[1] += 3 * pa * pa * gi[3]; gi[2] += 3 * pa * gi[3]; }
    }
    shift_poly(pa, li, vi);
    shift_poly(pb, lj, vj);
    double v0 = 0, v1 = 0, v2 = 0, g0 = 0, g1 = 0, g2 = 0;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        double sv = 0.0, sg = 0.0;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int b = k - a;
            if (b >= 0 && b <= 2) { if (a <= 2) sv += vi[a] * vj[b]; sg += gi[a] * vj[b]; }
        }
        if (k <= li + lj + 1) {
            const double m0 = s1d[k], m1 = s1d[k + 1] + rpc * s1d[k], m2 = s1d[k + 2] + 2 * rpc * s1d[k + 1] + rpc * rpc * s1d[k];
            v0 += m0 * sv; v1 += m1 * sv; v2 += m2 * sv;
            g0 += m0 * sg; g1 += m1 * sg; g2 += m2 * sg;
        }
    }
    val[0] = v0; val[1] = v1; val[2] = v2;
    gra[0] = g0; gra[1] = g1; gra[2] = g2;
}

// The same for all exponent pairs (a <= LI, b <= LJ) of one direction: out[(a (LJ + 1) + b) 6 + {val 0..2, gra 0..2}], static
// loops (LI / LJ are uniform over the warp), sums in the order of moments_grad_1d.
template <int LI, int LJ>
__device__ __forceinline__ void moments_grad_table(double ai, double pa, double pb, double rpc, const double* s1d, double* out) {
    constexpr int NK = LI + LJ + 2;   // powers 0 .. LI + LJ + 1 of (x - P)
    double m0[NK], m1[NK], m2[NK];
#pragma unroll
    for (int k = 0; k < NK; ++k) {
        m0[k] = s1d[k];
        m1[k] = s1d[k + 1] + rpc * s1d[k];
        m2[k] = s1d[k + 2] + 2 * rpc * s1d[k + 1] + rpc * rpc * s1d[k];
    }
    double pj[LJ + 1][3];
#pragma unroll
    for (int b = 0; b <= LJ; ++b) {
        pj[b][0] = pj[b][1] = pj[b][2] = 0.0;
        pj[b][b] = 1.0;
        shift_poly(pb, b, pj[b]);
    }
#pragma unroll
    for (int a = 0; a <= LI; ++a) {
        double vi[4] = {0, 0, 0, 0}, gi[5] = {0, 0, 0, 0, 0};
        vi[a] = 1.0;
        gi[a + 1] = 2 * ai;
        if (a > 0) gi[a - 1] = -a;
        if (a - 1 >= 1) shift_poly(pa, a - 1, gi);
        if (a + 1 == 1) gi[0] += pa * gi[1];
        else if (a + 1 == 2) { gi[0] += pa * pa * gi[2]; gi[1] += 2 * pa * gi[2]; }
        else { gi[0] += pa * pa * pa * gi[3]; gi[1] += 3 * pa * pa * gi[3]; gi[2] += 3 * pa * gi[3]; }
        shift_poly(pa, a, vi);
#pragma unroll
        for (int b = 0; b <= LJ; ++b) {
            double v0 = 0, v1 = 0, v2 = 0, g0 = 0, g1 = 0, g2 = 0;
#pragma unroll
            for (int k = 0; k <= a + b + 1; ++k) {
                double sv = 0.0, sg = 0.0;
#pragma unroll
                for (int q = 0; q <= a + 1; ++q) {
                    const int r = k - q;
                    if (r >= 0 && r <= b) { if (q <= a) sv += vi[q] * pj[b][r]; sg += gi[q] * pj[b][r]; }
                }
                v0 += m0[k] * sv; v1 += m1[k] * sv; v2 += m2[k] * sv;
                g0 += m0[k] * sg; g1 += m1[k] * sg; g2 += m2[k] * sg;
            }
            double* o = out + 6 * (a * (LJ + 1) + b);
            o[0] = v0; o[1] = v1; o[2] = v2; o[3] = g0; o[4] = g1; o[5] = g2;
        }
    }
}
__device__ __forceinline__ void moments_grad_table_dyn(int li, int lj, double ai, double pa, double pb, double rpc, const double* s1d,
                                                       double* out) {
    switch (3 * li + lj) {
        case 0: moments_grad_table<0, 0>(ai, pa, pb, rpc, s1d, out); break;
        case 1: moments_grad_table<0, 1>(ai, pa, pb, rpc, s1d, out); break;
        case 2: moments_grad_table<0, 2>(ai, pa, pb, rpc, s1d, out); break;
        case 3: moments_grad_table<1, 0>(ai, pa, pb, rpc, s1d, out); break;
        case 4: moments_grad_table<1, 1>(ai, pa, pb, rpc, s1d, out); break;
        case 5: moments_grad_table<1, 2>(ai, pa, pb, rpc, s1d, out); break;
        case 6: moments_grad_table<2, 0>(ai, pa, pb, rpc, s1d, out); break;
        case 7: moments_grad_table<2, 1>(ai, pa, pb, rpc, s1d, out); break;
        default: moments_grad_table<2, 2>(ai, pa, pb, rpc, s1d, out); break;
    }
}
constexpr int kGradTab = 888;         // doubles of the moment table per shell pair (12 primitive pairs of a p-p shell pair: 16 take two fillings, but three CTAs fit an SM)

__constant__ int c_ord[10][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {2, 0, 0}, {0, 2, 0}, {0, 0, 2}, {1, 1, 0}, {1, 0, 1}, {0, 1, 1}};

// Hamiltonian gradient: one warp per off-site shell pair (is > js, different atoms, r^2 <= 1600).
// The spherical weights are transformed to the Cartesian basis so that each lane contracts its own
// Cartesian component pair; result per pair: gpair[p] = (gx, gy, gz, dCN).
constexpr size_t kGradSmem = (size_t)kIntPairs * kGradTab * sizeof(double);   // dynamic shared memory of k_grad_h0 (the tables)
// half a warp per shell pair, as k_sdqh0: `warp` = group of kIntLG lanes, `lane` = lane of the group
__global__ void __launch_bounds__(kIntWarps * 32, 3) k_grad_h0(Dev d) {
    const MolDesc m = d.mol[blockIdx.y];
    const int warp = threadIdx.x / kIntLG, lane = threadIdx.x % kIntLG;
    const unsigned gmask = kIntLG == 32 ? 0xffffffffu : ((1u << kIntLG) - 1u) << ((threadIdx.x & 31) / kIntLG * kIntLG);
    auto group_sum = [&](double v) {
#pragma unroll
        for (int o = kIntLG / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o);
        return v;
    };
    const long long npair = (long long)m.nsh * (m.nsh + 1) / 2;
    const long long pk = shard_pair_cta(d) * kIntPairs + warp;
    __shared__ double s_alp[kIntPairs][2][6], s_cont[kIntPairs][2][6];
    __shared__ double s_wf[kIntPairs][25], s_wp[kIntPairs][25];
    extern __shared__ __align__(16) double s_dyn_grad[];                  // kGradSmem bytes
    double (*s_tabs)[kGradTab] = reinterpret_cast<double (*)[kGradTab]>(s_dyn_grad);
    __shared__ double s_kab[kIntPairs][kIntMaxPP];
    __shared__ unsigned char s_ipjp[kIntPairs][kIntMaxPP];
    double* tab = &s_tabs[warp][0];
    if (pk >= npair) return;
    const long long p = d.pair_order[d.pord0[blockIdx.y] + pk];
    int is = (int)((sqrt(8.0 * (double)p + 1.0) - 1.0) * 0.5);
    while ((long long)is * (is + 1) / 2 > p) --is;
    while ((long long)(is + 1) * (is + 2) / 2 <= p) ++is;
    const int js = (int)(p - (long long)is * (is + 1) / 2);
    double* out = d.gpair + 4 * (d.pair0[blockIdx.y] + p);
    const ShellPar& si = d.shells[m.sh0 + is];
    const ShellPar& sj = d.shells[m.sh0 + js];
    const int li = si.l, lj = sj.l, iat = si.atom, jat = sj.atom;
    const double* ri = d.xyz + 3 * (m.at0 + iat);
    const double* rj = d.xyz + 3 * (m.at0 + jat);
    const double rij[3] = {ri[0] - rj[0], ri[1] - rj[1], ri[2] - rj[2]};
    const double rij2 = rij[0] * rij[0] + rij[1] * rij[1] + rij[2] * rij[2];
    if (iat == jat || rij2 > 1600.0) {
        if (lane < 4) out[lane] = 0.0;
        return;
    }
    if (lane == 0) cnt_add(d, blockIdx.y, 2, 1);
    const int nci = c_ncart[li], ncj = c_ncart[lj], npi = si.nprim, npj = sj.nprim;
    const int nsi = c_nsph[li], nsj = c_nsph[lj];
    if (lane < 6) { s_alp[warp][0][lane] = si.alp[lane]; s_cont[warp][0][lane] = si.cont[lane]; }
    else if (lane < 12) { s_alp[warp][1][lane - 6] = sj.alp[lane - 6]; s_cont[warp][1][lane - 6] = sj.cont[lane - 6]; }
    for (int q = lane; q < kIntMaxPP; q += kIntLG) {
        const int ip = q / npj;
        s_ipjp[warp][q] = (unsigned char)(ip | ((q - ip * npj) << 4));
    }
    const int nao = m.nao;
    // spherical weights + shell-pair scalars
    const double ves = (d.sshift[m.sh0 + is] + d.ashift[m.at0 + iat]) + (d.sshift[m.sh0 + js] + d.ashift[m.at0 + jat]);
    const double vss = d.vs[m.at0 + iat] + d.vs[m.at0 + jat];
    const double hav = 0.5 * (d.selfE[m.sh0 + is] + d.selfE[m.sh0 + js]);
    double hps = 0.0;   // sum HPij * S
    for (int e = lane; e < nsi * nsj; e += kIntLG) {
        const int ii = e / nsj, jj = e % nsj;
        const long long a = m.mat0 + (sj.ao0 + jj) + (long long)nao * (si.ao0 + ii);
        const double Pij = d.P[a];
        const double HPij = d.H0n[a] * kEVtoAU * Pij;
        s_wf[warp][e] = 2 * HPij - 2 * d.H[a] - Pij * ves + Pij * vss;   // d.H holds the energy weighted density here
        s_wp[warp][e] = Pij;
        hps += HPij * d.S[a];
    }
    hps = group_sum(hps);
    __syncwarp(gmask);
    const double intcut = d.intcut;
    const double sqrtpi = sqrt(kPi);
    double gx[3] = {0, 0, 0};
    // lpc lanes share one Cartesian component pair and split its primitive pairs (everything below is linear in the
    // primitive contributions, so per-lane partial sums are contracted independently and added by the warp reduction)
    const int ncomp = nci * ncj, npp = npi * npj;
    const int lpc = ncomp >= kIntLG ? 1 : kIntLG / ncomp;
    const int ngrp = kIntLG / lpc;
    const int grp = lane / lpc, sub = lane % lpc;
    // One-dimensional moments and their derivatives are tabulated per primitive pair and direction for all exponent pairs of the
    // shell pair (k_sdqh0 does the same), the prefactors once per primitive pair; primitive pairs in chunks that fit the table.
    const int nbj = lj + 1, nab = (li + 1) * nbj;
    const int ppstride = 18 * nab + 1;
    const int chunk = min(npp, kGradTab / ppstride);
    for (int c0 = 0; c0 < ncomp; c0 += ngrp) {
        const int c = c0 + grp;
        const bool on = (grp < ngrp) && (c < ncomp);
        const int mli = on ? c / ncj : 0, mlj = on ? c % ncj : 0;
        const int ix = 6 * (c_lx[c_itt[li] + mli] * nbj + c_lx[c_itt[lj] + mlj]);
        const int iy = 6 * (nab + c_ly[c_itt[li] + mli] * nbj + c_ly[c_itt[lj] + mlj]);
        const int iz = 6 * (2 * nab + c_lz[c_itt[li] + mli] * nbj + c_lz[c_itt[lj] + mlj]);
        const double fi = (li == 2) ? c_dcart[mli] : 1.0, fj = (lj == 2) ? c_dcart[mlj] : 1.0;
        // Cartesian weights: W_c(mli,mlj) = sum_{ii,jj} T_i(mli,ii) W(ii,jj) T_j(mlj,jj)
        double wf = 0.0, wp = 0.0;
        if (on)
        for (int ii = 0; ii < nsi; ++ii) {
            const double ti = (li == 2) ? c_dtrafo[mli][ii] : (ii == mli ? 1.0 : 0.0);
            if (ti == 0.0) continue;
            for (int jj = 0; jj < nsj; ++jj) {
                const double tj = (lj == 2) ? c_dtrafo[mlj][jj] : (jj == mlj ? 1.0 : 0.0);
                if (tj == 0.0) continue;
                wf += ti * tj * s_wf[warp][ii * nsj + jj];
                wp += ti * tj * s_wp[warp][ii * nsj + jj];
            }
        }
        double sdq[10], sdqg[3][10];
#pragma unroll
        for (int k = 0; k < 10; ++k) { sdq[k] = 0.0; sdqg[0][k] = sdqg[1][k] = sdqg[2][k] = 0.0; }
        for (int p0 = 0; p0 < npp; p0 += chunk) {
            const int nch = min(chunk, npp - p0);
            __syncwarp(gmask);
            for (int q = lane; q < nch; q += kIntLG) {   // prefactors (negative: screened out)
                const int ij = s_ipjp[warp][p0 + q], ip = ij & 15, jp = ij >> 4;
                const double alpi = s_alp[warp][0][ip], alpj = s_alp[warp][1][jp];
                double est = alpi * alpj * rij2;
                double kab = -1.0;
                if (!(est > intcut * (alpi + alpj))) {
                    if (c0 == 0) cnt_add(d, blockIdx.y, 3, 1);
                    const double ab = 1.0 / (alpi + alpj);
                    est = est * ab;
                    const double sq = sqrtpi * sqrt(ab);
                    kab = exp(-est) * (sq * sq * sq);
                }
                s_kab[warp][q] = kab;
            }
            __syncwarp(gmask);
            for (int e = lane; e < 3 * nch; e += kIntLG) {   // one lane per (primitive pair, direction)
                const int q = e / 3, dim = e - 3 * q;
                if (s_kab[warp][q] < 0.0) continue;
                const int ij = s_ipjp[warp][p0 + q], ip = ij & 15, jp = ij >> 4;
                const double alpi = s_alp[warp][0][ip], alpj = s_alp[warp][1][jp];
                const double ab = 1.0 / (alpi + alpj), gm = 0.5 * ab;
                double t[9];
#pragma unroll
                for (int k = 0; k < 9; ++k) t[k] = olapp(k, gm);
                const double xi = ri[dim], xj = rj[dim];
                const double rp = (alpi * xi + alpj * xj) * ab;
                // operator origin on atom j (rc = rj), intgrad.f90:1020
                moments_grad_table_dyn(li, lj, alpi, rp - xi, rp - xj, rp - xj, t, tab + q * ppstride + 6 * dim * nab);
            }
            __syncwarp(gmask);
            if (on)
            for (int pp = p0 + sub; pp < p0 + nch; pp += lpc) {
                const int q = pp - p0;
                const double kab = s_kab[warp][q];
                if (kab < 0.0) continue;
                const int ij = s_ipjp[warp][pp], ip = ij & 15, jp = ij >> 4;
                const double ci = s_cont[warp][0][ip] * fi;
                const double cc = kab * (s_cont[warp][1][jp] * fj) * ci;
                double val[3][3], gra[3][3];
                {
                    const double* bx = tab + q * ppstride + ix;
                    const double* by = tab + q * ppstride + iy;
                    const double* bz = tab + q * ppstride + iz;
#pragma unroll
                    for (int u = 0; u < 3; ++u) {
                        val[0][u] = bx[u]; gra[0][u] = bx[3 + u];
                        val[1][u] = by[u]; gra[1][u] = by[3 + u];
                        val[2][u] = bz[u]; gra[2][u] = bz[3 + u];
                    }
                }
                constexpr int ord[10][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {2, 0, 0}, {0, 2, 0}, {0, 0, 2}, {1, 1, 0}, {1, 0, 1}, {0, 1, 1}};
#pragma unroll
                for (int k = 0; k < 10; ++k) {
                    const int ox = ord[k][0], oy = ord[k][1], oz = ord[k][2];
                    sdq[k] += val[0][ox] * val[1][oy] * val[2][oz] * cc;
                    sdqg[0][k] += gra[0][ox] * val[1][oy] * val[2][oz] * cc;
                    sdqg[1][k] += val[0][ox] * gra[1][oy] * val[2][oz] * cc;
                    sdqg[2][k] += val[0][ox] * val[1][oy] * gra[2][oz] * cc;
                }
            }
        }
        if (!on) continue;
        // multipole potentials of the two atoms: loaded here, not before the primitive loop (18 doubles less to keep in registers)
        double vdi[3], vdj[3], vqi[6], vqj[6];
#pragma unroll
        for (int k = 0; k < 3; ++k) { vdi[k] = d.vd[3 * (m.at0 + iat) + k]; vdj[k] = d.vd[3 * (m.at0 + jat) + k]; }
#pragma unroll
        for (int k = 0; k < 6; ++k) { vqi[k] = d.vq[6 * (m.at0 + iat) + k]; vqj[k] = d.vq[6 * (m.at0 + jat) + k]; }
        // shiftintg (intgrad.f90:703-728): operator moved from j to i, then contract
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double* g = sdqg[k];
            double g11 = g[1] - rij[0] * g[0];
            double g12 = g[2] - rij[1] * g[0];
            double g13 = g[3] - rij[2] * g[0];
            double g14 = g[4] - 2 * rij[0] * g[1] + rij[0] * rij[0] * g[0];
            double g15 = g[5] - 2 * rij[1] * g[2] + rij[1] * rij[1] * g[0];
            double g16 = g[6] - 2 * rij[2] * g[3] + rij[2] * rij[2] * g[0];
            double g17 = g[7] - rij[0] * g[2] - rij[1] * g[1] + rij[0] * rij[1] * g[0];
            double g18 = g[8] - rij[0] * g[3] - rij[2] * g[1] + rij[0] * rij[2] * g[0];
            double g19 = g[9] - rij[1] * g[3] - rij[2] * g[2] + rij[1] * rij[2] * g[0];
            if (k == 0) {
                g11 -= sdq[0];
                g14 += -2 * sdq[1] + 2 * rij[0] * sdq[0];
                g17 += -sdq[2] + rij[1] * sdq[0];
                g18 += -sdq[3] + rij[2] * sdq[0];
            } else if (k == 1) {
                g12 -= sdq[0];
                g15 += -2 * sdq[2] + 2 * rij[1] * sdq[0];
                g17 += -sdq[1] + rij[0] * sdq[0];
                g19 += -sdq[3] + rij[2] * sdq[0];
            } else {
                g13 -= sdq[0];
                g16 += -2 * sdq[3] + 2 * rij[2] * sdq[0];
                g18 += -sdq[1] + rij[0] * sdq[0];
                g19 += -sdq[2] + rij[1] * sdq[0];
            }
            const double dtmp = (g11 * vdi[0] + g12 * vdi[1] + g13 * vdi[2]) + (g[1] * vdj[0] + g[2] * vdj[1] + g[3] * vdj[2]);
            const double qtmp = (g14 * vqi[0] + g15 * vqi[1] + g16 * vqi[2] + g17 * vqi[3] + g18 * vqi[4] + g19 * vqi[5]) +
                                (g[4] * vqj[0] + g[5] * vqj[1] + g[6] * vqj[2] + g[7] * vqj[3] + g[8] * vqj[4] + g[9] * vqj[5]);
            gx[k] += g[0] * wf + wp * (dtmp + qtmp);
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) gx[k] = group_sum(gx[k]);
    if (lane == 0) {
        // distance dependent polynomial, grad_core.f90:40-72
        const AtomPar& a = d.atoms[m.at0 + iat];
        const AtomPar& b = d.atoms[m.at0 + jat];
        const double r = sqrt(rij2) / (a.rad + b.rad);
        const double rf1 = 1.0 + 0.01 * si.poly * sqrt(r), rf2 = 1.0 + 0.01 * sj.poly * sqrt(r);
        const double shpoly = rf1 * rf2;
        const double dum = 0.5 * (0.01 * si.poly * sqrt(r) * rf2 + rf1 * 0.01 * sj.poly * sqrt(r)) / rij2;
        for (int k = 0; k < 3; ++k) out[k] = gx[k] + 2 * hps * (dum * rij[k]) / shpoly;
        out[3] = hps / hav;
    }
}

// accumulate shell-pair results per atom in a fixed order; adds the on-site dE/dCN term
// (hamiltonian.F90:741-771); thread per atom
__global__ void k_grad_h0_collect(Dev d) {
    const MolDesc m = d.mol[blockIdx.y];
    const int tg = blockIdx.x * blockDim.x + threadIdx.x, sl = tg % kAS;
    const bool active = tg / kAS < m.nat;
    const int i = active ? tg / kAS : 0;
    const AtomPar& a = d.atoms[m.at0 + i];
    const double* gp = d.gpair + 4 * d.pair0[blockIdx.y];
    double g[3] = {0, 0, 0}, dh = 0;
    for (int s = a.sh0; s < (active ? a.sh0 + a.nsh : 0); ++s) {
        const ShellPar& sp = d.shells[m.sh0 + s];
        const double dse = -sp.kcn;
        for (int t = sl; t < m.nsh; t += kAS) {
            if (t >= a.sh0 && t < a.sh0 + a.nsh) continue;
            const int hi = s > t ? s : t, lo = s > t ? t : s;
            const double* o = gp + 4 * ((long long)hi * (hi + 1) / 2 + lo);
            const double sgn = s > t ? 1.0 : -1.0;
            g[0] += sgn * o[0]; g[1] += sgn * o[1]; g[2] += sgn * o[2];
            dh += o[3] * dse;
        }
        if (d.srank == 0 && sl == 0)   // on-site term: counted once when the shell pairs are split over ranks
            for (int k = 0; k < c_nsph[sp.l]; ++k) {
                const long long ao = sp.ao0 + k;
                dh += d.P[m.mat0 + ao + (long long)m.nao * ao] * dse * kEVtoAU;
            }
    }
    dh = slice_sum(dh);
    for (int k = 0; k < 3; ++k) g[k] = slice_sum(g[k]);
    if (!active || sl) return;
    for (int k = 0; k < 3; ++k) d.grad[3 * (m.at0 + i) + k] += g[k];
    d.dhdcn[m.at0 + i] = dh;
}

// anisotropic electrostatics gradient (aniso_grad, aespot.F90:1070-1165) + dradcn factor; thread per atom
__global__ void k_grad_aes(Dev d) {
    const MolDesc m = d.mol[blockIdx.y];
    const int t = blockIdx.x * blockDim.x + threadIdx.x, sl = t % kAS;
    const bool active = t / kAS < m.nat && shard_owns_atom(d, t / kAS);
    const int i = active ? t / kAS : 0;
    const double* xyz = d.xyz + 3 * m.at0;
    const double* q = d.q + m.at0;
    const double* dipm = d.dipm + 3 * m.at0;
    const double* qp = d.qp + 6 * m.at0;
    const double* radcn = d.radcn + m.at0;
    const double q1 = q[i];
    const double dip[3] = {dipm[3 * i], dipm[3 * i + 1], dipm[3 * i + 2]};
    double qp1[6];
    for (int k = 0; k < 6; ++k) qp1[k] = qp[6 * i + k];
    double gi[3] = {0, 0, 0}, tmp2 = 0;
    for (int j = sl; j < (active ? m.nat : 0); j += kAS) {
        if (j == i) continue;
        const double rij[3] = {xyz[3 * j] - xyz[3 * i], xyz[3 * j + 1] - xyz[3 * i + 1], xyz[3 * j + 2] - xyz[3 * i + 2]};
        const double r2 = rij[0] * rij[0] + rij[1] * rij[1] + rij[2] * rij[2];
        const double rabi = 1.0 / sqrt(r2);
        const double rab = 0.50 * (radcn[i] + radcn[j]);
        const double damp1 = 1.0 / (1.0 + 6.0 * pow(rab * rabi, 3.0));
        const double ddamp1 = -3.0 * rabi * (damp1 * damp1 - damp1);
        const double dgab3 = (-3.0 * rabi * pow(rabi, 3.0)) * damp1 + ddamp1 * pow(rabi, 3.0);
        const double damp2 = 1.0 / (1.0 + 6.0 * pow(rab * rabi, 4.0));
        const double ddamp2 = -4.0 * rabi * (damp2 * damp2 - damp2);
        const double dgab5 = (-5.0 * rabi * pow(rabi, 5.0)) * damp2 + ddamp2 * pow(rabi, 5.0);
        const double g3 = d.gab3[m.nn0 + j + (long long)m.nat * i], g5 = d.gab5[m.nn0 + j + (long long)m.nat * i];
        double ed = 0, edd = 0, eq = 0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            ed += q[j] * dip[k] * rij[k];
            ed -= dipm[3 * j + k] * q1 * rij[k];
            const double tt0 = q1 * dipm[3 * j + k] - q[j] * dip[k];
            gi[k] += g3 * tt0;
            double ddm2 = 0, ddm3a = 0, ddm3b = 0, qqa = 0, qqb = 0;
#pragma unroll
            for (int l = 0; l < 3; ++l) {
                const int kl = lin0(l, k);
                const double t2 = rij[l] * rij[k], tt3 = 3.0 * t2;
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
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double dxi = rij[k] * rabi;
            gi[k] -= ed * dgab3 * dxi;
            gi[k] -= (eq + edd) * dgab5 * dxi;
        }
        tmp2 += ed * 3.0 * g3 * (damp1 / rab) * pow(rab * rabi, 3.0);
        tmp2 += (eq + edd) * 4.0 * g5 * (damp2 / rab) * pow(rab * rabi, 4.0);
    }
    tmp2 = slice_sum(tmp2);
    for (int k = 0; k < 3; ++k) gi[k] = slice_sum(gi[k]);
    if (!active || sl) return;
    tmp2 = 3.0 * tmp2;
    const AtomPar& a = d.atoms[m.at0 + i];
    const double t1 = exp(-4.0 * (d.cn[m.at0 + i] - a.valcn - 1.2));
    const double t2 = (5.0 - a.multirad) / (1.0 + 2.0 * t1 + t1 * t1) * 4.0 * t1;
    d.w_aes[m.at0 + i] = -tmp2 * t2;
    for (int k = 0; k < 3; ++k) d.grad[3 * (m.at0 + i) + k] += gi[k];
}

// D4 reference weights for the gradient (weight_references, dftd4.F90:948-1065); thread per atom
// zw[atom][0..4][7] = zetavec, zerovec, zetadcn, zerodcn, (unused zetadq)
__global__ void k_d4_weights(Dev d) {
    const MolDesc m = d.mol[blockIdx.y];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m.nat) return;
    const AtomPar& a = d.atoms[m.at0 + i];
    const RefPar* r0 = d.refs + m.ref0 + a.ref0;
    const double cn = d.cn_cov[m.at0 + i], q = d.q[m.at0 + i];
    const double zi = a.d4_zeff, gi = 2.0 * a.d4_gam;
    double norm = 0, dnorm = 0, maxcn = -1e300;
    for (int r = 0; r < a.nref; ++r) {
        for (int c = 1; c <= r0[r].ncount; ++c) {
            const double twf = c * 6.0, g = cngw(twf, cn, r0[r].covcn);
            norm += g;
            dnorm += 2 * twf * (r0[r].covcn - cn) * g;
        }
        maxcn = fmax(maxcn, r0[r].covcn);
    }
    norm = norm > 1e-80 ? 1.0 / norm : 0.0;
    double* zw = d.zw + 35 * (long long)(m.at0 + i);
    for (int r = 0; r < 7; ++r) {
        double zv = 0, z0 = 0, zd = 0, z0d = 0;
        if (r < a.nref) {
            double expw = 0, expd = 0;
            for (int c = 1; c <= r0[r].ncount; ++c) {
                const double twf = c * 6.0, g = cngw(twf, cn, r0[r].covcn);
                expw += g;
                expd += 2 * twf * (r0[r].covcn - cn) * g;
            }
            double gwk = expw * norm;
            if (norm == 0.0) gwk = (fabs(maxcn - r0[r].covcn) < 1e-12) ? 1.0 : 0.0;
            const double zq = d4_zeta(3.0, gi, r0[r].refq + zi, q + zi);
            const double z00 = d4_zeta(3.0, gi, r0[r].refq + zi, zi);
            const double dgwk = expd * norm - expw * dnorm * norm * norm;
            zv = zq * gwk; z0 = z00 * gwk; zd = zq * dgwk; z0d = z00 * dgwk;
        }
        zw[r] = zv; zw[7 + r] = z0; zw[14 + r] = zd; zw[21 + r] = z0d;
    }
}

// atomic C6 and dC6/dCN for charged (c6, dc6) and neutral (c60, dc60) weights (get_atomic_c6,
// dftd4.F90:1069-1147); thread per ordered pair (i,j): c6[i + n*j], dc6[i + n*j] = dC6_ij/dCN_i
__global__ void k_d4_c6(Dev d) {
    const MolDesc m = d.mol[blockIdx.y];
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)m.nat * m.nat) return;
    const int i = (int)(t % m.nat), j = (int)(t / m.nat);
    const AtomPar& a = d.atoms[m.at0 + i];
    const AtomPar& b = d.atoms[m.at0 + j];
    const double* zi = d.zw + 35 * (long long)(m.at0 + i);
    const double* zj = d.zw + 35 * (long long)(m.at0 + j);
    const double* ref = d.c6ref + ((size_t)a.etype * d.ntype + b.etype) * 49;
    double c6 = 0, dc6 = 0, c60 = 0, dc60 = 0;
    for (int ir = 0; ir < a.nref; ++ir)
        for (int jr = 0; jr < b.nref; ++jr) {
            const double r = ref[ir * 7 + jr];
            c6 += zi[ir] * zj[jr] * r;
            dc6 += zi[14 + ir] * zj[jr] * r;
            c60 += zi[7 + ir] * zj[7 + jr] * r;
            dc60 += zi[21 + ir] * zj[7 + jr] * r;
        }
    d.c6[m.nn0 + t] = c6; d.dc6[m.nn0 + t] = dc6; d.c60[m.nn0 + t] = c60; d.dc60[m.nn0 + t] = dc60;
}

// two-body D4 gradient (disp_gradient_latp, cutoff 60 bohr) + Coulomb derivative
// (getCoulombDerivsCluster + gemv); thread per atom, loops over partners
__global__ void k_grad_pairs(Dev d) {
    const MolDesc m = d.mol[blockIdx.y];
    const int t = blockIdx.x * blockDim.x + threadIdx.x, sl = t % kAS;
    const bool active = t / kAS < m.nat && shard_owns_atom(d, t / kAS);
    const int i = active ? t / kAS : 0;
    const AtomPar ai = d.atoms[m.at0 + i];
    const double* xi = d.xyz + 3 * (m.at0 + i);
    double g[3] = {0, 0, 0}, dE = 0;
    for (int j = sl; j < (active ? m.nat : 0); j += kAS) {
        if (j == i) continue;
        const AtomPar& aj = d.atoms[m.at0 + j];
        const double* xj = d.xyz + 3 * (m.at0 + j);
        const double rij[3] = {xi[0] - xj[0], xi[1] - xj[1], xi[2] - xj[2]};
        const double r2 = rij[0] * rij[0] + rij[1] * rij[1] + rij[2] * rij[2];
        if (!(r2 > 3600.0 || r2 < 1.0e-10)) {
            cnt_add(d, blockIdx.y, 8, 1);
            const double r4r2ij = 3 * ai.r4r2 * aj.r4r2;
            const double r0 = 0.52 * sqrt(r4r2ij) + 5.0;
            const double t6 = 1. / (r2 * r2 * r2 + pow(r0, 6.0)), t8 = 1. / (r2 * r2 * r2 * r2 + pow(r0, 8.0));
            const double d6 = -6 * r2 * r2 * t6 * t6, d8 = -8 * r2 * r2 * r2 * t8 * t8;
            const double disp = 1.0 * t6 + 2.7 * r4r2ij * t8;
            const double ddisp = 1.0 * d6 + 2.7 * r4r2ij * d8;
            const double c6ij = d.c6[m.nn0 + i + (long long)m.nat * j];
            for (int k = 0; k < 3; ++k) g[k] += -c6ij * ddisp * rij[k];
            dE -= d.dc6[m.nn0 + i + (long long)m.nat * j] * disp;
        }
        // Coulomb: g_i += sum_{sh_i, sh_j} q_i q_j * (-dG), dG = -vec r^(gExp-2) g1 g1^(1/gExp), vec = r_j - r_i
        const double r1 = sqrt(r2);
        double acc = 0;
        for (int s = 0; s < ai.nsh; ++s) {
            const double hs = d.shells[m.sh0 + ai.sh0 + s].hard, qs = d.qsh[m.sh0 + ai.sh0 + s];
            for (int t = 0; t < aj.nsh; ++t) {
                const double gij = 0.5 * (hs + d.shells[m.sh0 + aj.sh0 + t].hard);
                const double g1 = 1.0 / (pow(r1, 2.0) + pow(gij, -2.0));
                acc += qs * d.qsh[m.sh0 + aj.sh0 + t] * (g1 * pow(g1, 0.5));
            }
        }
        // -dG*q_i*q_j with vec = -rij :  -(-(-rij) * f) = -rij*f
        for (int k = 0; k < 3; ++k) g[k] -= rij[k] * acc;
    }
    dE = slice_sum(dE);
    for (int k = 0; k < 3; ++k) g[k] = slice_sum(g[k]);
    if (!active || sl) return;
    for (int k = 0; k < 3; ++k) d.grad[3 * (m.at0 + i) + k] += g[k];
    d.dEdcn[m.at0 + i] = dE;
}

// gradient of the point charge embedding (pcem_grad_gfn2, src/embedding.f90:329-365) with the mixed shell charges.
// Two passes without atomics: thread per atom for the molecular gradient, thread per point charge for theirs.
__global__ void k_grad_pcem_atoms(Dev d) {
    const MolDesc m = d.mol[blockIdx.y];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m.nat || !shard_owns_atom(d, i)) return;
    const AtomPar ai = d.atoms[m.at0 + i];
    const double* xi = d.xyz + 3 * (m.at0 + i);
    double g[3] = {0, 0, 0};
    for (int s = 0; s < ai.nsh; ++s) {
        const double hs = d.shells[m.sh0 + ai.sh0 + s].hard, qs = d.qsh[m.sh0 + ai.sh0 + s];
        for (long long j = d.pc0[blockIdx.y]; j < d.pc0[blockIdx.y + 1]; ++j) {
            const double dr[3] = {xi[0] - d.pc_xyz[3 * j], xi[1] - d.pc_xyz[3 * j + 1], xi[2] - d.pc_xyz[3 * j + 2]};
            double rr = 0.5 * (hs + d.pc_gam[j]);
            rr = 1.0 / (rr * rr);
            const double yy = qs * d.pc_q[j] * pow(dr[0] * dr[0] + dr[1] * dr[1] + dr[2] * dr[2] + rr, -1.5);
            for (int k = 0; k < 3; ++k) g[k] -= dr[k] * yy;
        }
    }
    for (int k = 0; k < 3; ++k) d.grad[3 * (m.at0 + i) + k] += g[k];
}
__global__ void k_grad_pcem_charges(Dev d) {
    const MolDesc m = d.mol[blockIdx.y];
    const long long j = d.pc0[blockIdx.y] + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= d.pc0[blockIdx.y + 1]) return;
    double g[3] = {0, 0, 0};
    for (int s = 0; s < m.nsh; ++s) {
        const ShellPar& sp = d.shells[m.sh0 + s];
        const double* xi = d.xyz + 3 * (m.at0 + sp.atom);
        const double dr[3] = {xi[0] - d.pc_xyz[3 * j], xi[1] - d.pc_xyz[3 * j + 1], xi[2] - d.pc_xyz[3 * j + 2]};
        double rr = 0.5 * (sp.hard + d.pc_gam[j]);
        rr = 1.0 / (rr * rr);
        const double yy = d.qsh[m.sh0 + s] * d.pc_q[j] * pow(dr[0] * dr[0] + dr[1] * dr[1] + dr[2] * dr[2] + rr, -1.5);
        for (int k = 0; k < 3; ++k) g[k] += dr[k] * yy;
    }
    for (int k = 0; k < 3; ++k) d.pc_grad[3 * j + k] = g[k];
}

// Axilrod-Teller-Muto three-body term with neutral reference weights (atm_gradient_latp,
// deriv_atm_triple): one CTA per atom a, threads over the partner pairs (b, c < b), flattened; each triple is
// visited from its three atoms, every CTA accumulates only its own gradient / dE/dCN and reduces in a fixed order
// (no atomics).  dynamic shared memory: nat doubles (row a of dC6/dCN, strided in HBM).
__global__ void __launch_bounds__(128) k_grad_atm(Dev d) {
    const MolDesc m = d.mol[blockIdx.y];
    const int a = blockIdx.x;
    if (a >= m.nat || !shard_owns_atom(d, a)) return;
    extern __shared__ double s_dc6a[];
    __shared__ double red[32];
    const double* xyz = d.xyz + 3 * m.at0;
    const double ra[3] = {xyz[3 * a], xyz[3 * a + 1], xyz[3 * a + 2]};
    const double r4a = d.atoms[m.at0 + a].r4r2;
    const double* c60 = d.c60 + m.nn0;
    const double* dc60 = d.dc60 + m.nn0;
    const int n = m.nat;
    for (int c = threadIdx.x; c < n; c += blockDim.x) s_dc6a[c] = dc60[a + (long long)n * c];
    __syncthreads();
    const double* c6a = c60 + (long long)n * a;     // C6(a, .) ; the matrix is symmetric
    double g[3] = {0, 0, 0}, dcn = 0, e3 = 0;
    // partner pairs (b, c), c < b, numbered b (b - 1) / 2 + c and dealt to the threads round-robin: every thread gets the same
    // number of pairs whatever the size of the molecule (a loop over b per thread leaves 2/3 of the CTA idle for 47 atoms)
    const long long npair = (long long)n * (n - 1) / 2;
    int b = 0, c = 0;
    {
        const long long p0 = threadIdx.x;
        b = (int)(0.5 * (1.0 + sqrt(1.0 + 8.0 * (double)p0)));
        while ((long long)b * (b - 1) / 2 > p0) --b;
        while ((long long)(b + 1) * b / 2 <= p0) ++b;
        c = (int)(p0 - (long long)b * (b - 1) / 2);
    }
    int bc = -1;                                   // atom whose pair quantities with a are in the registers below
    double rij[3] = {0, 0, 0}, r2ij = 0, r4b = 0, c6ij = 0, cij = 0, dc6ij = 0;
    const double* c6b = c60;
    for (long long p = threadIdx.x; p < npair; p += blockDim.x) {
        if (b != a && c != a) {
            if (b != bc) {
                bc = b;
                rij[0] = xyz[3 * b] - ra[0]; rij[1] = xyz[3 * b + 1] - ra[1]; rij[2] = xyz[3 * b + 2] - ra[2];
                r2ij = rij[0] * rij[0] + rij[1] * rij[1] + rij[2] * rij[2];
                r4b = d.atoms[m.at0 + b].r4r2;
                c6ij = c6a[b];
                cij = 0.52 * sqrt(3.0 * r4a * r4b) + 5.0;
                dc6ij = s_dc6a[b];
                c6b = c60 + (long long)n * b;
            }
            const double rik[3] = {xyz[3 * c] - ra[0], xyz[3 * c + 1] - ra[1], xyz[3 * c + 2] - ra[2]};
            const double r2ik = rik[0] * rik[0] + rik[1] * rik[1] + rik[2] * rik[2];
            const double rjk[3] = {rik[0] - rij[0], rik[1] - rij[1], rik[2] - rij[2]};
            const double r2jk = rjk[0] * rjk[0] + rjk[1] * rjk[1] + rjk[2] * rjk[2];
            if (r2ij <= 1600.0 && r2ik <= 1600.0 && r2jk <= 1600.0) {
                cnt_add(d, blockIdx.y, 9, 1);
                const double r4c = d.atoms[m.at0 + c].r4r2;
                const double c6ik = c6a[c], c6jk = c6b[c];
                const double cik = 0.52 * sqrt(3.0 * r4a * r4c) + 5.0, cjk = 0.52 * sqrt(3.0 * r4b * r4c) + 5.0;
                const double c9 = -sqrt(c6ij * c6ik * c6jk);
                const double ccc1 = cij * cjk * cik;
                const double rrr2 = r2ij * r2jk * r2ik, rrr1 = sqrt(rrr2), rrr3 = rrr1 * rrr2;
                const double ang_fact = 0.375 / (rrr2 * rrr3);
                const double ang = ang_fact * (r2ij + r2jk - r2ik) * (r2ij - r2jk + r2ik) * (-r2ij + r2jk + r2ik) + 1.0 / rrr3;
                const double crr = ccc1 / rrr1, crr2 = crr * crr;
                const double cralp = crr2 * crr2 * crr * cbrt(crr);   // (c / r)^(16/3) without the general pow
                const double fdmp = 1.0 / (1.0 + 6.0 * cralp);
                const double dfdmp = -(2.0 * 16 * cralp) * fdmp * fdmp;
                const double dE = -fdmp * ang * c9;
                const double dang1 = -ang_fact * (r2ij * r2ij * r2ij + r2ij * r2ij * (r2jk + r2ik) +
                                                  r2ij * (3.0 * r2jk * r2jk + 2.0 * r2jk * r2ik + 3.0 * r2ik * r2ik) -
                                                  5.0 * (r2jk - r2ik) * (r2jk - r2ik) * (r2jk + r2ik));
                const double dGr1 = (-dang1 * c9 * fdmp + dfdmp * c9 * ang) / r2ij;
                const double dang2 = -ang_fact * (r2ik * r2ik * r2ik + r2ik * r2ik * (r2jk + r2ij) +
                                                  r2ik * (3.0 * r2jk * r2jk + 2.0 * r2jk * r2ij + 3.0 * r2ij * r2ij) -
                                                  5.0 * (r2jk - r2ij) * (r2jk - r2ij) * (r2jk + r2ij));
                const double dGr2 = (-dang2 * c9 * fdmp + dfdmp * c9 * ang) / r2ik;
                const double scale = 5.0;   // s9
                for (int k = 0; k < 3; ++k) g[k] += (-dGr1 * rij[k] - dGr2 * rik[k]) * scale;
                e3 += dE * scale / 3.0;
                dcn += -ang * fdmp * (0.5 * c9 * (dc6ij / c6ij + s_dc6a[c] / c6ik)) * scale;
            }
        }
        c += blockDim.x;
        while (c >= b) { c -= b; ++b; }
    }
    for (int k = 0; k < 3; ++k) g[k] = block_sum(g[k], red);
    dcn = block_sum(dcn, red);
    e3 = block_sum(e3, red);
    if (threadIdx.x == 0) {
        for (int k = 0; k < 3; ++k) d.grad[3 * (m.at0 + a) + k] += g[k];
        d.dEdcn[m.at0 + a] += dcn;
        d.e_atom[4 * (m.at0 + a) + 2] = e3;
    }
}

// CN chain rule: g_i += sum_j dCN-pair(i,j) * (v_i + v_j) for the gfn CN (v = dE/dCN from H0 + AES radii)
// and the covalent CN (v = D4 dE/dCN); replaces the dense gemv with dcndr (scf_module.F90:726, dftd4.F90:1723)
__global__ void k_grad_cn(Dev d) {
    const MolDesc m = d.mol[blockIdx.y];
    const int t = blockIdx.x * blockDim.x + threadIdx.x, sl = t % kAS;
    const bool active = t / kAS < m.nat && shard_owns_atom(d, t / kAS);
    const int i = active ? t / kAS : 0;
    const AtomPar ai = d.atoms[m.at0 + i];
    const double* xi = d.xyz + 3 * (m.at0 + i);
    const double vi = d.dhdcn[m.at0 + i] + d.w_aes[m.at0 + i], wi = d.dEdcn[m.at0 + i];
    const double sqrtpi = sqrt(kPi);
    double g[3] = {0, 0, 0};
    for (int j = sl; j < (active ? m.nat : 0); j += kAS) {
        if (j == i) continue;
        const AtomPar& aj = d.atoms[m.at0 + j];
        const double* xj = d.xyz + 3 * (m.at0 + j);
        const double rij[3] = {xi[0] - xj[0], xi[1] - xj[1], xi[2] - xj[2]};
        const double r2 = rij[0] * rij[0] + rij[1] * rij[1] + rij[2] * rij[2];
        if (r2 > 1600.0 || r2 < 1.0e-12) continue;
        const double r1 = sqrt(r2), rc = ai.rcov + aj.rcov;
        const double dgfn = dexp_count(10.0, r1, rc) * exp_count(20.0, r1, rc + 2) + exp_count(10.0, r1, rc) * dexp_count(20.0, r1, rc + 2);
        const double dcov = en_den(ai.en, aj.en) * (-7.5 / sqrtpi / rc * exp(-7.5 * 7.5 * (r1 - rc) * (r1 - rc) / (rc * rc)));
        const double vj = d.dhdcn[m.at0 + j] + d.w_aes[m.at0 + j], wj = d.dEdcn[m.at0 + j];
        const double f = (dgfn * (vi + vj) + dcov * (wi + wj)) / r1;
        for (int k = 0; k < 3; ++k) g[k] += f * rij[k];
    }
    for (int k = 0; k < 3; ++k) g[k] = slice_sum(g[k]);
    if (!active || sl) return;
    for (int k = 0; k < 3; ++k) d.grad[3 * (m.at0 + i) + k] += g[k];
}

// Multi-GPU split: exchange buffer xbuf[4][natoms_total] = (dE/dCN of H0: partial sums over this rank's shell pairs |
// AES radius weight | D4 dE/dCN | ATM energy: owned atoms only, zero elsewhere).  One NCCL allreduce(sum) of xbuf turns
// it into the full vectors on every rank; k_shard_unpack writes them back.  thread per atom of the whole batch
__global__ void k_shard_pack(Dev d, double* xbuf, int na) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= na) return;
    // molecule of atom t: batches that are split hold one (or a few) large molecules
    int im = 0;
    while (im + 1 < d.nmol && d.mol[im + 1].at0 <= t) ++im;
    const bool own = shard_owns_atom(d, t - d.mol[im].at0);
    xbuf[t] = d.dhdcn[t];
    xbuf[na + t] = own ? d.w_aes[t] : 0.0;
    xbuf[2 * (long long)na + t] = own ? d.dEdcn[t] : 0.0;
    xbuf[3 * (long long)na + t] = own ? d.e_atom[4 * (long long)t + 2] : 0.0;
}
__global__ void k_shard_unpack(Dev d, const double* xbuf, int na) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= na) return;
    d.dhdcn[t] = xbuf[t];
    d.w_aes[t] = xbuf[na + t];
    d.dEdcn[t] = xbuf[2 * (long long)na + t];
    d.e_atom[4 * (long long)t + 2] = xbuf[3 * (long long)na + t];
}

// Wiberg/Mayer bond orders from PS (get_wiberg / get_unrestricted_wiberg, scc_core.f90:1208-1295);
// thread per atom pair.  PS in d.W (closed shell) or d.W (alpha) + d.Zm (beta).
__global__ void k_wbo(Dev d, int open_path) {
    const MolDesc m = d.mol[blockIdx.y];
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)m.nat * m.nat) return;
    const int i = (int)(t % m.nat), j = (int)(t / m.nat);
    double x = 0;
    if (i != j) {
        const double* xi = d.xyz + 3 * (m.at0 + i);
        const double* xj = d.xyz + 3 * (m.at0 + j);
        const double r2 = (xi[0] - xj[0]) * (xi[0] - xj[0]) + (xi[1] - xj[1]) * (xi[1] - xj[1]) + (xi[2] - xj[2]) * (xi[2] - xj[2]);
        if (r2 < 100.0) {
            cnt_add(d, blockIdx.y, 10, 1);
            const AtomPar& a = d.atoms[m.at0 + i];
            const AtomPar& b = d.atoms[m.at0 + j];
            const double* PS = d.W + m.mat0;
            const double* PSb = d.H + m.mat0;   // beta P S (engine.cu; H is dead after the gradient, Zm keeps the eigenvectors)
            const int n = m.nao;
            for (int k = a.ao0; k < a.ao0 + a.nao; ++k)
                for (int l = b.ao0; l < b.ao0 + b.nao; ++l) {
                    x += PS[k + (long long)n * l] * PS[l + (long long)n * k];
                    if (open_path) x += PSb[k + (long long)n * l] * PSb[l + (long long)n * k];
                }
            if (open_path) x *= 2.0;
        }
    }
    d.wbo[m.nn0 + t] = x;
}

// final reductions: total energy, gradient norm, dipole moment (src/dipole.f90:94-128); one CTA per molecule
__global__ void __launch_bounds__(256) k_finish(Dev d) {
    const int im = blockIdx.x;
    const MolDesc m = d.mol[im];
    MolState& st = d.st[im];
    __shared__ double red[32];
    const int tid = threadIdx.x, nt = blockDim.x;
    double p0 = 0, p1 = 0, p2 = 0;
    for (int i = tid; i < m.nat; i += nt) {
        p0 += d.e_atom[4 * (m.at0 + i) + 0];
        p1 += d.e_atom[4 * (m.at0 + i) + 2];
        for (int k = 0; k < 3; ++k) { const double g = d.grad[3 * (m.at0 + i) + k]; p2 += g * g; }
    }
    const double ep = block_sum(p0, red), embd = block_sum(p1, red), g2 = block_sum(p2, red);
    double dip[3];
    for (int k = 0; k < 3; ++k) {
        double p = 0;
        for (int i = tid; i < m.nat; i += nt) p += d.xyz[3 * (m.at0 + i) + k] * d.atoms[m.at0 + i].zval;
        for (int i = tid; i < m.nao; i += nt) p -= d.aop[12 * (long long)(m.ao0 + i) + 1 + k];
        dip[k] = block_sum(p, red);
    }
    if (tid == 0) {
        st.ep = ep; st.embd = embd; st.gnorm = sqrt(g2);
        st.energy = st.eel + ep + embd;
        for (int k = 0; k < 3; ++k) st.dipole[k] = dip[k];
    }
}

}  // namespace xtbb200
