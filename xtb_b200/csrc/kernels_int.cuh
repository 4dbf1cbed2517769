// Shell-pair integral kernels: overlap / dipole / quadrupole / H0 (build_SDQH0) and the
// screening masks.  One warp per shell pair, lanes over Cartesian component pairs, contracted
// Gaussian primitives of the two shells staged in shared memory.
//
// Reference: src/xtb/hamiltonian.F90:146-371 (build_SDQH0), src/intgrad.f90:95-111 (olapp),
// 357-471 (horizontal_shift, form_product), 509-555 (multipole_3d), 879-955 (get_multiints),
// 144-231 (dtrf2), src/scc_core.f90:644-680 (h0scal), 728-754 (shellPoly),
// src/aespot.F90:34-85 (setdqlist), src/xtb/hamiltonian.F90:796-848 (count_dpint/qpint).
#pragma once
#include "common.cuh"

namespace xtbb200 {

__constant__ int c_lx[10] = {0, 1, 0, 0, 2, 0, 0, 1, 1, 0};
__constant__ int c_ly[10] = {0, 0, 1, 0, 0, 2, 0, 1, 0, 1};
__constant__ int c_lz[10] = {0, 0, 0, 1, 0, 0, 2, 0, 1, 1};
__constant__ int c_itt[3] = {0, 1, 4};
__constant__ int c_ncart[3] = {1, 3, 6};
__constant__ int c_nsph[3] = {1, 3, 5};
// columns 2..6 of trafo (intgrad.f90:70-86): [cart m][sph a]
__constant__ double c_dtrafo[6][5] = {{0.86602540378443864676, 0.5, 0, 0, 0}, {-0.86602540378443864676, 0.5, 0, 0, 0},
                                      {0, -1.0, 0, 0, 0}, {0, 0, 1, 0, 0}, {0, 0, 0, 1, 0}, {0, 0, 0, 0, 1}};
__constant__ double c_dcart[6] = {1.0, 1.0, 1.0, 1.7320508075688772935, 1.7320508075688772935, 1.7320508075688772935};

// one-dimensional overlap moments, intgrad.f90:95-111
__device__ __forceinline__ double olapp(int l, double gm) {  // gm = 0.5/gamma
    switch (l) {
        case 0: return 1.0;
        case 2: return gm;
        case 4: return gm * gm * 3.0;
        case 6: return gm * gm * gm * 15.0;
        case 8: return gm * gm * gm * gm * 105.0;
        default: return 0.0;
    }
}

// expand (x - A)^l around the product centre: coefficients of (x-P)^k, ae = P - A
__device__ __forceinline__ void shift_poly(double ae, int l, double* c) {
    // c has c[l] = 1 on input (others 0), intgrad.f90:357-380
    switch (l) {
        case 1: c[0] += ae * c[1]; break;
        case 2: c[0] += ae * ae * c[2]; c[1] += 2 * ae * c[2]; break;
        case 3: c[0] += ae * ae * ae * c[3]; c[1] += 3 * ae * ae * c[3]; c[2] += 3 * ae * c[3]; break;
        default: break;
    }
}

// 1D moment integrals <i| (x-C)^{0,1,2} |j> for one Cartesian direction, intgrad.f90:509-538
template <int MAXLI>
__device__ __forceinline__ void moments_1d(int li, int lj, double pa, double pb, double rpc, const double* s1d, double* val) {
    double vi[MAXLI + 1], vj[3], vv[MAXLI + 3];
#pragma unroll
    for (int k = 0; k <= MAXLI; ++k) vi[k] = 0.0;
    vj[0] = vj[1] = vj[2] = 0.0;
    vi[li] = 1.0; vj[lj] = 1.0;
    shift_poly(pa, li, vi);
    shift_poly(pb, lj, vj);
#pragma unroll
    for (int k = 0; k < MAXLI + 3; ++k) {
        double s = 0.0;
#pragma unroll
        for (int a = 0; a <= MAXLI; ++a) {
            int b = k - a;
            if (b >= 0 && b <= 2) s += vi[a] * vj[b];
        }
        vv[k] = s;
    }
    double v0 = 0, v1 = 0, v2 = 0;
#pragma unroll
    for (int l = 0; l < MAXLI + 3; ++l) {
        if (l <= li + lj) {
            v0 += s1d[l] * vv[l];
            v1 += (s1d[l + 1] + rpc * s1d[l]) * vv[l];
            v2 += (s1d[l + 2] + 2 * rpc * s1d[l + 1] + rpc * rpc * s1d[l]) * vv[l];
        }
    }
    val[0] = v0; val[1] = v1; val[2] = v2;
}

// h0scal valence-valence branch (scc_core.f90:644-680), kScale/enScale from gfn2.f90:834-853
__device__ __forceinline__ double h0scal(int li, int lj, double eni, double enj) {
    const double kshell[3] = {1.85, 2.23, 2.23};
    double ks = 0.5 * (kshell[li] + kshell[lj]);
    if ((li == 0 && lj == 2) || (li == 2 && lj == 0)) ks = 2.0;
    if ((li == 1 && lj == 2) || (li == 2 && lj == 1)) ks = 2.0;
    double den = (eni - enj) * (eni - enj);
    double enpoly = (1.0 + 0.02 * den * (1.0 + 0.0 * den));
    return ks * enpoly * 1.0;
}

// The same integrals for all exponent pairs (a <= LI, b <= LJ) of one direction at once: out[(a (LJ + 1) + b) 3 + t].  The
// moment factors of every power of (x - P) are shared; the loops are static (the shell pair, hence LI / LJ, is uniform over
// the warp), the sums run in the order of moments_1d.
template <int LI, int LJ>
__device__ __forceinline__ void moments_table(double pa, double pb, double rpc, const double* s1d, double* out) {
    double m0[LI + LJ + 1], m1[LI + LJ + 1], m2[LI + LJ + 1];
#pragma unroll
    for (int l = 0; l <= LI + LJ; ++l) {
        m0[l] = s1d[l];
        m1[l] = s1d[l + 1] + rpc * s1d[l];
        m2[l] = s1d[l + 2] + 2 * rpc * s1d[l + 1] + rpc * rpc * s1d[l];
    }
    double pi[LI + 1][LI + 1], pj[LJ + 1][LJ + 1];   // (x - A)^a and (x - B)^b in powers of (x - P)
#pragma unroll
    for (int a = 0; a <= LI; ++a) {
#pragma unroll
        for (int k = 0; k <= LI; ++k) pi[a][k] = 0.0;
        pi[a][a] = 1.0;
        if (a == 1) pi[a][0] += pa * pi[a][1];
        else if (a == 2) { pi[a][0] += pa * pa * pi[a][2]; pi[a][1] += 2 * pa * pi[a][2]; }
    }
#pragma unroll
    for (int b = 0; b <= LJ; ++b) {
#pragma unroll
        for (int k = 0; k <= LJ; ++k) pj[b][k] = 0.0;
        pj[b][b] = 1.0;
        if (b == 1) pj[b][0] += pb * pj[b][1];
        else if (b == 2) { pj[b][0] += pb * pb * pj[b][2]; pj[b][1] += 2 * pb * pj[b][2]; }
    }
#pragma unroll
    for (int a = 0; a <= LI; ++a)
#pragma unroll
        for (int b = 0; b <= LJ; ++b) {
            double v0 = 0, v1 = 0, v2 = 0;
#pragma unroll
            for (int l = 0; l <= a + b; ++l) {
                double vv = 0.0;
#pragma unroll
                for (int k = 0; k <= a; ++k)
                    if (l - k >= 0 && l - k <= b) vv += pi[a][k] * pj[b][l - k];
                v0 += m0[l] * vv; v1 += m1[l] * vv; v2 += m2[l] * vv;
            }
            double* o = out + 3 * (a * (LJ + 1) + b);
            o[0] = v0; o[1] = v1; o[2] = v2;
        }
}
__device__ __forceinline__ void moments_table_dyn(int li, int lj, double pa, double pb, double rpc, const double* s1d, double* out) {
    switch (3 * li + lj) {
        case 0: moments_table<0, 0>(pa, pb, rpc, s1d, out); break;
        case 1: moments_table<0, 1>(pa, pb, rpc, s1d, out); break;
        case 2: moments_table<0, 2>(pa, pb, rpc, s1d, out); break;
        case 3: moments_table<1, 0>(pa, pb, rpc, s1d, out); break;
        case 4: moments_table<1, 1>(pa, pb, rpc, s1d, out); break;
        case 5: moments_table<1, 2>(pa, pb, rpc, s1d, out); break;
        case 6: moments_table<2, 0>(pa, pb, rpc, s1d, out); break;
        case 7: moments_table<2, 1>(pa, pb, rpc, s1d, out); break;
        default: moments_table<2, 2>(pa, pb, rpc, s1d, out); break;
    }
}

constexpr int kIntWarps = 4;
// One-dimensional moment integrals are shared by all Cartesian components of a shell pair: for every primitive pair and
// direction there are only (li + 1) (lj + 1) different ones, while a p-p pair has 9 components that each need three.  They are
// tabulated per warp in shared memory ([primitive pair][direction][(a, b)][moment], primitive pairs in chunks that fit), the
// prefactors exp(-est) (pi / (alpha_i + alpha_j))^(3/2) once per primitive pair; the component lanes then only multiply.
constexpr int kIntTab = 432;          // doubles per shell pair
constexpr int kIntMaxPP = 36;         // primitive pairs of a shell pair (6 x 6)

// Half a warp per shell pair: most of the instructions of a shell pair do not depend on its size (pair decoding, shell data,
// shell-pair constants, Cartesian -> spherical transformation, scatter), and 1128 + 1363 of the 3081 shell pairs of a molecule
// like rivaroxaban are s-s / s-p pairs with 1 or 3 components; two pairs per warp share that instruction stream.  Everything
// below is written for groups of kIntLG lanes (`lane` = lane of the group, `slot` = group of the CTA) and synchronises the
// group only (__syncwarp(gmask)), so the two halves of a warp may take different paths.
constexpr int kIntLG = 16;                          // lanes per shell pair
constexpr int kIntPairs = kIntWarps * (32 / kIntLG); // shell pairs per CTA

constexpr size_t kSdqSmem = (size_t)kIntPairs * (360 + kIntTab) * sizeof(double);   // dynamic shared memory of k_sdqh0

// S, D, Q, H0, H0_noovlp for all shell pairs (is >= js).  grid = (ceil(npair/kIntPairs), nmol)
__global__ void __launch_bounds__(kIntWarps * 32, 4) k_sdqh0(Dev d) {
    const MolDesc m = d.mol[blockIdx.y];
    const int warp = threadIdx.x / kIntLG, lane = threadIdx.x % kIntLG;   // "warp": the group of kIntLG lanes of this shell pair
    const unsigned gmask = kIntLG == 32 ? 0xffffffffu : ((1u << kIntLG) - 1u) << ((threadIdx.x & 31) / kIntLG * kIntLG);
    long long npair = (long long)m.nsh * (m.nsh + 1) / 2;
    long long p = shard_pair_cta(d) * kIntPairs + warp;
    extern __shared__ __align__(16) double s_dyn_int[];                    // kSdqSmem bytes: s_blk | s_tabs (more than 48 kB together)
    double (*s_blk)[10][36] = reinterpret_cast<double (*)[10][36]>(s_dyn_int);
    double (*s_tabs)[kIntTab] = reinterpret_cast<double (*)[kIntTab]>(s_dyn_int + kIntPairs * 360);
    __shared__ double s_kab[kIntPairs][kIntMaxPP];
    __shared__ double s_alp[kIntPairs][2][6], s_cont[kIntPairs][2][6];
    __shared__ unsigned char s_ipjp[kIntPairs][kIntMaxPP], s_slot[kIntPairs][kIntMaxPP];   // (ip, jp) of a primitive pair; block slot of a component
    static_assert(kIntTab >= 10 * (kIntLG + 1), "the lane partial sums of the reduction reuse the table");
    double (*s_part)[kIntLG + 1] = reinterpret_cast<double (*)[kIntLG + 1]>(&s_tabs[warp][0]);   // [10][kIntLG + 1], after the table is dead
    double* tab = &s_tabs[warp][0];
    if (p >= npair) return;   // the whole group exits together
    p = d.pair_order[d.pord0[blockIdx.y] + p];
    int is = (int)((sqrt(8.0 * (double)p + 1.0) - 1.0) * 0.5);
    while ((long long)is * (is + 1) / 2 > p) --is;
    while ((long long)(is + 1) * (is + 2) / 2 <= p) ++is;
    int js = (int)(p - (long long)is * (is + 1) / 2);
    const ShellPar& si = d.shells[m.sh0 + is];
    const ShellPar& sj = d.shells[m.sh0 + js];
    const int li = si.l, lj = sj.l, iat = si.atom, jat = sj.atom;
    const int nci = c_ncart[li], ncj = c_ncart[lj], npi = si.nprim, npj = sj.nprim;
    if (lane < 6) { s_alp[warp][0][lane] = si.alp[lane]; s_cont[warp][0][lane] = si.cont[lane]; }
    else if (lane < 12) { s_alp[warp][1][lane - 6] = sj.alp[lane - 6]; s_cont[warp][1][lane - 6] = sj.cont[lane - 6]; }
    const int ncomp = nci * ncj;
    for (int q = lane; q < kIntMaxPP; q += kIntLG) {   // the divisions by the run-time npj / ncj once per shell pair, not per term
        const int ip = q / npj, cc = q / ncj;
        s_ipjp[warp][q] = (unsigned char)(ip | ((q - ip * npj) << 4));
        s_slot[warp][q] = (unsigned char)(cc * 6 + (q - cc * ncj));
    }
    __syncwarp(gmask);
    const double* ri = d.xyz + 3 * (m.at0 + iat);
    const double* rj = d.xyz + 3 * (m.at0 + jat);
    const double rij[3] = {ri[0] - rj[0], ri[1] - rj[1], ri[2] - rj[2]};
    const double rij2 = rij[0] * rij[0] + rij[1] * rij[1] + rij[2] * rij[2];
    const double intcut = d.intcut;
    const double sqrtpi = sqrt(kPi);
    const bool inrange = rij2 <= 2000.0;
    if (lane == 0 && inrange) cnt_add(d, blockIdx.y, 0, 1);
    // Work decomposition: lpc lanes share one Cartesian component pair and split its primitive pairs, so that s-s,
    // s-p and p-p shell pairs (1, 3, 9 components) still occupy most of the warp.
    const int npp = npi * npj;
    const int lpc = ncomp >= kIntLG ? 1 : kIntLG / ncomp;
    const int ngrp = kIntLG / lpc;
    const int grp = lane / lpc, sub = lane % lpc;
    const int nbj = lj + 1, nab = (li + 1) * nbj;
    const int ppstride = 9 * nab + 1;                       // + 1: the lanes of a component group hit different banks
    const int chunk = min(npp, kIntTab / ppstride);         // primitive pairs per table filling (>= 3 for d-d)
    for (int c0 = 0; c0 < ncomp; c0 += ngrp) {
        const int c = c0 + grp;
        const bool on = (grp < ngrp) && (c < ncomp);
        double acc[10];
#pragma unroll
        for (int k = 0; k < 10; ++k) acc[k] = 0.0;
        int ix = 0, iy = 0, iz = 0;
        double fi = 1.0, fj = 1.0;
        if (on) {
            const int mli = c / ncj, mlj = c % ncj;
            ix = 3 * (c_lx[c_itt[li] + mli] * nbj + c_lx[c_itt[lj] + mlj]);
            iy = 3 * (nab + c_ly[c_itt[li] + mli] * nbj + c_ly[c_itt[lj] + mlj]);
            iz = 3 * (2 * nab + c_lz[c_itt[li] + mli] * nbj + c_lz[c_itt[lj] + mlj]);
            fi = (li == 2) ? c_dcart[mli] : 1.0; fj = (lj == 2) ? c_dcart[mlj] : 1.0;
        }
        for (int p0 = 0; inrange && p0 < npp; p0 += chunk) {
            const int nch = min(chunk, npp - p0);
            // ---- prefactors, one lane per primitive pair (negative: screened out)
            for (int q = lane; q < nch; q += kIntLG) {
                const int ij = s_ipjp[warp][p0 + q], ip = ij & 15, jp = ij >> 4;
                const double alpi = s_alp[warp][0][ip], alpj = s_alp[warp][1][jp];
                double est = alpi * alpj * rij2;
                double kab = -1.0;
                if (!(est > intcut * (alpi + alpj))) {
                    if (c0 == 0) cnt_add(d, blockIdx.y, 1, 1);
                    const double ab = 1.0 / (alpi + alpj);
                    est = est * ab;
                    const double sq = sqrtpi * sqrt(ab);
                    kab = exp(-est) * (sq * sq * sq);
                }
                s_kab[warp][q] = kab;
            }
            __syncwarp(gmask);
            // ---- moments <(x - A)^a | (x - C)^{0,1,2} | (x - B)^b>, one lane per (primitive pair, direction)
            for (int e = lane; e < 3 * nch; e += kIntLG) {
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
                moments_table_dyn(li, lj, rp - xi, rp - xj, rp, t, tab + q * ppstride + 3 * dim * nab);
            }
            __syncwarp(gmask);
            // ---- this lane's component: products of three tabulated moments per primitive pair
            if (on) {
                for (int pp = p0 + sub; pp < p0 + nch; pp += lpc) {
                    const int q = pp - p0;
                    const double kab = s_kab[warp][q];
                    if (kab < 0.0) continue;
                    const int ij = s_ipjp[warp][pp], ip = ij & 15, jp = ij >> 4;
                    const double ci = s_cont[warp][0][ip] * fi;
                    const double cc = kab * (s_cont[warp][1][jp] * fj) * ci;
                    const double* vx = tab + q * ppstride + ix;
                    const double* vy = tab + q * ppstride + iy;
                    const double* vz = tab + q * ppstride + iz;
                    const double x0 = vx[0], x1 = vx[1], x2 = vx[2], y0 = vy[0], y1 = vy[1], y2 = vy[2], z0 = vz[0], z1 = vz[1], z2 = vz[2];
                    acc[0] += x0 * y0 * z0 * cc;
                    acc[1] += x1 * y0 * z0 * cc;
                    acc[2] += x0 * y1 * z0 * cc;
                    acc[3] += x0 * y0 * z1 * cc;
                    acc[4] += x2 * y0 * z0 * cc;
                    acc[5] += x0 * y2 * z0 * cc;
                    acc[6] += x0 * y0 * z2 * cc;
                    acc[7] += x1 * y1 * z0 * cc;
                    acc[8] += x1 * y0 * z1 * cc;
                    acc[9] += x0 * y1 * z1 * cc;
                }
            }
            __syncwarp(gmask);
        }
        // fixed-order reduction over the lanes of each component group (deterministic)
#pragma unroll
        for (int k = 0; k < 10; ++k) s_part[k][lane] = acc[k];
        __syncwarp(gmask);
        for (int e = lane; e < ngrp * 10; e += kIntLG) {
            const int g = e / 10, k = e % 10, cc2 = c0 + g;
            if (cc2 < ncomp) {
                double v = 0.0;
                for (int u = 0; u < lpc; ++u) v += s_part[k][g * lpc + u];
                s_blk[warp][k][s_slot[warp][cc2]] = v;
            }
        }
        __syncwarp(gmask);
    }
    // CAO -> SAO and scatter.  Output element (ii, jj) of the spherical block, 10 operators.
    const int nsi = c_nsph[li], nsj = c_nsph[lj];
    const int nao = m.nao;
    double hav = 0.0, shpoly = 1.0;
    const bool offsite = (iat != jat);
    if (offsite) {
        const AtomPar& a = d.atoms[m.at0 + iat];
        const AtomPar& b = d.atoms[m.at0 + jat];
        double hii = d.selfE[m.sh0 + is], hjj = d.selfE[m.sh0 + js];
        double zetaij = sqrt(2 * sqrt(si.zeta * sj.zeta) / (si.zeta + sj.zeta));
        double km = h0scal(li, lj, a.en, b.en);
        hav = 0.5 * km * (hii + hjj) * zetaij;
        double r = sqrt(rij2) / (b.rad + a.rad);
        shpoly = (1.0 + 0.01 * si.poly * sqrt(r)) * (1.0 + 0.01 * sj.poly * sqrt(r));
    }
    for (int e = lane; e < nsi * nsj; e += kIntLG) {
        const int ii = e / nsj, jj = e % nsj;
        const int iao = si.ao0 + ii, jao = sj.ao0 + jj;
        if (!offsite && is == js && jao > iao) continue;
        double out[10];
#pragma unroll
        for (int k = 0; k < 10; ++k) {
            double v = 0.0;
            if (li == 2 && lj == 2) {
                for (int a = 0; a < 6; ++a) {
                    double ta = c_dtrafo[a][ii];
                    if (ta == 0.0) continue;
                    for (int b = 0; b < 6; ++b) v += ta * c_dtrafo[b][jj] * s_blk[warp][k][a * 6 + b];
                }
            } else if (li == 2) {
                for (int a = 0; a < 6; ++a) v += c_dtrafo[a][ii] * s_blk[warp][k][a * 6 + jj];
            } else if (lj == 2) {
                for (int b = 0; b < 6; ++b) v += c_dtrafo[b][jj] * s_blk[warp][k][ii * 6 + b];
            } else {
                v = s_blk[warp][k][ii * 6 + jj];
            }
            out[k] = v;
        }
        const long long a1 = m.mat0 + iao + (long long)nao * jao, a2 = m.mat0 + jao + (long long)nao * iao;
        if (offsite) {
            d.S[a1] = out[0]; d.S[a2] = out[0];
            double h = hav * shpoly * out[0], hn = hav * shpoly;
            d.H0[a1] = h; d.H0[a2] = h; d.H0n[a1] = hn; d.H0n[a2] = hn;
        } else {
            double sv = (iao == jao) ? 1.0 : 0.0;
            double hv = (iao == jao) ? d.selfE[m.sh0 + is] : 0.0;
            d.S[a1] = sv; d.S[a2] = sv; d.H0[a1] = hv; d.H0[a2] = hv; d.H0n[a1] = hv; d.H0n[a2] = hv;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) { d.D[k * d.tot_mat + a1] = out[1 + k]; d.D[k * d.tot_mat + a2] = out[1 + k]; }
#pragma unroll
        for (int k = 0; k < 6; ++k) { d.Q[k * d.tot_mat + a1] = out[4 + k]; d.Q[k * d.tot_mat + a2] = out[4 + k]; }
    }
}

// dipole / quadrupole screening masks + list lengths (aespot.F90:34-85); thread per (i,j)
__global__ void k_dqmask(Dev d) {
    const MolDesc m = d.mol[blockIdx.y];
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int nd = 0, nq = 0;
    if (t < (long long)m.nao * m.nao) {
        int i = (int)(t % m.nao), j = (int)(t / m.nao);
        long long a = m.mat0 + t;
        double thr2 = d.neglect * 1.0e-2;
        double tmp1 = 0, tmp2 = 0;
        for (int k = 0; k < 3; ++k) {
            double dv = d.D[k * d.tot_mat + a], qv = d.Q[k * d.tot_mat + a];
            tmp1 += dv * dv; tmp2 -= qv * qv;
        }
        for (int k = 0; k < 6; ++k) { double qv = d.Q[k * d.tot_mat + a]; tmp2 += 2.0 * qv * qv; }
        uint8_t mk = 0;
        if (tmp1 > thr2) mk |= 1;
        if (tmp2 > thr2) mk |= 2;
        d.mask[a] = mk;
        if (j <= i) { nd = mk & 1; nq = (mk >> 1) & 1; }
    }
    // integer counts: order independent, atomics are exact
    unsigned bd = __ballot_sync(0xffffffffu, nd), bq = __ballot_sync(0xffffffffu, nq);
    if ((threadIdx.x & 31) == 0) {
        if (bd) atomicAdd(&d.st[blockIdx.y].ndp, __popc(bd));
        if (bq) atomicAdd(&d.st[blockIdx.y].nqp, __popc(bq));
    }
}

}  // namespace xtbb200
