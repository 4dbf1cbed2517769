// Geometry-dependent setup kernels (once per single point):
//   coordination numbers, CN-shifted levels, damped multipole kernels, Klopman-Ohno matrix,
//   D4 Gaussian weights + self-consistent dispersion matrix, repulsion, EEQ guess charges.
// grid = (tiles, nmol); one thread per output element, no atomics (bitwise reproducible).
#pragma once
#include "common.cuh"

namespace xtbb200 {

// ---- counting functions, src/disp/coordinationnumber.F90:483-560
__device__ __forceinline__ double exp_count(double k, double r, double r0) { return 1.0 / (1.0 + exp(-k * (r0 / r - 1.0))); }
__device__ __forceinline__ double dexp_count(double k, double r, double r0) {
    double e = exp(-k * (r0 / r - 1.0));
    return (-k * r0 * e) / (r * r * ((e + 1.0) * (e + 1.0)));
}
__device__ __forceinline__ double en_den(double eni, double enj) {  // k4*exp(-(|dEN|+k5)^2/k6)
    const double k4 = 4.10451, k5 = 19.08857, k6 = 2 * 11.28174 * 11.28174;
    double a = fabs(eni - enj) + k5;
    return k4 * exp(-(a * a) / k6);
}

// CN(gfn), CN(cov) [coordinationnumber.F90:351-479], ncoord_d4 / ncoord_erf [src/disp/ncoord.f90:469-617]
__global__ void k_cn(Dev d) {
    const MolDesc m = d.mol[blockIdx.y];
    const int t = blockIdx.x * blockDim.x + threadIdx.x, sl = t % kAS;
    const bool active = t / kAS < m.nat;
    const int i = active ? t / kAS : 0;
    const AtomPar ai = d.atoms[m.at0 + i];
    const double* xi = d.xyz + 3 * (m.at0 + i);
    double cg = 0, cc = 0, c4 = 0, ce = 0;
    for (int j = sl; j < (active ? m.nat : 0); j += kAS) {
        if (j == i) continue;
        const AtomPar& aj = d.atoms[m.at0 + j];
        const double* xj = d.xyz + 3 * (m.at0 + j);
        double dx = xi[0] - xj[0], dy = xi[1] - xj[1], dz = xi[2] - xj[2];
        double r2 = dx * dx + dy * dy + dz * dz;
        if (r2 > 1600.0) continue;
        cnt_add(d, blockIdx.y, 5, 1);
        double r1 = sqrt(r2);
        if (r2 >= 1.0e-12) {
            cnt_add(d, blockIdx.y, 4, 1);
            double rc = ai.rcov + aj.rcov;
            cg += exp_count(10.0, r1, rc) * exp_count(20.0, r1, rc + 2);
            cc += en_den(ai.en, aj.en) * 0.5 * (1.0 + erf(-7.5 * (r1 - rc) / rc));
        }
        double rco = aj.nc_rcov + ai.nc_rcov;
        double ec = 0.5 * (1.0 + erf(-7.5 * (r1 - rco) / rco));
        ce += ec;
        c4 += en_den(ai.nc_en, aj.nc_en) * ec;
    }
    cg = slice_sum(cg); cc = slice_sum(cc); c4 = slice_sum(c4); ce = slice_sum(ce);
    if (!active || sl) return;
    d.cn[m.at0 + i] = cg; d.cn_cov[m.at0 + i] = cc; d.cn_d4[m.at0 + i] = c4; d.cn_erf[m.at0 + i] = ce;
}

__device__ __forceinline__ double cngw(double wf, double cn, double cnref) {  // dftd4.F90:451-466
    double val = -wf * (cn - cnref) * (cn - cnref);
    return val < -200.0 ? 0.0 : exp(val);
}
__device__ __forceinline__ double d4_zeta(double a, double c, double qref, double qmod) {  // dftd4.F90:348-362
    if (qmod < 0.0) return exp(a);
    return exp(a * (1.0 - exp(c * (1.0 - qref / qmod))));
}
__device__ __forceinline__ double d4_dzeta(double a, double c, double qref, double qmod) {  // dftd4.F90:364-379
    if (qmod < 0.0) return 0.0;
    return -a * c * exp(c * (1.0 - qref / qmod)) * d4_zeta(a, c, qref, qmod) * qref / (qmod * qmod);
}

// per-atom / per-shell / per-reference quantities that depend on the CNs:
//   radcn (aespot.F90:1267-1281), selfE (hamiltonian.F90:98-141), gw (dftd4.F90:588-657)
__global__ void k_cn_dependent(Dev d) {
    const MolDesc m = d.mol[blockIdx.y];
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < m.nat) {
        const AtomPar a = d.atoms[m.at0 + t];
        double t1 = d.cn[m.at0 + t] - a.valcn - 1.2;
        d.radcn[m.at0 + t] = a.multirad + (5.0 - a.multirad) / (1.0 + exp(-4.0 * t1));
    }
    if (t < m.nsh) {
        const ShellPar& s = d.shells[m.sh0 + t];
        d.selfE[m.sh0 + t] = s.level - s.kcn * d.cn[m.at0 + s.atom];
    }
    if (t < m.ndim) {
        const RefPar r = d.refs[m.ref0 + t];
        const AtomPar& a = d.atoms[m.at0 + r.atom];
        double covcn = d.cn_d4[m.at0 + r.atom];
        const RefPar* r0 = d.refs + m.ref0 + a.ref0;
        double norm = 0, maxcn = -1e300;
        for (int ii = 0; ii < a.nref; ++ii) {
            for (int c = 1; c <= r0[ii].ncount; ++c) norm += cngw(c * 6.0, covcn, r0[ii].covcn);
            maxcn = fmax(maxcn, r0[ii].covcn);
        }
        norm = norm > 1e-80 ? 1.0 / norm : 0.0;
        double g = 0;
        for (int c = 1; c <= r.ncount; ++c) g += cngw(c * 6.0, covcn, r.covcn) * norm;
        if (norm == 0.0) g = (fabs(maxcn - r.covcn) < 1e-12) ? 1.0 : 0.0;
        d.gw[m.ref0 + t] = g;
    }
}

// damped multipole kernels gab3/gab5 (aespot.F90:1219-1255, 1317-1341); thread per (i,j)
__global__ void k_gab(Dev d) {
    const MolDesc m = d.mol[blockIdx.y];
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)m.nat * m.nat) return;
    int i = (int)(t / m.nat), j = (int)(t % m.nat);
    double g3 = 0, g5 = 0;
    if (i != j) {
        const double* xi = d.xyz + 3 * (m.at0 + i);
        const double* xj = d.xyz + 3 * (m.at0 + j);
        double dx = xj[0] - xi[0], dy = xj[1] - xi[1], dz = xj[2] - xi[2];
        double rinv = 1.0 / sqrt(dx * dx + dy * dy + dz * dz);
        double rco = 0.5 * (d.radcn[m.at0 + i] + d.radcn[m.at0 + j]);
        double damp3 = 1.0 / (1.0 + 6.0 * pow(rco * rinv, 3.0));
        double damp5 = 1.0 / (1.0 + 6.0 * pow(rco * rinv, 4.0));
        g3 = damp3 * pow(rinv, 3.0);
        g5 = damp5 * pow(rinv, 5.0);
    }
    d.gab3[m.nn0 + t] = g3;
    d.gab5[m.nn0 + t] = g5;
}

// Klopman-Ohno shell-resolved Coulomb matrix, src/coulomb/klopmanohno.F90:257-312 (arithmetic average, gExp=2)
__global__ void k_jmat(Dev d) {
    const MolDesc m = d.mol[blockIdx.y];
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)m.nsh * m.nsh) return;
    int is = (int)(t / m.nsh), js = (int)(t % m.nsh);
    const ShellPar& si = d.shells[m.sh0 + is];
    const ShellPar& sj = d.shells[m.sh0 + js];
    double v;
    if (is == js) v = si.hard;
    else if (si.atom == sj.atom) v = 0.5 * (si.hard + sj.hard);
    else {
        const double* xi = d.xyz + 3 * (m.at0 + si.atom);
        const double* xj = d.xyz + 3 * (m.at0 + sj.atom);
        double dx = xj[0] - xi[0], dy = xj[1] - xi[1], dz = xj[2] - xi[2];
        double r1 = sqrt(dx * dx + dy * dy + dz * dz);
        double gij = 0.5 * (si.hard + sj.hard);
        v = 1.0 / pow(pow(r1, 2.0) + pow(gij, -2.0), 0.5);
    }
    d.jmat[m.ss0 + t] = v;
}

// self-consistent D4 matrix, src/disp/dftd4.F90:660-732 (s6=1, s8=2.7, a1=0.52, a2=5, s10=0)
__global__ void k_dispmat(Dev d) {
    const MolDesc m = d.mol[blockIdx.y];
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)m.ndim * m.ndim) return;
    int k = (int)(t / m.ndim), l = (int)(t % m.ndim);
    const RefPar rk = d.refs[m.ref0 + k], rl = d.refs[m.ref0 + l];
    double v = 0.0;
    if (rk.atom != rl.atom) {
        const AtomPar& a = d.atoms[m.at0 + rk.atom];
        const AtomPar& b = d.atoms[m.at0 + rl.atom];
        const double* xi = d.xyz + 3 * (m.at0 + rk.atom);
        const double* xj = d.xyz + 3 * (m.at0 + rl.atom);
        double dx = xj[0] - xi[0], dy = xj[1] - xi[1], dz = xj[2] - xi[2];
        double r2 = dx * dx + dy * dy + dz * dz;
        double gk = d.gw[m.ref0 + k], gl = d.gw[m.ref0 + l];
        if (r2 <= 72.0 * 72.0 && !(gk * gl < 1.0e-7)) {
            cnt_add(d, blockIdx.y, 7, 1);
            double r4r2ij = 3.0 * a.r4r2 * b.r4r2;
            double cutoff = 0.52 * sqrt(r4r2ij) + 5.0;
            double oor6 = 1.0 / (r2 * r2 * r2 + pow(cutoff, 6.0));
            double oor8 = 1.0 / (r2 * r2 * r2 * r2 + pow(cutoff, 8.0));
            // reference C6 always taken with the higher-index atom first (dftd4.F90:646-651)
            double c6ab = (rk.atom > rl.atom)
                              ? d.c6ref[((size_t)a.etype * d.ntype + b.etype) * 49 + rk.iref * 7 + rl.iref]
                              : d.c6ref[((size_t)b.etype * d.ntype + a.etype) * 49 + rl.iref * 7 + rk.iref];
            double c8 = r4r2ij * c6ab;
            v = gk * gl * (-1.0 * (c6ab * oor6) - 2.7 * (c8 * oor8) - 0.0);
        }
    }
    d.dispmat[m.dd0 + t] = v;
}

// classical repulsion energy + gradient, src/xtb/repulsion.F90:41-158 ; thread per atom
__global__ void k_repulsion(Dev d) {
    const MolDesc m = d.mol[blockIdx.y];
    const int t = blockIdx.x * blockDim.x + threadIdx.x, sl = t % kAS;
    const bool active = t / kAS < m.nat;
    const int i = active ? t / kAS : 0;
    const AtomPar ai = d.atoms[m.at0 + i];
    const double* xi = d.xyz + 3 * (m.at0 + i);
    double e = 0, g[3] = {0, 0, 0};
    for (int j = sl; j < (active ? m.nat : 0); j += kAS) {
        if (j == i) continue;
        const AtomPar& aj = d.atoms[m.at0 + j];
        const double* xj = d.xyz + 3 * (m.at0 + j);
        double rij[3] = {xi[0] - xj[0], xi[1] - xj[1], xi[2] - xj[2]};
        double r2 = rij[0] * rij[0] + rij[1] * rij[1] + rij[2] * rij[2];
        if (r2 > 1600.0 || r2 < 1.0e-8) continue;
        cnt_add(d, blockIdx.y, 6, 1);
        double alpha = sqrt(ai.rep_alpha * aj.rep_alpha);
        double zeff = ai.rep_zeff * aj.rep_zeff;
        double kExp = (ai.Z > 2 || aj.Z > 2) ? 1.5 : 1.0;
        double r1 = sqrt(r2);
        double t16 = pow(r1, kExp), t26 = exp(-alpha * t16);
        double dE = zeff * t26 / r1;
        double f = -(alpha * t16 * kExp + 1.0) * dE / r2;
        for (int k = 0; k < 3; ++k) g[k] += f * rij[k];
        e += 0.5 * dE;
    }
    e = slice_sum(e);
    for (int k = 0; k < 3; ++k) g[k] = slice_sum(g[k]);
    if (!active || sl) return;
    d.e_atom[4 * (m.at0 + i) + 0] = e;
    for (int k = 0; k < 3; ++k) d.grad[3 * (m.at0 + i) + k] = g[k];  // first gradient contribution: overwrite
}

// ---- EEQ guess: src/eeq_model.f90:58-80 (A matrix), 782-870 (X vector, bordered system)
__global__ void k_eeq_build(Dev d) {
    const MolDesc m = d.mol[blockIdx.y];
    int n1 = m.nat + 1;
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)n1 * (n1 + 1)) return;
    double* A = d.eeqA + d.eeq0[blockIdx.y];   // column major n1 x (n1+1); last column = rhs
    int i = (int)(t % n1), j = (int)(t / n1);
    double v;
    if (j == n1) {  // rhs
        if (i == m.nat) v = m.charge;
        else {
            const AtomPar& a = d.atoms[m.at0 + i];
            double cn = d.cn_erf[m.at0 + i];
            double tmp = a.eeq_kappa / (sqrt(cn) + 1e-14);
            v = -a.eeq_en + tmp * cn;
        }
    } else if (i == m.nat || j == m.nat) {
        v = (i == j) ? 0.0 : 1.0;
    } else if (i == j) {
        const AtomPar& a = d.atoms[m.at0 + i];
        v = a.eeq_gam + sqrt(2.0 / kPi) / a.eeq_alpha;
    } else {
        const AtomPar& a = d.atoms[m.at0 + i];
        const AtomPar& b = d.atoms[m.at0 + j];
        const double* xi = d.xyz + 3 * (m.at0 + i);
        const double* xj = d.xyz + 3 * (m.at0 + j);
        double dx = xj[0] - xi[0], dy = xj[1] - xi[1], dz = xj[2] - xi[2];
        double r1 = sqrt(dx * dx + dy * dy + dz * dz);
        double gam = 1.0 / sqrt(a.eeq_alpha * a.eeq_alpha + b.eeq_alpha * b.eeq_alpha);
        v = erf(gam * r1) / r1;
    }
    A[i + (long long)n1 * j] = v;
}

// back substitution on the eliminated augmented system and the charge guess derived from it (one CTA)
__device__ __forceinline__ void eeq_backsub_and_guess(Dev& d, const MolDesc& m, double* A, int n1) {
    // back substitution (serial in k, parallel in rows)
    double* x = A + (long long)n1 * n1;
    for (int k = n1 - 1; k >= 0; --k) {
        if (threadIdx.x == 0) x[k] = x[k] / A[k + (long long)n1 * k];
        __syncthreads();
        double xk = x[k];
        for (int i = threadIdx.x; i < k; i += blockDim.x) x[i] -= A[i + (long long)n1 * k] * xk;
        __syncthreads();
    }
    // charges -> shell charges by reference-occupation fractions; multipoles start at zero
    for (int i = threadIdx.x; i < m.nat; i += blockDim.x) {
        const AtomPar& a = d.atoms[m.at0 + i];
        double qi = x[i];
        d.q[m.at0 + i] = qi;
        for (int s = 0; s < a.nsh; ++s) {
            const ShellPar& sp = d.shells[m.sh0 + a.sh0 + s];
            d.qsh[m.sh0 + a.sh0 + s] = sp.refocc / a.zval * qi;
        }
        for (int k = 0; k < 3; ++k) d.dipm[3 * (m.at0 + i) + k] = 0.0;
        for (int k = 0; k < 6; ++k) d.qp[6 * (m.at0 + i) + k] = 0.0;
    }
}

// Gaussian elimination with partial pivoting on the augmented system (one CTA per molecule).
// Replaces the dsysv call of src/eeq_model.f90:858-862; then iniqshell (scc_core.f90:1475-1506).
__global__ void k_eeq_solve(Dev d) {
    const MolDesc m = d.mol[blockIdx.x];
    int n1 = m.nat + 1, nc = n1 + 1;
    double* A = d.eeqA + d.eeq0[blockIdx.x];
    __shared__ int s_piv;
    __shared__ double s_red[64];
    __shared__ int s_idx[64];
    for (int k = 0; k < n1; ++k) {
        // pivot search in column k
        double best = -1.0; int bi = k;
        for (int i = k + threadIdx.x; i < n1; i += blockDim.x) {
            double v = fabs(A[i + (long long)n1 * k]);
            if (v > best) { best = v; bi = i; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            double ob = __shfl_xor_sync(0xffffffffu, best, o);
            int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if ((threadIdx.x & 31) == 0) { s_red[threadIdx.x >> 5] = best; s_idx[threadIdx.x >> 5] = bi; }
        __syncthreads();
        if (threadIdx.x == 0) {
            int nw = (blockDim.x + 31) >> 5;
            double b = s_red[0]; int ix = s_idx[0];
            for (int w = 1; w < nw; ++w) if (s_red[w] > b || (s_red[w] == b && s_idx[w] < ix)) { b = s_red[w]; ix = s_idx[w]; }
            s_piv = ix;
            if (b == 0.0) d.st[blockIdx.x].fail = 10;
        }
        __syncthreads();
        int p = s_piv;
        if (p != k)
            for (int j = k + threadIdx.x; j < nc; j += blockDim.x) {
                double t = A[k + (long long)n1 * j];
                A[k + (long long)n1 * j] = A[p + (long long)n1 * j];
                A[p + (long long)n1 * j] = t;
            }
        __syncthreads();
        double pinv = 1.0 / A[k + (long long)n1 * k];
        // eliminate: thread over (i, j) with i > k, j > k
        int nr = n1 - k - 1, ncol = nc - k - 1;
        for (long long t = threadIdx.x; t < (long long)nr * ncol; t += blockDim.x) {
            int i = k + 1 + (int)(t % nr), j = k + 1 + (int)(t / nr);
            double f = A[i + (long long)n1 * k] * pinv;
            A[i + (long long)n1 * j] -= f * A[k + (long long)n1 * j];
        }
        __syncthreads();
    }
    eeq_backsub_and_guess(d, m, A, n1);
}

// ---- the same elimination for large molecules, spread over the GPU: per column one pivot kernel (one CTA per
// molecule: search + row swap) and one elimination kernel (thread per trailing element).  Same arithmetic and pivoting
// as k_eeq_solve, so both give identical charges.
__global__ void __launch_bounds__(1024) k_eeq_pivot(Dev d, int k) {
    const MolDesc m = d.mol[blockIdx.x];
    const int n1 = m.nat + 1, nc = n1 + 1;
    if (k >= n1) return;
    double* A = d.eeqA + d.eeq0[blockIdx.x];
    __shared__ int s_piv;
    __shared__ double s_red[32];
    __shared__ int s_idx[32];
    double best = -1.0; int bi = k;
    for (int i = k + threadIdx.x; i < n1; i += blockDim.x) {
        double v = fabs(A[i + (long long)n1 * k]);
        if (v > best) { best = v; bi = i; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        double ob = __shfl_xor_sync(0xffffffffu, best, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { s_red[threadIdx.x >> 5] = best; s_idx[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        int nw = (blockDim.x + 31) >> 5;
        double b = s_red[0]; int ix = s_idx[0];
        for (int w = 1; w < nw; ++w) if (s_red[w] > b || (s_red[w] == b && s_idx[w] < ix)) { b = s_red[w]; ix = s_idx[w]; }
        s_piv = ix;
        if (b == 0.0) d.st[blockIdx.x].fail = 10;
    }
    __syncthreads();
    const int p = s_piv;
    if (p != k)
        for (int j = k + threadIdx.x; j < nc; j += blockDim.x) {
            double t = A[k + (long long)n1 * j];
            A[k + (long long)n1 * j] = A[p + (long long)n1 * j];
            A[p + (long long)n1 * j] = t;
        }
}
__global__ void k_eeq_elim(Dev d, int k) {
    const MolDesc m = d.mol[blockIdx.y];
    const int n1 = m.nat + 1, nc = n1 + 1;
    if (k >= n1) return;
    const int nr = n1 - k - 1, ncol = nc - k - 1;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)nr * ncol) return;
    double* A = d.eeqA + d.eeq0[blockIdx.y];
    const double pinv = 1.0 / A[k + (long long)n1 * k];
    const int i = k + 1 + (int)(t % nr), j = k + 1 + (int)(t / nr);
    const double f = A[i + (long long)n1 * k] * pinv;
    A[i + (long long)n1 * j] -= f * A[k + (long long)n1 * j];
}
__global__ void __launch_bounds__(1024) k_eeq_finish(Dev d) {
    const MolDesc m = d.mol[blockIdx.x];
    eeq_backsub_and_guess(d, m, d.eeqA + d.eeq0[blockIdx.x], m.nat + 1);
}

}  // namespace xtbb200
