// SCC iteration kernels: potentials, Fock assembly, Fermi smearing, populations, energies,
// convergence bookkeeping and Broyden mixing.
//
// Reference: src/scc_core.f90:255-640 (scc), 154-251 (buildIsoAnisotropicH1), 983-1064 (fermismear),
// 1408-1444 (mpopsh, qsh2qat), 686-722 (electro); src/xtb/coulomb.f90:89-142; src/xtb/thirdorder.f90:100-164;
// src/aespot.F90:289-427 (mmompop), 569-662 (aniso_electro), 727-848 (setvsdq), 861-952 (setdvsdq);
// src/disp/dftd4.F90:735-864 (disppot, edisp_scc); src/broyden.f90:37-184.
#pragma once
#include "common.cuh"
#include "kernels_geom.cuh"

namespace xtbb200 {

// initial guess without EEQ: q = charge/n ("sad"-like, test/unit/test_gfn2.f90:126-128)
__global__ void k_guess_uniform(Dev d) {
    const MolDesc m = d.mol[blockIdx.y];
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m.nat) return;
    const AtomPar& a = d.atoms[m.at0 + i];
    double qi = m.charge / (double)m.nat;
    d.q[m.at0 + i] = qi;
    for (int s = 0; s < a.nsh; ++s) d.qsh[m.sh0 + a.sh0 + s] = d.shells[m.sh0 + a.sh0 + s].refocc / a.zval * qi;
    for (int k = 0; k < 3; ++k) d.dipm[3 * (m.at0 + i) + k] = 0.0;
    for (int k = 0; k < 6; ++k) d.qp[6 * (m.at0 + i) + k] = 0.0;
}
// q(atom) = sum of shell charges (qsh2qat, scc_core.f90:1429-1444)
__global__ void k_qsh2qat(Dev d) {
    const MolDesc m = d.mol[blockIdx.y];
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m.nat) return;
    const AtomPar& a = d.atoms[m.at0 + i];
    double s = 0;
    for (int k = 0; k < a.nsh; ++k) s += d.qsh[m.sh0 + a.sh0 + k];
    d.q[m.at0 + i] = s;
}

// ---- potentials -------------------------------------------------------------------------------
// Row of y = M x for a dense symmetric column-major n x n matrix.  CTA = 32 rows x kMvSlices column slices (256
// threads, launched with grid2(8 n, 256, nmol)): lanes read 32 consecutive rows (coalesced), the slices share the columns
// so that 8 times more loads are in flight than with one thread per row; partial sums are combined in a fixed order.
// Returns the row index (or -1) and the sum in the threads of slice 0.
constexpr int kMvSlices = 8;
__device__ __forceinline__ int matvec_row(const double* __restrict__ M, int n, const double* __restrict__ x, double& out) {
    __shared__ double sp[kMvSlices][33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int row = blockIdx.x * 32 + lane;
    double a0 = 0.0, a1 = 0.0;
    if (row < n) {
        int t = w;
        for (; t + kMvSlices < n; t += 2 * kMvSlices) {
            a0 += M[row + (long long)n * t] * x[t];
            a1 += M[row + (long long)n * (t + kMvSlices)] * x[t + kMvSlices];
        }
        if (t < n) a0 += M[row + (long long)n * t] * x[t];
    }
    sp[w][lane] = a0 + a1;
    __syncthreads();
    if (w != 0 || row >= n) return -1;
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < kMvSlices; ++k) s += sp[k][lane];
    out = s;
    return row;
}

// shell potential: Gamma_sh q_sh^2 + (J q)_sh
__global__ void __launch_bounds__(32 * kMvSlices) k_pot_shell(Dev d) {
    const MolDesc m = d.mol[blockIdx.y];
    if (d.st[blockIdx.y].done) return;
    if ((int)blockIdx.x * 32 >= m.nsh) return;
    const double* q = d.qsh + m.sh0;
    double acc = 0;
    const int s = matvec_row(d.jmat + m.ss0, m.nsh, q, acc);
    if (s < 0) return;
    // shellShift(:) = externShift (scc_core.f90:438): the potential of the external point charges, if any
    const double ext = d.pc0 ? d.vpc[m.sh0 + s] : 0.0;
    d.sshift[m.sh0 + s] = ext + q[s] * q[s] * d.shells[m.sh0 + s].gam3 + acc;
}

// potential of the external point charges on the shells (jpot_pcem_gfn2, src/embedding.f90:225-256); thread per shell
__global__ void k_pcem_pot(Dev d) {
    const MolDesc m = d.mol[blockIdx.y];
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= m.nsh) return;
    const ShellPar& sp = d.shells[m.sh0 + s];
    const double* ri = d.xyz + 3 * (m.at0 + sp.atom);
    double eh1 = 0.0;
    for (long long kk = d.pc0[blockIdx.y]; kk < d.pc0[blockIdx.y + 1]; ++kk) {
        const double dx = d.pc_xyz[3 * kk] - ri[0], dy = d.pc_xyz[3 * kk + 1] - ri[1], dz = d.pc_xyz[3 * kk + 2] - ri[2];
        const double xj = 0.5 * (sp.hard + d.pc_gam[kk]);
        eh1 += d.pc_q[kk] * (1.0 / sqrt(dx * dx + dy * dy + dz * dz + 1.0 / (xj * xj)));
    }
    d.vpc[m.sh0 + s] = eh1;
}

// anisotropic potentials vs, vd, vq; SHIFT = true: setvsdq (SCC), false: setdvsdq (gradient).  CTA = 32 atoms x
// kMvSlices slices of the partner loop (launched with grid2(8 nat, 256, nmol)); partial sums combined in a fixed order
template <bool SHIFT>
__global__ void __launch_bounds__(32 * kMvSlices) k_pot_aes(Dev d) {
    const MolDesc m = d.mol[blockIdx.y];
    if (SHIFT && d.st[blockIdx.y].done) return;
    if ((int)blockIdx.x * 32 >= m.nat) return;
    __shared__ double sp[kMvSlices][10][33];
    const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
    const int i = min(blockIdx.x * 32 + lane, m.nat - 1);
    const bool valid = (int)(blockIdx.x * 32 + lane) < m.nat;
    const double* xyz = d.xyz + 3 * m.at0;
    const double* q = d.q + m.at0;
    const double* dipm = d.dipm + 3 * m.at0;
    const double* qp = d.qp + 6 * m.at0;
    const double ra[3] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
    double stmp = 0, dtmp[3] = {0, 0, 0}, qtmp[6] = {0, 0, 0, 0, 0, 0};
    for (int j = slice; j < m.nat; j += kMvSlices) {
        const double g3 = d.gab3[m.nn0 + i + (long long)m.nat * j], g5 = d.gab5[m.nn0 + i + (long long)m.nat * j];
        const double dra[3] = {ra[0] - xyz[3 * j], ra[1] - xyz[3 * j + 1], ra[2] - xyz[3 * j + 2]};
        const double qj = q[j];
        double dum5a = 0, r2a = 0, r2ab = 0, t1a = 0, t2a = 0, t3a = 0;
#pragma unroll
        for (int l1 = 0; l1 < 3; ++l1) {
            if (SHIFT) { r2a += ra[l1] * ra[l1]; t1a += ra[l1] * dra[l1]; t3a += ra[l1] * dipm[3 * j + l1]; }
            r2ab += dra[l1] * dra[l1];
            t2a += dipm[3 * j + l1] * dra[l1];
#pragma unroll
            for (int l2 = 0; l2 < 3; ++l2) {
                const int ll = lin0(l1, l2);
                if (SHIFT) dum5a = dum5a - qp[6 * j + ll] * dra[l1] * dra[l2] - 1.50 * qj * dra[l1] * dra[l2] * ra[l1] * ra[l2];
                else dum5a = dum5a - qp[6 * j + ll] * dra[l1] * dra[l2];
                if (l2 >= l1) continue;
                qtmp[l1 + l2 + 2] -= 3.0 * qj * g5 * dra[l2] * dra[l1];
            }
            qtmp[l1] -= 1.50 * qj * g5 * dra[l1] * dra[l1];
        }
        if (SHIFT) {
            double dum3a = -t1a * qj - t2a;
            dum5a = dum5a + t3a * r2ab - 3.0 * t1a * t2a + 0.50 * qj * r2a * r2ab;
            stmp = stmp + dum5a * g5 + dum3a * g3;
        } else {
            stmp = stmp + (-t2a) * g3 + dum5a * g5;
        }
#pragma unroll
        for (int l1 = 0; l1 < 3; ++l1) {
            double d3 = dra[l1] * qj, d5;
            if (SHIFT) d5 = 3.0 * dra[l1] * t2a - r2ab * dipm[3 * j + l1] - qj * r2ab * ra[l1] + 3.0 * qj * dra[l1] * t1a;
            else d5 = 3.0 * dra[l1] * t2a - r2ab * dipm[3 * j + l1];
            dtmp[l1] += d3 * g3 + d5 * g5;
            qtmp[l1] += 0.50 * r2ab * qj * g5;
        }
    }
    sp[slice][0][lane] = stmp;
#pragma unroll
    for (int k = 0; k < 3; ++k) sp[slice][1 + k][lane] = dtmp[k];
#pragma unroll
    for (int k = 0; k < 6; ++k) sp[slice][4 + k][lane] = qtmp[k];
    __syncthreads();
    if (slice != 0 || !valid) return;
    double vs = 0.0, vd[3] = {0.0, 0.0, 0.0}, vq[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    for (int w = 0; w < kMvSlices; ++w) {
        vs += sp[w][0][lane];
#pragma unroll
        for (int k = 0; k < 3; ++k) vd[k] += sp[w][1 + k][lane];
#pragma unroll
        for (int k = 0; k < 6; ++k) vq[k] += sp[w][4 + k][lane];
    }
    const AtomPar& a = d.atoms[m.at0 + i];
    const double qs1 = a.dipk * 2.0, qs2 = a.quadk * 6.0;
    double t3a = 0, t2a = 0;
#pragma unroll
    for (int l1 = 0; l1 < 3; ++l1) {
        if (SHIFT) t3a += ra[l1] * dipm[3 * i + l1] * qs1;
        vd[l1] -= qs1 * dipm[3 * i + l1];
#pragma unroll
        for (int l2 = 0; l2 < 3; ++l2) {
            if (l2 >= l1) continue;
            const int ll = lin0(l1, l2), ki = l1 + l2 + 2;
            vq[ki] -= qp[6 * i + ll] * qs2;
            if (SHIFT) {
                t3a -= ra[l1] * ra[l2] * qp[6 * i + ll] * qs2;
                vd[l1] += ra[l2] * qp[6 * i + ll] * qs2;
                vd[l2] += ra[l1] * qp[6 * i + ll] * qs2;
            }
        }
        const int ll = lin0(l1, l1);
        vq[l1] -= qp[6 * i + ll] * qs2 * 0.50;
        if (SHIFT) {
            t3a -= ra[l1] * ra[l1] * qp[6 * i + ll] * qs2 * 0.50;
            vd[l1] += ra[l1] * qp[6 * i + ll] * qs2;
        }
        t2a += qp[6 * i + ll];
    }
    if (SHIFT) vs += t3a;
    t2a *= a.quadk;
#pragma unroll
    for (int l1 = 0; l1 < 3; ++l1) {
        vq[l1] += t2a;
        if (SHIFT) { vd[l1] -= 2.0 * ra[l1] * t2a; vs += t2a * ra[l1] * ra[l1]; }
    }
    d.vs[m.at0 + i] = vs;
    for (int k = 0; k < 3; ++k) d.vd[3 * (m.at0 + i) + k] = vd[k];
    for (int k = 0; k < 6; ++k) d.vq[6 * (m.at0 + i) + k] = vq[k];
}

// D4 zeta / dzeta vectors for the current atomic charges (dftd4.F90:775-788, 848-857); thread per reference
__global__ void k_d4_zeta(Dev d, int use_out) {
    const MolDesc m = d.mol[blockIdx.y];
    if (d.st[blockIdx.y].done) return;
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m.ndim) return;
    const RefPar r = d.refs[m.ref0 + k];
    const AtomPar& a = d.atoms[m.at0 + r.atom];
    double z = 0, dz = 0;
    if (!(d.gw[m.ref0 + k] < 1.0e-7)) {
        double q = use_out ? d.q_out[m.at0 + r.atom] : d.q[m.at0 + r.atom];
        z = d4_zeta(3.0, a.d4_gam * 2.0, r.refq + a.d4_zeff, q + a.d4_zeff);
        dz = d4_dzeta(3.0, a.d4_gam * 2.0, r.refq + a.d4_zeff, q + a.d4_zeff);
    }
    d.zeta[m.ref0 + k] = z;
    d.dzeta[m.ref0 + k] = dz;
}
// dvec = dispmat * zeta
__global__ void __launch_bounds__(32 * kMvSlices) k_d4_matvec(Dev d) {
    const MolDesc m = d.mol[blockIdx.y];
    if (d.st[blockIdx.y].done) return;
    if ((int)blockIdx.x * 32 >= m.ndim) return;
    double acc = 0;
    const int k = matvec_row(d.dispmat + m.dd0, m.ndim, d.zeta + m.ref0, acc);
    if (k >= 0) d.dvec[m.ref0 + k] = acc;
}
// atomic shift from dispersion: sum_ref dvec*dzeta (dftd4.F90:805-811); thread per atom
__global__ void k_d4_shift(Dev d) {
    const MolDesc m = d.mol[blockIdx.y];
    if (d.st[blockIdx.y].done) return;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m.nat) return;
    const AtomPar& a = d.atoms[m.at0 + i];
    double s = 0;
    if (d.solv_on) s = d.sv_shift[m.at0 + i];   // solvation%addShift precedes scD4%addShift, scc_core.f90:444-452
    for (int r = 0; r < a.nref; ++r) s += d.dvec[m.ref0 + a.ref0 + r] * d.dzeta[m.ref0 + a.ref0 + r];
    d.ashift[m.at0 + i] = s;
}

// ---- Fock assembly, scc_core.f90:154-251 ; thread per matrix element (both triangles written) ----
__global__ void k_fock(Dev d) {
    const MolDesc m = d.mol[blockIdx.y];
    if (d.st[blockIdx.y].done) return;
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)m.nao * m.nao) return;
    int i = (int)(t % m.nao), j = (int)(t / m.nao);
    const long long a = m.mat0 + t;
    const int ish = d.ao2sh[m.ao0 + i], jsh = d.ao2sh[m.ao0 + j];
    const int iat = d.ao2at[m.ao0 + i], jat = d.ao2at[m.ao0 + j];
    const double dum = d.S[a] * kAutoEV * 0.5;
    const double shi = d.sshift[m.sh0 + ish] + d.ashift[m.at0 + iat];
    const double shj = d.sshift[m.sh0 + jsh] + d.ashift[m.at0 + jat];
    double eh1 = -dum * (shi + shj);
    double h = d.H0[a] + eh1;
    eh1 = dum * (d.vs[m.at0 + iat] + d.vs[m.at0 + jat]);
    h = h + eh1;
    const uint8_t mk = d.mask[a];
    if (mk & 1) {
        double e = 0;
#pragma unroll
        for (int l = 0; l < 3; ++l) e += d.D[l * d.tot_mat + a] * (d.vd[3 * (m.at0 + iat) + l] + d.vd[3 * (m.at0 + jat) + l]);
        h += 0.50 * e * kAutoEV;
    }
    if (mk & 2) {
        double e = 0;
#pragma unroll
        for (int l = 0; l < 6; ++l) e += d.Q[l * d.tot_mat + a] * (d.vq[6 * (m.at0 + iat) + l] + d.vq[6 * (m.at0 + jat) + l]);
        h += 0.50 * e * kAutoEV;
    }
    d.H[a] = h;
}

// ---- Fermi smearing + occupations, scc_core.f90:983-1064 ; one warp per molecule ----------------
__device__ void fermi_one(int norbs, int nel, double t, const double* eig, double* occ, double& ef, double& s) {
    const int lane = threadIdx.x & 31;
    const double boltz = kBoltz * kAutoEV, thr = 1e-9;
    const double bkt = boltz * t;
    for (int i = lane; i < norbs; i += 32) occ[i] = (i < nel) ? 1.0 : 0.0;
    __syncwarp();
    if (nel + 1 > norbs) ef = eig[nel - 1];
    else if (nel == 0) ef = eig[0];
    else {
        ef = 0.5 * (eig[nel - 1] + eig[nel]);
        const double occt = nel;
        for (int ncycle = 1; ncycle <= 200; ++ncycle) {
            double tn = 0, td = 0;
            for (int i = lane; i < norbs; i += 32) {
                double f = 0, df = 0;
                double x = (eig[i] - ef) / bkt;
                if (x < 50) {
                    double ex = exp(x);
                    f = 1.0 / (ex + 1.0);
                    df = ex / (bkt * (ex + 1.0) * (ex + 1.0));
                }
                occ[i] = f;
                tn += f; td += df;
            }
            tn = warp_sum(tn); td = warp_sum(td);
            ef += (occt - tn) / td;   // unconditional update, scc_core.f90:1041
            if (fabs(occt - tn) <= thr) break;
        }
    }
    __syncwarp();
    double ss = 0;
    for (int i = lane; i < norbs; i += 32) {
        double o = occ[i];
        if (o > thr && 1.0 - o > thr) ss += o * log(o) + (1.0 - o) * log(1.0 - o);
    }
    ss = warp_sum(ss);
    s = ss * kBoltz * t;
}
__global__ void k_fermi(Dev d) {
    const int im = blockIdx.x;
    const MolDesc m = d.mol[im];
    MolState& st = d.st[im];
    if (st.done) return;
    const double* emo = d.emo + m.ao0;
    double *fa = d.focca + m.ao0, *fb = d.foccb + m.ao0, *f = d.focc + m.ao0;
    const int lane = threadIdx.x;
    double efa = 0, efb = 0, ga = 0, gb = 0;
    if (d.etemp > 0.1) {
        fermi_one(m.nao, m.ihomoa, d.etemp, emo, fa, efa, ga);
        fermi_one(m.nao, m.ihomob, d.etemp, emo, fb, efb, gb);
    } else {
        for (int i = lane; i < m.nao; i += 32) { fa[i] = i < m.ihomoa ? 1.0 : 0.0; fb[i] = i < m.ihomob ? 1.0 : 0.0; }
    }
    __syncwarp();
    int kocc = 0;
    for (int i = lane; i < m.nao; i += 32) {
        double v = fa[i] + fb[i];
        f[i] = v;
        if (v != 0.0) kocc = max(kocc, i + 1);
    }
    for (int o = 16; o > 0; o >>= 1) kocc = max(kocc, __shfl_xor_sync(0xffffffffu, kocc, o));
    if (lane == 0) {
        st.kocc = kocc; st.ga = ga; st.gb = gb; st.efa = efa; st.efb = efb;
        int ih = m.ihomoa;
        if (ih + 1 <= m.nao && ih >= 1) st.egap = emo[ih] - emo[ih - 1];
    }
}

// T[:, k] = C[:, k] * w[k] for k < kocc, zero otherwise. mode 0: w = focc, 1: focc*emo*evtoau, 2: focca, 3: foccb
__global__ void k_scale_cols(Dev d, int mode) {
    const MolDesc m = d.mol[blockIdx.y];
    if (mode == 0 && d.st[blockIdx.y].done) return;
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)m.nao * m.nao) return;
    int k = (int)(t / m.nao);
    double w;
    if (mode == 0) w = d.focc[m.ao0 + k];
    else if (mode == 1) w = d.focc[m.ao0 + k] * d.emo[m.ao0 + k] * kEVtoAU;
    else if (mode == 2) w = d.focca[m.ao0 + k];
    else w = d.foccb[m.ao0 + k];
    d.T[m.mat0 + t] = d.C[m.mat0 + t] * w;
}

// ---- populations: per-AO row partial sums of P.S, P.D, P.Q, P.H0 -------------------------------
// aop[ao][0] = sum_j P_ij S_ij ; [1..3] = sum_j P_ij D_k,ij ; [4..9] = sum_j P_ij Q_k,ij ; [10] = sum_j P_ij H0_ij
// (mpopsh scc_core.f90:1408-1426, mmompop aespot.F90:289-427, electro scc_core.f90:686-722 in row form)
// one warp per AO row: lanes stride over j (coalesced since all matrices are symmetric)
__global__ void k_pop_rows(Dev d) {
    const MolDesc m = d.mol[blockIdx.y];
    if (d.st[blockIdx.y].done) return;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= m.nao) return;
    const long long row = m.mat0 + (long long)m.nao * warp;
    double acc[11];
#pragma unroll
    for (int k = 0; k < 11; ++k) acc[k] = 0.0;
    for (int j = lane; j < m.nao; j += 32) {
        const double p = d.P[row + j];
        acc[0] += p * d.S[row + j];
#pragma unroll
        for (int k = 0; k < 3; ++k) acc[1 + k] += p * d.D[k * d.tot_mat + row + j];
#pragma unroll
        for (int k = 0; k < 6; ++k) acc[4 + k] += p * d.Q[k * d.tot_mat + row + j];
        acc[10] += p * d.H0[row + j];
    }
#pragma unroll
    for (int k = 0; k < 11; ++k) acc[k] = warp_sum(acc[k]);
    if (lane == 0) {
        double* o = d.aop + 12 * (long long)(m.ao0 + warp);
#pragma unroll
        for (int k = 0; k < 11; ++k) o[k] = acc[k];
    }
}

// per-atom output charges and multipoles from the AO partials; thread per atom
__global__ void k_pop_atoms(Dev d) {
    const MolDesc m = d.mol[blockIdx.y];
    if (d.st[blockIdx.y].done) return;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m.nat) return;
    const AtomPar& a = d.atoms[m.at0 + i];
    const double* ra = d.xyz + 3 * (m.at0 + i);
    double qat = 0, dp[3] = {0, 0, 0}, qq[6] = {0, 0, 0, 0, 0, 0};
    for (int s = 0; s < a.nsh; ++s) {
        const ShellPar& sp = d.shells[m.sh0 + a.sh0 + s];
        const int na = c_nsph[sp.l];
        double pop = 0;
        for (int k = 0; k < na; ++k) {
            const double* o = d.aop + 12 * (long long)(m.ao0 + sp.ao0 + k);
            const double ps = o[0];
            pop += ps;
            // cumulative atomic multipoles, origin shifted to the atom (aespot.F90:329-362)
            for (int c = 0; c < 3; ++c) {
                const double xk = ra[c], pdmk = o[1 + c];
                dp[c] += xk * ps - pdmk;
                for (int l = 0; l < c; ++l) {
                    const int kl = c * (c + 1) / 2 + l, kj = c + l + 2;
                    const double xl = ra[l], pdml = o[1 + l], pqm = o[4 + kj];
                    qq[kl] += pdmk * xl + pdml * xk - xl * xk * ps - pqm;
                }
                qq[c * (c + 1) / 2 + c] += 2.0 * pdmk * xk - xk * xk * ps - o[4 + c];
            }
        }
        const double qs = sp.refocc - pop;
        d.qsh_out[m.sh0 + a.sh0 + s] = qs;
        qat += qs;
    }
    d.q_out[m.at0 + i] = qat;
    const double tii = 0.50 * (qq[0] + qq[2] + qq[5]);
#pragma unroll
    for (int k = 0; k < 6; ++k) qq[k] *= 1.50;
    qq[0] -= tii; qq[2] -= tii; qq[5] -= tii;
    for (int k = 0; k < 3; ++k) d.dipm_out[3 * (m.at0 + i) + k] = dp[k];
    for (int k = 0; k < 6; ++k) d.qp_out[6 * (m.at0 + i) + k] = qq[k];
}

// (J q_out) per shell for the ES energy; thread per shell
__global__ void __launch_bounds__(32 * kMvSlices) k_es_rows(Dev d) {
    const MolDesc m = d.mol[blockIdx.y];
    if (d.st[blockIdx.y].done) return;
    if ((int)blockIdx.x * 32 >= m.nsh) return;
    double acc = 0;
    const int s = matvec_row(d.jmat + m.ss0, m.nsh, d.qsh_out + m.sh0, acc);
    if (s >= 0) d.jq[m.sh0 + s] = acc;
}

// anisotropic ES energy rows (aniso_electro, aespot.F90:569-662): per atom i, sum over j < i
// e_atom[i][1] = sum_j ed*gab3 + (eq+edd)*gab5 ; e_atom[i][3] = on-site kernel energy
__global__ void k_aes_rows(Dev d) {
    const MolDesc m = d.mol[blockIdx.y];
    if (d.st[blockIdx.y].done) return;
    const int t = blockIdx.x * blockDim.x + threadIdx.x, sl = t % kAS;
    const bool active = t / kAS < m.nat;
    const int i = active ? t / kAS : 0;
    const int idx[3][3] = {{0, 1, 3}, {1, 2, 4}, {3, 4, 5}};
    const double* xyz = d.xyz + 3 * m.at0;
    const double* q = d.q_out + m.at0;
    const double* dipm = d.dipm_out + 3 * m.at0;
    const double* qp = d.qp_out + 6 * m.at0;
    const double q1 = q[i];
    const double dp1[3] = {dipm[3 * i], dipm[3 * i + 1], dipm[3 * i + 2]};
    double qp1[6];
    for (int k = 0; k < 6; ++k) qp1[k] = qp[6 * i + k];
    double e01 = 0, e02 = 0, e11 = 0;
    for (int j = sl; j < (active ? i : 0); j += kAS) {
        const double rij[3] = {xyz[3 * j] - xyz[3 * i], xyz[3 * j + 1] - xyz[3 * i + 1], xyz[3 * j + 2] - xyz[3 * i + 2]};
        const double r2 = rij[0] * rij[0] + rij[1] * rij[1] + rij[2] * rij[2];
        double ed = 0, eq = 0, edd = 0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            ed += q[j] * dp1[k] * rij[k];
            ed -= dipm[3 * j + k] * q1 * rij[k];
#pragma unroll
            for (int l = 0; l < 3; ++l) {
                const int kl = idx[l][k];
                const double tt = rij[l] * rij[k], tt3 = 3.0 * tt;
                eq += q[j] * qp1[kl] * tt;
                eq += qp[6 * j + kl] * q1 * tt;
                edd -= dipm[3 * j + k] * dp1[l] * tt3;
            }
            edd += dipm[3 * j + k] * dp1[k] * r2;
        }
        const double g3 = d.gab3[m.nn0 + j + (long long)m.nat * i], g5 = d.gab5[m.nn0 + j + (long long)m.nat * i];
        e01 += ed * g3; e02 += eq * g5; e11 += edd * g5;
    }
    e01 = slice_sum(e01); e02 = slice_sum(e02); e11 = slice_sum(e11);
    if (!active || sl) return;
    const AtomPar& a = d.atoms[m.at0 + i];
    double tt = dp1[0] * dp1[0] + dp1[1] * dp1[1] + dp1[2] * dp1[2], tt3 = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int l = 0; l < 3; ++l) tt3 += qp1[idx[l][k]] * qp1[idx[l][k]];
    d.e_atom[4 * (m.at0 + i) + 1] = e01 + e02 + e11;
    d.e_atom[4 * (m.at0 + i) + 3] = a.dipk * tt + tt3 * a.quadk;
}

// ---- end of iteration: energies, convergence, Broyden mixing; one CTA per molecule --------------
// solves beta*gamma = c by Gaussian elimination with partial pivoting (replaces matinv = dgetrf+dgetri
// followed by gamma = c*beta^-1, src/broyden.f90:117-121,156-184; beta is symmetric)
__device__ void small_solve(double* B, double* c, int n, int ld) {
    __shared__ int s_p;
    for (int k = 0; k < n; ++k) {
        if (threadIdx.x == 0) {
            int p = k; double best = fabs(B[k + ld * k]);
            for (int i = k + 1; i < n; ++i) { double v = fabs(B[i + ld * k]); if (v > best) { best = v; p = i; } }
            s_p = p;
        }
        __syncthreads();
        const int p = s_p;
        if (p != k) {
            for (int j = threadIdx.x; j < n; j += blockDim.x) { double t = B[k + ld * j]; B[k + ld * j] = B[p + ld * j]; B[p + ld * j] = t; }
            if (threadIdx.x == 0) { double t = c[k]; c[k] = c[p]; c[p] = t; }
        }
        __syncthreads();
        const double pinv = 1.0 / B[k + ld * k];
        const int nr = n - k - 1;
        for (int t = threadIdx.x; t < nr * nr; t += blockDim.x) {
            int i = k + 1 + t % nr, j = k + 1 + t / nr;
            B[i + ld * j] -= B[i + ld * k] * pinv * B[k + ld * j];
        }
        __syncthreads();
        for (int i = k + 1 + threadIdx.x; i < n; i += blockDim.x) c[i] -= B[i + ld * k] * pinv * c[k];
        __syncthreads();
    }
    for (int k = n - 1; k >= 0; --k) {
        if (threadIdx.x == 0) c[k] /= B[k + ld * k];
        __syncthreads();
        const double ck = c[k];
        for (int i = threadIdx.x; i < k; i += blockDim.x) c[i] -= B[i + ld * k] * ck;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) k_scc_finish(Dev d) {
    const int im = blockIdx.x;
    const MolDesc m = d.mol[im];
    MolState& st = d.st[im];
    if (st.done) return;
    __shared__ double red[32];
    __shared__ double s_c[128], s_om[128];
    const int tid = threadIdx.x, nt = blockDim.x;
    const int iter = st.iter + 1;   // 1-based iteration number
    // --- energies
    double p = 0;
    for (int s = tid; s < m.nsh; s += nt) {
        const double qs = d.qsh_out[m.sh0 + s];
        p += 0.5 * d.jq[m.sh0 + s] * qs + qs * qs * qs * d.shells[m.sh0 + s].gam3 / 3.0;
    }
    const double ees = block_sum(p, red);
    p = 0;
    for (int i = tid; i < m.nao; i += nt) p += d.aop[12 * (long long)(m.ao0 + i) + 10];
    const double eh0 = block_sum(p, red);   // sum_ij P_ij H0_ij = 2*h of electro()
    double p1 = 0, p2 = 0;
    for (int i = tid; i < m.nat; i += nt) { p1 += d.e_atom[4 * (m.at0 + i) + 1]; p2 += d.e_atom[4 * (m.at0 + i) + 3]; }
    const double eaes = block_sum(p1, red);
    const double epol = block_sum(p2, red);
    p = 0;
    for (int k = tid; k < m.ndim; k += nt) p += 0.5 * d.dvec[m.ref0 + k] * d.zeta[m.ref0 + k];
    const double ed = block_sum(p, red);
    double epcem = 0.0;
    if (d.pc0) {   // electro_pcem, src/embedding.f90:261-280 (scc_core.f90:539-541)
        p = 0;
        for (int s = tid; s < m.nsh; s += nt) p += d.qsh_out[m.sh0 + s] * d.vpc[m.sh0 + s];
        epcem = block_sum(p, red);
    }
    double gsolv = 0.0;
    if (d.solv_on) {   // solvation%getEnergy with the output charges, scc_core.f90:543-548 -> gbsa.f90:510-535
        p = 0;
        for (int i = tid; i < m.nat; i += nt) p += d.q_out[m.at0 + i] * (0.5 * d.sv_bq[m.at0 + i]);
        gsolv = block_sum(p, red) + d.sv_mol[24 * (size_t)im] + d.solv_gshift;
    }
    const double eold = st.eel;
    double eel = ees + eh0 * kEVtoAU;
    eel = eel + eaes + epol;
    eel = eel + ed;
    eel = eel + epcem;
    eel = eel + gsolv;
    eel = eel + st.ga + st.gb;
    const bool econv = fabs(eel - eold) < d.scfconv;
    // --- charge difference vector (gfn2broyden_save/diff, aespot.F90:1360-1425)
    double* qin = d.bqin + m.q0;
    double* dq = d.bdq + m.q0;
    p = 0;
    for (int k = tid; k < m.nbr; k += nt) {
        double vin, vout;
        if (k < m.nsh) { vin = d.qsh[m.sh0 + k]; vout = d.qsh_out[m.sh0 + k]; }
        else {
            int r = k - m.nsh, i = r / 9, c = r % 9;
            if (c < 3) { vin = d.dipm[3 * (m.at0 + i) + c]; vout = d.dipm_out[3 * (m.at0 + i) + c]; }
            else { vin = d.qp[6 * (m.at0 + i) + c - 3]; vout = d.qp_out[6 * (m.at0 + i) + c - 3]; }
        }
        qin[k] = vin;
        const double dv = vout - vin;
        dq[k] = dv;
        p += dv * dv;
    }
    const double dq2 = block_sum(p, red);
    const double rmsq = sqrt(dq2 / (double)m.nat);
    const bool qconv = rmsq < d.qconv;
    // --- Broyden (src/broyden.f90:37-150); n == 1 keeps the plain output (broy = mol%n > 1)
    double* qlast = d.bqlast + m.q0;
    double* dqlast = d.bdqlast + m.q0;
    const double alpha = d.broydamp;
    double omegap = 0.0;      // weight of this iteration's Broyden vector (printed by the reference as omega)
    if (m.nat > 1) {
        if (iter == 1) {
            for (int k = tid; k < m.nbr; k += nt) { dqlast[k] = dq[k]; qlast[k] = qin[k]; qin[k] = qin[k] + alpha * dq[k]; }
        } else if (iter - 1 > d.maxhist) {
            if (tid == 0) st.fail = 20;   // Broyden history exhausted
        } else {
            const int it1 = iter - 1, c1 = it1 - 1;
            double* df = d.bdf + m.br0;
            double* u = d.bu + m.br0;
            double* a = d.ba + (long long)im * d.maxhist * d.maxhist;
            double* beta = d.bbeta + (long long)im * d.maxhist * d.maxhist;
            double* omega = d.bomega + (long long)im * d.maxhist;
            const int mh = d.maxhist;
            double om = sqrt(dq2);
            om = (om > (0.01 / 100000.0)) ? 0.01 / om : 100000.0;
            if (om < 1.0) om = 1.0;
            omegap = om;
            double* dfc = df + (long long)c1 * m.nbr;
            p = 0;
            for (int k = tid; k < m.nbr; k += nt) { double v = dq[k] - dqlast[k]; dfc[k] = v; p += v * v; }
            const double inv = 1.0 / sqrt(block_sum(p, red));
            for (int k = tid; k < m.nbr; k += nt) dfc[k] *= inv;
            __syncthreads();
            if (tid == 0) omega[c1] = om;
            // a(i,it1) = <df_i|df_it1>, c(i) = omega_i <df_i|dq> : one warp per history vector
            {
                const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
                for (int i = warp; i < it1; i += nw) {
                    const double* dfi = df + (long long)i * m.nbr;
                    double s = 0, t = 0;
                    for (int k = lane; k < m.nbr; k += 32) { s += dfi[k] * dfc[k]; t += dfi[k] * dq[k]; }
                    s = warp_sum(s); t = warp_sum(t);
                    if (lane == 0) {
                        a[i + mh * c1] = s; a[c1 + mh * i] = s;
                        s_c[i] = (i == c1 ? om : omega[i]) * t;
                        s_om[i] = (i == c1 ? om : omega[i]);
                    }
                }
            }
            __syncthreads();
            for (int t = tid; t < it1 * it1; t += nt) {
                int i = t % it1, j = t / it1;
                double v = s_om[i] * s_om[j] * a[i + mh * j];
                if (i == j) v += 0.01 * 0.01;
                beta[i + mh * j] = v;
            }
            __syncthreads();
            small_solve(beta, s_c, it1, mh);   // s_c <- gamma
            double* uc = u + (long long)c1 * m.nbr;
            for (int k = tid; k < m.nbr; k += nt) {
                uc[k] = alpha * dfc[k] + inv * (qin[k] - qlast[k]);
            }
            __syncthreads();
            for (int k = tid; k < m.nbr; k += nt) {
                dqlast[k] = dq[k];
                qlast[k] = qin[k];
                double v = qin[k] + alpha * dq[k];
                for (int i = 0; i < it1; ++i) v -= s_om[i] * s_c[i] * u[(long long)i * m.nbr + k];
                qin[k] = v;
            }
        }
        __syncthreads();
        // unpack mixed vector (gfn2broyden_out)
        for (int k = tid; k < m.nbr; k += nt) {
            const double v = qin[k];
            if (k < m.nsh) d.qsh[m.sh0 + k] = v;
            else {
                int r = k - m.nsh, i = r / 9, c = r % 9;
                if (c < 3) d.dipm[3 * (m.at0 + i) + c] = v;
                else d.qp[6 * (m.at0 + i) + c - 3] = v;
            }
        }
    } else {
        for (int k = tid; k < m.nbr; k += nt) {
            const double v = qin[k] + dq[k];
            if (k < m.nsh) d.qsh[m.sh0 + k] = v;
            else {
                int r = k - m.nsh, c = r % 9;
                if (c < 3) d.dipm[3 * m.at0 + c] = v; else d.qp[6 * m.at0 + c - 3] = v;
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < m.nat; i += nt) {   // qsh2qat on the mixed charges
        const AtomPar& a = d.atoms[m.at0 + i];
        double s = 0;
        for (int k = 0; k < a.nsh; ++k) s += d.qsh[m.sh0 + a.sh0 + k];
        d.q[m.at0 + i] = s;
    }
    if (tid == 0) {
        st.eold = eold; st.eel = eel; st.ees = ees; st.eaes = eaes; st.epol = epol; st.ed = ed; st.eh0 = eh0; st.epcem = epcem;
        st.rmsq = rmsq;
        st.iter = iter;
        if (d.iterlog && im == 0 && iter <= d.maxiter) {
            double* lg = d.iterlog + 5 * (size_t)(iter - 1);
            lg[0] = eel; lg[1] = eel - eold; lg[2] = rmsq; lg[3] = st.egap; lg[4] = omegap;
        }
        if (econv && qconv) {
            st.converged = 1;
            if (st.lastdiag) st.done = 1;
            st.lastdiag = 1;
        }
        if (iter >= d.maxiter) st.done = 1;
        if (st.fail) st.done = 1;
    }
}

}  // namespace xtbb200
