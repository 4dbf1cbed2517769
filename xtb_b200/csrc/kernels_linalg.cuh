// Dense FP64 linear algebra for the generalized eigenproblem H C = S C e, batched over molecules.
//
//   once per geometry : S = U^T U (Cholesky), X = U^-1            [k_chol_inv]
//   per SCC iteration : A' = X^T H X  (2 GEMMs on the FP64 tensor path, DMMA m8n8k4)
//                       A' Z = Z e    (one-sided Jacobi in shared memory)   [k_jacobi]
//                       C = X Z       (GEMM),  P = (C f) C^T (GEMM over occupied columns)
//
// Replaces: dpotrf (src/scc_core.f90:411-413), dsygst + dsyevd + dtrsm
// (src/mctc/lapack/eigensolve.F90:218-246), dgemm in dmat (src/scc_core.f90:1177-1205).
#pragma once
#include "common.cuh"

namespace xtbb200 {

// ---------------------------------------------------------------------------------------------
// batched GEMM  Cm = op(A) * op(B),  column-major, leading dimension nao; FP64 tensor cores
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma_884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

constexpr int GT = 64;     // CTA tile
constexpr int GK = 16;     // k tile
constexpr int GLD = GT + 4;

// kmode: 0 -> K = nao ; 1 -> K = st.kocc (density build over occupied columns)
// skip_done: 1 = skip molecules that already left the SCC loop; 2 = additionally skip molecules without a valid
// eigenvector guess (warm-start product A' Z_prev)
template <bool TA, bool TB>
__global__ void __launch_bounds__(128) k_gemm(Dev d, const double* __restrict__ Ag, const double* __restrict__ Bg,
                                              double* __restrict__ Cg, int kmode, int skip_done) {
    const MolDesc m = d.mol[blockIdx.y];
    if (skip_done && d.st[blockIdx.y].done) return;
    if (skip_done == 2 && !d.st[blockIdx.y].zvalid) return;
    const int n = m.nao;
    const int tiles = (n + GT - 1) / GT;
    if ((int)blockIdx.x >= tiles * tiles) return;
    const int r0 = (blockIdx.x % tiles) * GT, c0 = (blockIdx.x / tiles) * GT;
    const int K = kmode == 1 ? d.st[blockIdx.y].kocc : n;
    const double* A = Ag + m.mat0;
    const double* B = Bg + m.mat0;
    double* C = Cg + m.mat0;
    __shared__ double As[GK][GLD], Bs[GK][GLD];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wr = (warp & 1) * 32, wc = (warp >> 1) * 32;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int k0 = 0; k0 < K; k0 += GK) {
        // stage tiles: As[kk][r] = op(A)(r0+r, k0+kk), Bs[kk][c] = op(B)(k0+kk, c0+c)
        if (!TA) {
            for (int t = tid; t < GT * GK; t += 128) {
                int r = t % GT, kk = t / GT;
                int gr = r0 + r, gk = k0 + kk;
                As[kk][r] = (gr < n && gk < K) ? A[gr + (long long)n * gk] : 0.0;
            }
        } else {
            for (int t = tid; t < GT * GK; t += 128) {
                int kk = t % GK, r = t / GK;
                int gr = r0 + r, gk = k0 + kk;
                As[kk][r] = (gr < n && gk < K) ? A[gk + (long long)n * gr] : 0.0;
            }
        }
        if (!TB) {
            for (int t = tid; t < GT * GK; t += 128) {
                int kk = t % GK, c = t / GK;
                int gc = c0 + c, gk = k0 + kk;
                Bs[kk][c] = (gc < n && gk < K) ? B[gk + (long long)n * gc] : 0.0;
            }
        } else {
            for (int t = tid; t < GT * GK; t += 128) {
                int c = t % GT, kk = t / GT;
                int gc = c0 + c, gk = k0 + kk;
                Bs[kk][c] = (gc < n && gk < K) ? B[gc + (long long)n * gk] : 0.0;
            }
        }
        __syncthreads();
#pragma unroll
        for (int ks = 0; ks < GK; ks += 4) {
            double af[4], bf[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) af[i] = As[ks + (lane & 3)][wr + 8 * i + (lane >> 2)];
#pragma unroll
            for (int j = 0; j < 4; ++j) bf[j] = Bs[ks + (lane & 3)][wc + 8 * j + (lane >> 2)];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma_884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int gr = r0 + wr + 8 * i + (lane >> 2);
            int gc = c0 + wc + 8 * j + 2 * (lane & 3);
            if (gr < n) {
                if (gc < n) C[gr + (long long)n * gc] = acc[i][j][0];
                if (gc + 1 < n) C[gr + (long long)n * (gc + 1)] = acc[i][j][1];
            }
        }
}

// ---------------------------------------------------------------------------------------------
// Cholesky S = L L^T (U = L^T) and X = U^-1 = L^-T ; one CTA per molecule; matrix in shared memory
// when it fits (SM = true), otherwise in global scratch (W).
// ---------------------------------------------------------------------------------------------
template <bool SM>
__global__ void __launch_bounds__(512) k_chol_inv(Dev d) {
    const int im = blockIdx.x;
    const MolDesc m = d.mol[im];
    const int n = m.nao;
    extern __shared__ double smem[];
    const int ld = SM ? (n | 1) : n;
    double* L = SM ? smem : d.W + m.mat0;
    const double* S = d.S + m.mat0;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (long long t = tid; t < (long long)n * n; t += nt) {
        int i = (int)(t % n), j = (int)(t / n);
        L[i + (long long)ld * j] = S[t];
    }
    __syncthreads();
    for (int k = 0; k < n; ++k) {
        const double akk = L[k + (long long)ld * k];
        if (!(akk > 0.0)) { if (tid == 0) d.st[im].fail = 11; return; }  // uniform: all threads read the same value
        const double lkk = sqrt(akk);
        __syncthreads();
        for (int i = k + tid; i < n; i += nt) L[i + (long long)ld * k] = (i == k) ? lkk : L[i + (long long)ld * k] / lkk;
        __syncthreads();
        const int nr = n - k - 1;
        for (long long t = tid; t < (long long)nr * nr; t += nt) {
            int i = k + 1 + (int)(t % nr), j = k + 1 + (int)(t / nr);
            if (i >= j) L[i + (long long)ld * j] -= L[i + (long long)ld * k] * L[j + (long long)ld * k];
        }
        __syncthreads();
    }
    // Linv by forward substitution, one column per thread; X(i,j) = Linv(j,i)
    double* X = d.X + m.mat0;
    for (int j = tid; j < n; j += nt) {
        // column j of Linv is stored into row j of X (upper triangle): X[j + n*i], i >= j
        for (int i = 0; i < j; ++i) X[j + (long long)n * i] = 0.0;
        for (int i = j; i < n; ++i) {
            double s = (i == j) ? 1.0 : 0.0;
            for (int k = j; k < i; ++k) s -= L[i + (long long)ld * k] * X[j + (long long)n * k];
            X[j + (long long)n * i] = s / L[i + (long long)ld * i];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// symmetric eigensolver: one-sided (Hestenes) Jacobi on G = A' + sigma*I (SPD), one CTA per matrix.
// Columns of G are orthogonalised by plane rotations in a round-robin parallel ordering;
// at convergence column norms are the shifted eigenvalues and the normalised columns the
// eigenvectors.  TG lanes cooperate on one column pair.
// ---------------------------------------------------------------------------------------------
template <int TG, bool SM>
__global__ void __launch_bounds__(1024) k_jacobi(Dev d) {
    const int im = blockIdx.x;
    const MolDesc m = d.mol[im];
    MolState& st = d.st[im];
    if (st.done) return;
    const int n = m.nao;
    extern __shared__ double smem[];
    const int ld = SM ? (n | 1) : n;
    double* G = SM ? smem : d.W + m.mat0;
    __shared__ double red[32];
    __shared__ double s_sigma;
    const int tid = threadIdx.x, nt = blockDim.x;
    const double* A = d.W + m.mat0;
    // Gershgorin bound
    double mx = 0.0;
    for (int i = tid; i < n; i += nt) {
        double s = 0;
        for (int j = 0; j < n; ++j) s += fabs(A[i + (long long)n * j]);
        mx = fmax(mx, s);
    }
    mx = warp_max(mx);
    if ((tid & 31) == 0) red[tid >> 5] = mx;
    __syncthreads();
    if (tid == 0) {
        double v = 0;
        for (int w = 0; w < (nt + 31) / 32; ++w) v = fmax(v, red[w]);
        s_sigma = 1.05 * v + 1.0;
    }
    __syncthreads();
    const double sigma = s_sigma;
    if (st.zvalid) {
        // warm start: G = (A' + sigma I) Z_prev = T + sigma Z_prev, T = A' Z_prev from the preceding k_gemm.
        // Z_prev is orthonormal, so the columns of G are already nearly orthogonal and a few sweeps finish the job.
        const double* T = d.T + m.mat0;
        const double* Zp = d.Zm + m.mat0;
        for (long long t = tid; t < (long long)n * n; t += nt) {
            int i = (int)(t % n), j = (int)(t / n);
            G[i + (long long)ld * j] = T[t] + sigma * Zp[t];
        }
    } else if (SM) {
        for (long long t = tid; t < (long long)n * n; t += nt) {
            int i = (int)(t % n), j = (int)(t / n);
            // symmetrise on load: the two triangles of X^T H X agree only to rounding
            double v = 0.5 * (A[i + (long long)n * j] + A[j + (long long)n * i]);
            G[i + (long long)ld * j] = v + (i == j ? sigma : 0.0);
        }
    } else {
        __syncthreads();
        for (long long t = tid; t < (long long)n * n; t += nt) {
            int i = (int)(t % n), j = (int)(t / n);
            if (i > j) {
                double v = 0.5 * (G[i + (long long)n * j] + G[j + (long long)n * i]);
                G[i + (long long)n * j] = v; G[j + (long long)n * i] = v;
            }
        }
        __syncthreads();
        for (int i = tid; i < n; i += nt) G[i + (long long)n * i] += sigma;
    }
    __syncthreads();
    const int mm = (n + 1) & ~1;        // players in the tournament
    const int npairs = mm / 2;
    const int grp = tid / TG, gl = tid % TG, ngrp = nt / TG;
    const double tol = sqrt((double)n) * 2.220446049250313e-16;
    int sweep = 0;
    for (; sweep < 40; ++sweep) {
        int rotated = 0;
        for (int step = 0; step < mm - 1; ++step) {
            for (int pr0 = 0; pr0 < npairs; pr0 += ngrp) {
                // all lanes run the same trip count so the full-mask shuffles below stay convergent
                const int pr = pr0 + grp;
                int p = 0, q = n;
                if (pr < npairs) {
                    if (pr == 0) { p = step % (mm - 1); q = mm - 1; }
                    else { p = (step + pr) % (mm - 1); q = (step - pr + (mm - 1)) % (mm - 1); }
                    if (p > q) { int t = p; p = q; q = t; }
                }
                const bool active = (pr < npairs) && (q < n);
                double* gp = G + (long long)ld * p;
                double* gq = G + (long long)ld * (active ? q : p);
                double al = 0, be = 0, ga = 0;
                if (active)
                    for (int r = gl; r < n; r += TG) {
                        const double a = gp[r], b = gq[r];
                        al += a * a; be += b * b; ga += a * b;
                    }
#pragma unroll
                for (int o = TG / 2; o > 0; o >>= 1) {
                    al += __shfl_xor_sync(0xffffffffu, al, o);
                    be += __shfl_xor_sync(0xffffffffu, be, o);
                    ga += __shfl_xor_sync(0xffffffffu, ga, o);
                }
                if (active && fabs(ga) > tol * sqrt(al * be)) {
                    rotated = 1;
                    const double zeta = (be - al) / (2.0 * ga);
                    const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                    const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                    for (int r = gl; r < n; r += TG) {
                        const double a = gp[r], b = gq[r];
                        gp[r] = c * a - s * b;
                        gq[r] = s * a + c * b;
                    }
                }
            }
            __syncthreads();
        }
        if (!__syncthreads_or(rotated)) { ++sweep; break; }
    }
    // eigenvalues, ordering, normalised eigenvectors
    double* nrm = d.nrm + m.ao0;
    for (int j = tid; j < n; j += nt) {
        double s = 0;
        const double* g = G + (long long)ld * j;
        for (int r = 0; r < n; ++r) s += g[r] * g[r];
        nrm[j] = sqrt(s);
    }
    __syncthreads();
    double* Z = d.Zm + m.mat0;
    double* emo = d.emo + m.ao0;
    int* rank = d.rank + m.ao0;
    for (int j = tid; j < n; j += nt) {
        const double v = nrm[j];
        int rk = 0;
        for (int k = 0; k < n; ++k) { const double w = nrm[k]; rk += (w < v || (w == v && k < j)) ? 1 : 0; }
        emo[rk] = v - sigma;
        rank[j] = rk;
    }
    __syncthreads();
    for (long long t = tid; t < (long long)n * n; t += nt) {
        const int r = (int)(t % n), j = (int)(t / n);
        Z[r + (long long)n * rank[j]] = G[r + (long long)ld * j] / nrm[j];
    }
    if (tid == 0) { st.sweeps = sweep; st.sweeps_total += sweep; st.zvalid = 1; if (sweep >= 40) st.fail = 12; }
}

}  // namespace xtbb200
