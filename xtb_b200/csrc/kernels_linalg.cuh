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
//        2 -> B is upper triangular (T = H X): K stops at the last row of the tile's columns
//        3 -> op(A) = X^T with X upper triangular and the product is symmetric (W = X^T T): K stops at the last row
//             index of the tile, only tiles on or below the diagonal are computed and mirrored into the upper triangle
//        4 -> A is upper triangular (C = X Z): K starts at the first row of the tile
// skip_done: 1 = skip molecules that already left the SCC loop; 2 = additionally skip molecules without a valid
// eigenvector guess (warm-start product A' Z_prev); 3 = as 2, and only molecules that follow a leader (first
// diagonalisation of an ensemble, see k_lead_copy)
// GTT: CTA tile, 64 or 48 (2 x 2 warps of GTT/2 x GTT/2); 48 fits orders like 144 = 3 x 48 without the 44 % of padding a
// 64-wide tiling of 192 x 192 multiplies (the engine picks the tile with the smaller padded area for the batch)
template <bool TA, bool TB, int GTT>
// tcb / tce: only the tile columns [tcb, tce) are computed (one large system with the GEMM columns split over the GPUs)
__global__ void __launch_bounds__(128) k_gemm(Dev d, const double* __restrict__ Ag, const double* __restrict__ Bg,
                                              double* __restrict__ Cg, int kmode, int skip_done, int tcb = 0, int tce = 1 << 30) {
    const MolDesc m = d.mol[blockIdx.y];
    if (skip_done && d.st[blockIdx.y].done) return;
    if (skip_done >= 2 && !d.st[blockIdx.y].zvalid) return;
    if (skip_done == 3 && d.leader[blockIdx.y] == (int)blockIdx.y) return;
    const int n = m.nao;
    constexpr int GTL = GTT + 4, NI = GTT / 16;
    const int tiles = (n + GTT - 1) / GTT;
    if ((int)blockIdx.x >= tiles * tiles) return;
    const int r0 = (blockIdx.x % tiles) * GTT, c0 = (blockIdx.x / tiles) * GTT;
    if ((int)(blockIdx.x / tiles) < tcb || (int)(blockIdx.x / tiles) >= tce) return;
    if (kmode == 3 && r0 < c0) return;
    int K = n, kbeg = 0;
    if (kmode == 1) K = d.st[blockIdx.y].kocc;
    else if (kmode == 2) K = min(n, c0 + GTT);
    else if (kmode == 3) K = min(n, r0 + GTT);
    else if (kmode == 4) kbeg = r0;
    const double* A = Ag + m.mat0;
    const double* B = Bg + m.mat0;
    double* C = Cg + m.mat0;
    // operand tiles: [k][row] with pitch GTL when consecutive threads of the staging fetch consecutive rows, [row][k] with pitch
    // GK + 4 when they fetch consecutive k (TA / !TB): staging stores and m8n8k4 fragment loads are free of bank conflicts in both
    constexpr int PK = GK + 4;
    __shared__ double As[TA ? GTT * PK : GK * GTL], Bs[TB ? GK * GTL : GTT * PK];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wr = (warp & 1) * (GTT / 2), wc = (warp >> 1) * (GTT / 2);
    double acc[NI][NI][2];
#pragma unroll
    for (int i = 0; i < NI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    // software pipeline: the global loads of tile k+1 are in flight (registers) while tile k is multiplied from shared memory
    double ra[GTT * GK / 128], rb[GTT * GK / 128];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int u = 0; u < GTT * GK / 128; ++u) {
            const int t = tid + 128 * u;
            if (!TA) {
                const int r = t % GTT, kk = t / GTT, gr = r0 + r, gk = k0 + kk;
                ra[u] = (gr < n && gk < K) ? A[gr + (long long)n * gk] : 0.0;
            } else {
                const int kk = t % GK, r = t / GK, gr = r0 + r, gk = k0 + kk;
                ra[u] = (gr < n && gk < K) ? A[gk + (long long)n * gr] : 0.0;
            }
            if (!TB) {
                const int kk = t % GK, c = t / GK, gc = c0 + c, gk = k0 + kk;
                rb[u] = (gc < n && gk < K) ? B[gk + (long long)n * gc] : 0.0;
            } else {
                const int c = t % GTT, kk = t / GTT, gc = c0 + c, gk = k0 + kk;
                rb[u] = (gc < n && gk < K) ? B[gc + (long long)n * gk] : 0.0;
            }
        }
    };
    auto stage = [&]() {   // As[kk][r] = op(A)(r0+r, k0+kk), Bs[kk][c] = op(B)(k0+kk, c0+c)
#pragma unroll
        for (int u = 0; u < GTT * GK / 128; ++u) {
            const int t = tid + 128 * u;
            if (!TA) As[(t / GTT) * GTL + t % GTT] = ra[u]; else As[(t / GK) * PK + t % GK] = ra[u];
            if (!TB) Bs[(t / GK) * PK + t % GK] = rb[u]; else Bs[(t / GTT) * GTL + t % GTT] = rb[u];
        }
    };
    if (K > kbeg) fetch(kbeg);
    for (int k0 = kbeg; k0 < K; k0 += GK) {
        stage();
        __syncthreads();
        if (k0 + GK < K) fetch(k0 + GK);
#pragma unroll
        for (int ks = 0; ks < GK; ks += 4) {
            double af[NI], bf[NI];
#pragma unroll
            for (int i = 0; i < NI; ++i)
                af[i] = TA ? As[(wr + 8 * i + (lane >> 2)) * PK + ks + (lane & 3)] : As[(ks + (lane & 3)) * GTL + wr + 8 * i + (lane >> 2)];
#pragma unroll
            for (int j = 0; j < NI; ++j)
                bf[j] = TB ? Bs[(ks + (lane & 3)) * GTL + wc + 8 * j + (lane >> 2)] : Bs[(wc + 8 * j + (lane >> 2)) * PK + ks + (lane & 3)];
#pragma unroll
            for (int i = 0; i < NI; ++i)
#pragma unroll
                for (int j = 0; j < NI; ++j) dmma_884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        __syncthreads();
    }
    const bool mirror = kmode == 3 && r0 > c0;
#pragma unroll
    for (int i = 0; i < NI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) {
            int gr = r0 + wr + 8 * i + (lane >> 2);
            int gc = c0 + wc + 8 * j + 2 * (lane & 3);
            if (gr < n) {
                if (gc < n) C[gr + (long long)n * gc] = acc[i][j][0];
                if (gc + 1 < n) C[gr + (long long)n * (gc + 1)] = acc[i][j][1];
                if (mirror) {
                    if (gc < n) C[gc + (long long)n * gr] = acc[i][j][0];
                    if (gc + 1 < n) C[gc + 1 + (long long)n * gr] = acc[i][j][1];
                }
            }
        }
}

// upper triangle <- lower triangle of an n x n matrix (after the column blocks of a symmetric product computed on its lower
// tiles were gathered from the ranks that own them); thread per element.
// Only the tiles above the diagonal tiles are mirrored (gt = tile width of the product): the diagonal tiles were computed
// whole by the owner of their column block, exactly as in the unsplit run (their two triangles differ in the last bit).
__global__ void k_mirror_lower(double* __restrict__ A, int n, int gt) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)n * n) return;
    const int i = (int)(t % n), j = (int)(t / n);
    if (i / gt < j / gt) A[t] = A[j + (long long)n * i];
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

// ---------------------------------------------------------------------------------------------
// Block one-sided Jacobi on the FP64 tensor path (the ensemble-sized solver, nao <= ~164).
//
// Same mathematics as k_jacobi (Hestenes orthogonalisation of the columns of G = (A' + sigma I) Z_prev) but the
// columns are handled in blocks of 4: one warp owns one pair of blocks (8 columns, an n x 8 panel W in shared memory)
// per tournament step and
//   1. forms the 8x8 Gram matrix M = W^T W with DMMA m8n8k4 (A and B fragments are the same shared-memory element,
//      one conflict-free 8-byte load per lane and k-step),
//   2. skips the pair when every relative off-diagonal |M_pq|/sqrt(M_pp M_qq) is below tol,
//   3. otherwise diagonalises M by a two-sided cyclic Jacobi in per-warp shared memory (32 lanes = 4 disjoint
//      rotations x 8 rows), accumulating the 8x8 orthogonal J,
//   4. applies W <- W J with DMMA (2 k-steps per 8-row tile).
// Shared-memory traffic per rotated column pair drops ~10x against the scalar kernel, which is shared-memory bound.
// Layout: G column-major, leading dimension ld = nr + 4 (nr = n rounded up to 8) so that ld = 4 mod 8: the 4 rows x 4
// columns touched by a half-warp fall into 16 distinct 8-byte banks.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void jb_block_pair(int step, int pr, int nb, int& P, int& Q) {
    // step < nb - 1 and pr < nb / 2: the residues modulo nb - 1 by one conditional subtraction / addition (two integer divisions
    // by a run-time divisor at the head of every tournament step were 5 % of the kernel's stall samples)
    const int mm1 = nb - 1;
    int p, q;
    if (pr == 0) { p = step; q = mm1; }
    else {
        p = step + pr; if (p >= mm1) p -= mm1;
        q = step - pr; if (q < 0) q += mm1;
    }
    P = p < q ? p : q; Q = p < q ? q : p;
}


// 1/sqrt(x) for x in the normal float range: single-precision seed + two Newton steps (full double accuracy,
// no special-case branches)
__device__ __forceinline__ double jb_rsqrt(double x) {
    double y = (double)rsqrtf((float)x);
    const double h = 0.5 * x;
    y = y * (1.5 - h * y * y);
    y = y * (1.5 - h * y * y);
    return y;
}

// release the two column blocks of a finished rotation to the warps that need them in the next tournament step
__device__ __forceinline__ void jb_publish(volatile int* ver, int P, int Q, int v, int lane) {
    __threadfence_block();
    __syncwarp();
    if (lane == 0) { ver[P] = v; ver[Q] = v; }
}

// default of Dev::jb_qtol.  The columns of G are then orthogonal to ~1e-14 after the last sweep, but the rotation angle between
// two eigenvectors is larger than the non-orthogonality of their columns by (lambda + sigma) / (2 gap): at 1e-5 the density
// matrix carried errors of 1e-6 (energy still right to 3e-12, but gradients only to 1e-7 and numerical Hessians to 8e-6), and
// at 1e-6 (52 instead of 56 sweeps per single point) the numerical Hessian test still missed its 1e-6
constexpr double kJbQuadraticStop = 1.0e-7;
constexpr int kJbInnerSweeps = 1;  // inner sweeps per Gram block (early exit when a sweep rotates nothing)
constexpr int kJbScratch = 192;   // doubles of per-warp scratch: M and J, 8 x 8 with row stride 9 (jb_sweeps) or kJs (jb_sweeps2)
// jb_sweeps2: largest |sin| of the 16 P x Q rotation angles of a block pair up to which they are taken simultaneously from the
// Gram matrix as it stands (0: always the sequential rounds)
// jb_sweeps2: row tiles (8 rows each) per pass of the apply W <- W J.  Measured on the ensemble (144 AOs = 18 row tiles): 6 tiles
// per pass 13.81 ms per launch, 4 tiles 13.60 (the fifth pass multiplies two tiles of zeros), 3 tiles 13.2, 2 tiles 13.09, 1 tile
// 12.57: the warps of an SM sub-partition share its FP64 pipe and the shared-memory pipe best in small bursts
#ifndef XTBB200_JB_APPLY_TILES
#define XTBB200_JB_APPLY_TILES 1
#endif
constexpr int kJbApplyTiles = XTBB200_JB_APPLY_TILES;
#ifndef XTBB200_JB_SIMUL
#define XTBB200_JB_SIMUL 0.1
#endif
constexpr double kJbSimulMaxSin = XTBB200_JB_SIMUL;
constexpr int kJbLoadBatch = 8;   // global loads in flight per thread and operand while k_jacobi_blk fills its shared memory
constexpr int kJs = 12;           // jb_sweeps2: with lane = (idx = lane >> 2, slot = lane & 3) the P x Q rounds are free of bank conflicts

struct JbDims { int nr, ld, nb; };
__host__ __device__ inline JbDims jb_dims(int n) {
    JbDims r;
    r.nr = (n + 7) & ~7;
    r.ld = r.nr + 4;
    r.nb = (n + 3) / 4;
    if (r.nb & 1) r.nb++;
    if (r.nb < 2) r.nb = 2;
    return r;
}
inline size_t jb_smem_bytes(int n, int nwarps) {
    JbDims g = jb_dims(n);
    // matrix + per-warp scratch + the tracked 4x4 Gram matrix of every column block (jb_sweeps2)
    return ((size_t)g.ld * g.nb * 4 + (size_t)nwarps * kJbScratch + (size_t)16 * g.nb) * sizeof(double);
}

#ifdef XTBB200_JB_TIMING
#define JB_T(k) do { long long t_ = clock64(); jbt[k] += t_ - jbt0; jbt0 = t_; } while (0)
#else
#define JB_T(k)
#endif
// The sweep loop of the block Jacobi solver on a matrix G resident in shared memory (layout of jb_dims(n)); every
// thread of the CTA calls it.  Msh: this warp's 8x9 + 8x9 scratch; ver: 64 ints of shared memory.  Returns the number
// of sweeps executed (the last one is the verification sweep when converged; == maxsweeps means not converged).
__device__ __forceinline__ int jb_sweeps(double* G, double* Msh, volatile int* ver, int n, double tol, int inner, int maxsweeps,
                                         double qtol
#ifdef XTBB200_JB_TIMING
                                         , long long* jbt, long long& jbt0
#endif
) {
    const JbDims dm = jb_dims(n);
    const int nr = dm.nr, ld = dm.ld, nb = dm.nb;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    double* Jsh = Msh + 72;
    if (tid < 64) ver[tid] = 0;
    __syncthreads();
    const double tol2 = tol * tol;
    // quadratic stop: a sweep in which no pair was further from orthogonality than qtol leaves the columns orthogonal
    // to ~qtol^2 (cyclic Jacobi converges quadratically), so the verification sweep can be skipped.  0: off.
    const double qtol2 = qtol * qtol;
    const double tol_in2 = 2.220446049250313e-16 * 2.220446049250313e-16;
    const int nbp = nb / 2;
    const int l4 = lane >> 2, l3 = lane & 3;
    // inner tournament of the 8 panel columns: lane = (slot, idx); slot's pair at inner step s7 packed 6 bits per step
    const int slot = lane >> 3, idx = lane & 7;
    unsigned long long pairtab = 0;
    for (int s7 = 0; s7 < 7; ++s7) {
        int p, q;
        if (slot == 0) { p = s7; q = 7; }
        else { const int a = (s7 + slot) % 7, b = (s7 - slot + 7) % 7; p = a < b ? a : b; q = a < b ? b : a; }
        pairtab |= (unsigned long long)(p | (q << 3)) << (6 * s7);
    }
    int sweep = 0;
    int tstep = 0;     // tournament steps completed so far (all sweeps)
    for (; sweep < maxsweeps; ++sweep) {
        int rotated = 0, rotbig = 0;
        for (int step = 0; step < nb - 1; ++step, ++tstep) {
            for (int pr = warp; pr < nbp; pr += nwarps) {
                int P, Q;
                jb_block_pair(step, pr, nb, P, Q);
                // wait until both blocks have finished the previous step, work, then publish (see jb_publish)
                if (lane == 0) { while (ver[P] < tstep || ver[Q] < tstep) {} }
                __syncwarp();
                __threadfence_block();
                JB_T(5);
                if (4 * Q >= n) { jb_publish(ver, P, Q, tstep + 1, lane); continue; }   // dummy block (all-zero columns)
                JB_T(0);
                // ---- Gram matrix of the 8 panel columns
                const int gcol = l4 < 4 ? 4 * P + l4 : 4 * Q + l4 - 4;
                const double* wp = G + (size_t)ld * gcol + l3;
                double a0 = 0, a1 = 0, b0 = 0, b1 = 0, c0 = 0, c1 = 0, e0 = 0, e1 = 0;
                int k0 = 0;
                for (; k0 + 16 <= nr; k0 += 16) {
                    const double w0 = wp[k0], w1 = wp[k0 + 4], w2 = wp[k0 + 8], w3 = wp[k0 + 12];
                    dmma_884(a0, a1, w0, w0); dmma_884(b0, b1, w1, w1);
                    dmma_884(c0, c1, w2, w2); dmma_884(e0, e1, w3, w3);
                }
                for (; k0 < nr; k0 += 4) { const double w0 = wp[k0]; dmma_884(a0, a1, w0, w0); }
                const double m0 = (a0 + b0) + (c0 + e0), m1 = (a1 + b1) + (c1 + e1);
                __syncwarp();
                Msh[l4 * 9 + 2 * l3] = m0; Msh[l4 * 9 + 2 * l3 + 1] = m1;
                Jsh[l4 * 9 + 2 * l3] = (l4 == 2 * l3) ? 1.0 : 0.0; Jsh[l4 * 9 + 2 * l3 + 1] = (l4 == 2 * l3 + 1) ? 1.0 : 0.0;
                __syncwarp();
                JB_T(1);
                // ---- needs work?  (squared form of |M_pq| > tol sqrt(M_pp M_qq): no square roots on this path)
                int need = 0;
                {
                    const double dii = Msh[l4 * 10];
                    const int j0 = 2 * l3, j1 = 2 * l3 + 1;
                    if (j0 != l4) need |= (m0 * m0 > tol2 * dii * Msh[j0 * 10]) ? 1 : 0;
                    if (j1 != l4) need |= (m1 * m1 > tol2 * dii * Msh[j1 * 10]) ? 1 : 0;
                }
                if (!__any_sync(0xffffffffu, need)) { jb_publish(ver, P, Q, tstep + 1, lane); continue; }
                rotated = 1;
                if (qtol2 > 0.0) {
                    const double dii = Msh[l4 * 10];
                    const int j0 = 2 * l3, j1 = 2 * l3 + 1;
                    int big = 0;
                    if (j0 != l4) big |= (m0 * m0 > qtol2 * dii * Msh[j0 * 10]) ? 1 : 0;
                    if (j1 != l4) big |= (m1 * m1 > qtol2 * dii * Msh[j1 * 10]) ? 1 : 0;
                    if (__any_sync(0xffffffffu, big)) rotbig = 1;
                } else rotbig = 1;
                // ---- two-sided Jacobi on the 8x8 Gram matrix; lane = (pair slot, row/column index).
                // A bounded number of inner sweeps is enough: the outer loop re-forms every Gram matrix, so the inner
                // solve only sets the convergence rate, never the final accuracy.  M and J have row stride 9
                // (conflict-free column access).
                for (int isw = 0; isw < inner; ++isw) {
                    int anyrot = 0;
#pragma unroll
                    for (int s7 = 0; s7 < 7; ++s7) {
                        const int p = (int)(pairtab >> (6 * s7)) & 7, q = (int)(pairtab >> (6 * s7 + 3)) & 7;
                        const double app = Msh[p * 10], aqq = Msh[q * 10], apq = Msh[p * 9 + q];
                        double c = 1.0, s = 0.0;
                        if (apq * apq > tol_in2 * app * aqq) {
                            // rotation angle |theta| <= pi/4 from two reciprocal square roots (no divisions):
                            // cos 2t = |tau|/r, r = hypot(tau, 2 apq); c = sqrt((1 + cos 2t)/2); s = sgn(tau) apq/(r c)
                            const double tau = aqq - app;
                            const double ir = jb_rsqrt(tau * tau + 4.0 * apq * apq);
                            const double c2 = 0.5 + 0.5 * fabs(tau) * ir;
                            const double ic = jb_rsqrt(c2);
                            c = c2 * ic;
                            s = (tau >= 0.0 ? apq : -apq) * ir * ic;
                            anyrot = 1;
                        }
                        __syncwarp();
                        {   // columns p,q of M and J (row idx)
                            const double mp = Msh[idx * 9 + p], mq = Msh[idx * 9 + q];
                            const double jp = Jsh[idx * 9 + p], jq = Jsh[idx * 9 + q];
                            Msh[idx * 9 + p] = c * mp - s * mq; Msh[idx * 9 + q] = s * mp + c * mq;
                            Jsh[idx * 9 + p] = c * jp - s * jq; Jsh[idx * 9 + q] = s * jp + c * jq;
                        }
                        __syncwarp();
                        {   // rows p,q of M (column idx)
                            const double mp = Msh[p * 9 + idx], mq = Msh[q * 9 + idx];
                            Msh[p * 9 + idx] = c * mp - s * mq; Msh[q * 9 + idx] = s * mp + c * mq;
                        }
                        __syncwarp();
                    }
                    if (!__any_sync(0xffffffffu, anyrot)) break;
                }
                JB_T(2);
                // ---- W <- W J : four 8-row tiles in flight per iteration
                {
                    const double bj0 = Jsh[l3 * 9 + l4], bj1 = Jsh[(4 + l3) * 9 + l4];
                    const double* ap0 = G + (size_t)ld * (4 * P + l3) + l4;
                    const double* ap1 = G + (size_t)ld * (4 * Q + l3) + l4;
                    const int oc = (l3 < 2) ? 4 * P + 2 * l3 : 4 * Q + 2 * l3 - 4;
                    double* op = G + (size_t)ld * oc + l4;
                    for (int r0 = 0; r0 < nr; r0 += 32) {
                        double xa[4], xb[4], y0[4], y1[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const bool on = r0 + 8 * u < nr;
                            xa[u] = on ? ap0[r0 + 8 * u] : 0.0;
                            xb[u] = on ? ap1[r0 + 8 * u] : 0.0;
                            y0[u] = 0.0; y1[u] = 0.0;
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) dmma_884(y0[u], y1[u], xa[u], bj0);
#pragma unroll
                        for (int u = 0; u < 4; ++u) dmma_884(y0[u], y1[u], xb[u], bj1);
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (r0 + 8 * u < nr) { op[r0 + 8 * u] = y0[u]; op[r0 + 8 * u + ld] = y1[u]; }
                    }
                }
                jb_publish(ver, P, Q, tstep + 1, lane);
                JB_T(3);
            }
        }
        if (!__syncthreads_or(rotated)) { ++sweep; break; }
        if (!__syncthreads_or(rotbig)) { ++sweep; break; }
    }
    return sweep;
}

// ---------------------------------------------------------------------------------------------
// jb_sweeps2: the sweep loop of k_jacobi_blk, second form.  Same tournament, same dataflow synchronisation, same
// convergence logic as jb_sweeps; three changes that cut the work per block pair:
//   * Gram matrix: only the 4x4 cross block W_P^T W_Q is formed (18 DMMAs at n = 144 instead of 36: rows 0-3 of the A
//     fragment and columns 0-3 of the B fragment take one half of the k range, rows / columns 4-7 the other half, so
//     the two diagonal quadrants of the accumulator are the two half sums).  The 4x4 Gram matrix of each block with
//     itself is tracked in shared memory (Tsh): it is the corresponding corner of J^T M J of the block's previous
//     rotation.  The first tournament step of every sweep forms the full 8x8 Gram matrix and refreshes the tracked ones.
//   * inner solve: away from the first step of a sweep only the 16 rotations between a column of P and a column of Q
//     are done (4 rounds of 4 disjoint rotations instead of the 7 rounds of the full 8x8 tournament): the columns of a
//     block were orthogonalised against each other when the block was last rotated, what the later rotations put back
//     is of second order and is removed by the full tournament of the next sweep's first step.  Sweep counts are
//     unchanged (checked on the SCC sequences of the ensemble and on random matrices).
//   * rotation parameters: the kernel is bound by shared-memory wavefronts and by the FP64 pipe, which DMMA and the
//     scalar FP64 instructions share at the same flop rate (tools/microbench/fp64_pipes.cu: a DMMA occupies it 16
//     cycles per SM sub-partition, a DFMA 2.2).  The angle is evaluated in single precision and only the final
//     normalisation of (c, s) to unit length is done in double precision (7 FP64 instructions instead of 33): the
//     rotation is orthogonal to rounding for any angle, and an angle with 2^-22 relative error leaves 1e-7 of the
//     eliminated element, far below what the other rotations of the sweep put back.
// G must be scaled so that its column norms are O(1) (k_jacobi_blk divides by sigma): the angles are formed in float.
// Msh: this warp's kJbScratch doubles; Tsh: 16 doubles per column block.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void jb_rotation(double app, double aqq, double apq, double& c, double& s) {
    c = 1.0; s = 0.0;
    const float appf = (float)app, aqqf = (float)aqq, apqf = (float)apq;
    if (apqf * apqf > 4.930380657631324e-32f * appf * aqqf) {     // |apq| > eps sqrt(app aqq)
        // rotation angle (|theta| <= pi/4) in single precision: t = sgn(tau) 2 apq / (|tau| + hypot(tau, 2 apq))
        const float tf = (float)(aqq - app), af = 2.0f * apqf;
        const float s2 = fmaf(tf, tf, af * af);
        const float t = __fdividef(tf >= 0.0f ? af : -af, fabsf(tf) + s2 * rsqrtf(s2));
        const float cf = rsqrtf(fmaf(t, t, 1.0f));
        // (c, s) of unit length to double precision: c0^2 + s0^2 = 1 + e with |e| ~ 1e-7, scale by (1 + e)^(-1/2) to O(e^3)
        const double c0 = (double)cf, s0 = (double)(t * cf);
        const double e = fma(c0, c0, fma(s0, s0, -1.0));
        const double k = fma(e, fma(e, 0.375, -0.5), 1.0);
        c = c0 * k; s = s0 * k;
    }
}

__device__ __forceinline__ int jb_sweeps2(double* G, double* Msh, double* Tsh, volatile int* ver, int n, double tol, int inner,
                                          int maxsweeps, double qtol
#ifdef XTBB200_JB_TIMING
                                          , long long* jbt, long long& jbt0
#endif
) {
    const JbDims dm = jb_dims(n);
    const int nr = dm.nr, ld = dm.ld, nb = dm.nb;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    double* Jsh = Msh + 8 * kJs;
    if (tid < 64) ver[tid] = 0;
    __syncthreads();
    const double tol2 = tol * tol;
    const double qtol2 = qtol * qtol;
    const int nbp = nb / 2;
    const int l4 = lane >> 2, l3 = lane & 3;
    // inner rotations: lane = (slot, idx).  pairtab: slot's pair in round s7 of the full tournament of the 8 panel columns
    // (6 bits per round); crosstab: slot's pair (slot, 4 + (slot + r) % 4) in round r of the P x Q rotations
    const int slot = lane & 3, idx = lane >> 2;
    unsigned long long pairtab = 0;
    unsigned int crosstab = 0;
    for (int s7 = 0; s7 < 7; ++s7) {
        int p, q;
        if (slot == 0) { p = s7; q = 7; }
        else { const int a = (s7 + slot) % 7, b = (s7 - slot + 7) % 7; p = a < b ? a : b; q = a < b ? b : a; }
        pairtab |= (unsigned long long)(p | (q << 3)) << (6 * s7);
    }
    for (int r = 0; r < 4; ++r) crosstab |= (unsigned int)(slot | ((4 + (slot + r) % 4) << 3)) << (6 * r);
    // tracked block Gram: lane e -> block (e >> 4 ? Q : P), element (i, j)
    const int tb = lane >> 4, ti = (lane >> 2) & 3, tj = lane & 3;
    const int tm = (4 * tb + ti) * kJs + 4 * tb + tj;
    int sweep = 0;
    int tstep = 0;
    for (; sweep < maxsweeps; ++sweep) {
        int rotated = 0, rotbig = 0;
        for (int step = 0; step < nb - 1; ++step, ++tstep) {
            const bool full = step == 0;
            for (int pr = warp; pr < nbp; pr += nwarps) {
                int P, Q;
                jb_block_pair(step, pr, nb, P, Q);
                if (lane == 0) { while (ver[P] < tstep || ver[Q] < tstep) {} }
                __syncwarp();
                __threadfence_block();
                JB_T(5);
                if (4 * Q >= n && !full) { jb_publish(ver, P, Q, tstep + 1, lane); continue; }   // dummy block
                double* Tb = Tsh + 16 * (tb ? Q : P) + 4 * ti + tj;
                if (full) {
                    // ---- full 8x8 Gram matrix (A and B fragment are the same element)
                    const int gcol = l4 < 4 ? 4 * P + l4 : 4 * Q + l4 - 4;
                    const double* wp = G + (size_t)ld * gcol + l3;
                    double a0 = 0, a1 = 0, b0 = 0, b1 = 0, c0 = 0, c1 = 0, e0 = 0, e1 = 0;
                    int k0 = 0;
                    for (; k0 + 16 <= nr; k0 += 16) {
                        const double w0 = wp[k0], w1 = wp[k0 + 4], w2 = wp[k0 + 8], w3 = wp[k0 + 12];
                        dmma_884(a0, a1, w0, w0); dmma_884(b0, b1, w1, w1);
                        dmma_884(c0, c1, w2, w2); dmma_884(e0, e1, w3, w3);
                    }
                    for (; k0 < nr; k0 += 4) { const double w0 = wp[k0]; dmma_884(a0, a1, w0, w0); }
                    __syncwarp();
                    Msh[l4 * kJs + 2 * l3] = (a0 + b0) + (c0 + e0); Msh[l4 * kJs + 2 * l3 + 1] = (a1 + b1) + (c1 + e1);
                } else {
                    // ---- cross block only: lanes l4 < 4 take rows k0..k0+3, lanes l4 >= 4 rows k0+4..k0+7
                    const int cc = l4 & 3, hh = (l4 >> 2) * 4;
                    const double* pa = G + (size_t)ld * (4 * P + cc) + hh + l3;
                    const double* pb = G + (size_t)ld * (4 * Q + cc) + hh + l3;
                    double a0 = 0, a1 = 0, b0 = 0, b1 = 0, c0 = 0, c1 = 0;
                    int k0 = 0;
                    for (; k0 + 24 <= nr; k0 += 24) {
                        const double x0 = pa[k0], y0 = pb[k0], x1 = pa[k0 + 8], y1 = pb[k0 + 8], x2 = pa[k0 + 16], y2 = pb[k0 + 16];
                        dmma_884(a0, a1, x0, y0); dmma_884(b0, b1, x1, y1); dmma_884(c0, c1, x2, y2);
                    }
                    for (; k0 < nr; k0 += 8) { const double x0 = pa[k0], y0 = pb[k0]; dmma_884(a0, a1, x0, y0); }
                    double m0 = (a0 + b0) + c0, m1 = (a1 + b1) + c1;
                    m0 += __shfl_down_sync(0xffffffffu, m0, 18);
                    m1 += __shfl_down_sync(0xffffffffu, m1, 18);
                    const double tv = *Tb;
                    __syncwarp();
                    if (l4 < 4 && l3 < 2) {
                        Msh[l4 * kJs + 4 + 2 * l3] = m0; Msh[l4 * kJs + 5 + 2 * l3] = m1;
                        Msh[(4 + 2 * l3) * kJs + l4] = m0; Msh[(5 + 2 * l3) * kJs + l4] = m1;
                    }
                    Msh[tm] = tv;
                }
                __syncwarp();
                JB_T(1);
                // ---- needs work?  (squared form of |M_pq| > tol sqrt(M_pp M_qq))
                const int j0 = 2 * l3, j1 = 2 * l3 + 1;
                const double dii = Msh[l4 * (kJs + 1)], dj0 = Msh[j0 * (kJs + 1)], dj1 = Msh[j1 * (kJs + 1)];
                const double m0 = Msh[l4 * kJs + j0], m1 = Msh[l4 * kJs + j1];
                // the rotation of this lane's pair (simultaneous inner solve, see below: the lane of (slot, idx) evaluates the pair
                // of round idx & 3 of its slot -- the FP64 pipe is shared by all warps, which reach this point together, and four
                // rotations per lane were half of its load here) is evaluated ahead of the votes, so that its chain of dependent
                // instructions overlaps the test
                double c1 = 1.0, s1 = 0.0;
                if (!full && kJbSimulMaxSin > 0.0) {
                    const int q = 4 + ((slot + idx) & 3);
                    jb_rotation(Msh[slot * (kJs + 1)], Msh[q * (kJs + 1)], Msh[slot * kJs + q], c1, s1);
                }
                int need = 0, big = 0;
                {
                    const double d0 = dii * dj0, d1 = dii * dj1;
                    if (j0 != l4) { need |= (m0 * m0 > tol2 * d0) ? 1 : 0; big |= (m0 * m0 > qtol2 * d0) ? 1 : 0; }
                    if (j1 != l4) { need |= (m1 * m1 > tol2 * d1) ? 1 : 0; big |= (m1 * m1 > qtol2 * d1) ? 1 : 0; }
                }
                if (!__any_sync(0xffffffffu, need)) {
                    if (full) *Tb = Msh[tm];
                    jb_publish(ver, P, Q, tstep + 1, lane);
                    continue;
                }
                rotated = 1;
                if (qtol2 > 0.0) { if (__any_sync(0xffffffffu, big)) rotbig = 1; }
                else rotbig = 1;
                // ---- two-sided Jacobi on the 8x8 Gram matrix; lane = (pair slot, row/column index)
                __syncwarp();
                bool simul = false;
                if (!full && kJbSimulMaxSin > 0.0) {
                    // Small angles (every sweep but the first part of the first one): the 16 P x Q rotations with the angles of
                    // the Gram matrix as it stands -- they differ from the angles of the sequential rounds in second order, like
                    // what the rounds put back into each other's zeros.  Nothing of M has to be updated between the rounds then:
                    // J = product of the 16 (exactly orthogonal) rotations is accumulated in registers (the lane of (slot, idx)
                    // keeps J[idx][slot]; column 4 + (slot + r) % 4 comes from the lane of slot + 1 by one shuffle per round) and
                    // the tracked Gram blocks are the diagonal blocks of J^T M J, two 8 x 8 x 8 products = four DMMAs.  Shared-
                    // memory wavefronts of the inner solve: ~50 instead of ~125 per block pair (the phase counters show the
                    // inner solves of all warps in lock step: a third of the kernel with the FP64 pipe idle).
                    double cr[4], sr[4];
                    simul = __all_sync(0xffffffffu, fabs(s1) <= kJbSimulMaxSin);
                    JB_T(0);
                    if (simul) {
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            cr[r] = __shfl_sync(0xffffffffu, c1, 4 * r + slot);
                            sr[r] = __shfl_sync(0xffffffffu, s1, 4 * r + slot);
                        }
                        double jp = (idx == slot) ? 1.0 : 0.0, jq = (idx == 4 + slot) ? 1.0 : 0.0;
                        const int src = (lane & ~3) | ((lane + 1) & 3);
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            const double np = cr[r] * jp - sr[r] * jq, nq = sr[r] * jp + cr[r] * jq;
                            jp = np;
                            jq = r < 3 ? __shfl_sync(0xffffffffu, nq, src) : nq;
                        }
                        Jsh[idx * kJs + slot] = jp; Jsh[idx * kJs + 4 + ((slot + 3) & 3)] = jq;
                        __syncwarp();
                        JB_T(4);
                        // X = M J (A fragment M[l4][k], B fragment J[k][l4]), then M' = J^T X (A fragment J[k][l4], B fragment X[k][l4])
                        double x0 = 0.0, x1 = 0.0;
                        dmma_884(x0, x1, Msh[l4 * kJs + l3], Jsh[l3 * kJs + l4]);
                        dmma_884(x0, x1, Msh[l4 * kJs + 4 + l3], Jsh[(4 + l3) * kJs + l4]);
                        __syncwarp();
                        Msh[l4 * kJs + 2 * l3] = x0; Msh[l4 * kJs + 2 * l3 + 1] = x1;
                        __syncwarp();
                        double y0 = 0.0, y1 = 0.0;
                        dmma_884(y0, y1, Jsh[l3 * kJs + l4], Msh[l3 * kJs + l4]);
                        dmma_884(y0, y1, Jsh[(4 + l3) * kJs + l4], Msh[(4 + l3) * kJs + l4]);
                        __syncwarp();
                        Msh[l4 * kJs + 2 * l3] = y0; Msh[l4 * kJs + 2 * l3 + 1] = y1;
                        __syncwarp();
                    }
                }
                if (!simul) {
                Jsh[l4 * kJs + j0] = (l4 == j0) ? 1.0 : 0.0; Jsh[l4 * kJs + j1] = (l4 == j1) ? 1.0 : 0.0;
                __syncwarp();
                {
                    const unsigned long long tab = full ? pairtab : (unsigned long long)crosstab;
                    const int nround = full ? 7 : 4;
#pragma unroll 1
                    for (int s7 = 0; s7 < nround; ++s7) {
                        const int p = (int)(tab >> (6 * s7)) & 7, q = (int)(tab >> (6 * s7 + 3)) & 7;
                        double c, sn;
                        jb_rotation(Msh[p * (kJs + 1)], Msh[q * (kJs + 1)], Msh[p * kJs + q], c, sn);
                        __syncwarp();
                        {   // columns p,q of M and J (row idx)
                            const double mp = Msh[idx * kJs + p], mq = Msh[idx * kJs + q];
                            const double jp = Jsh[idx * kJs + p], jq = Jsh[idx * kJs + q];
                            Msh[idx * kJs + p] = c * mp - sn * mq; Msh[idx * kJs + q] = sn * mp + c * mq;
                            Jsh[idx * kJs + p] = c * jp - sn * jq; Jsh[idx * kJs + q] = sn * jp + c * jq;
                        }
                        __syncwarp();
                        {   // rows p,q of M (column idx)
                            const double mp = Msh[p * kJs + idx], mq = Msh[q * kJs + idx];
                            Msh[p * kJs + idx] = c * mp - sn * mq; Msh[q * kJs + idx] = sn * mp + c * mq;
                        }
                        __syncwarp();
                    }
                }
                }
                JB_T(2);
                // ---- tracked Gram matrices of the two blocks
                *Tb = Msh[tm];
                // ---- W <- W J
                {
                    const double bj0 = Jsh[l3 * kJs + l4], bj1 = Jsh[(4 + l3) * kJs + l4];
                    const double* ap0 = G + (size_t)ld * (4 * P + l3) + l4;
                    const double* ap1 = G + (size_t)ld * (4 * Q + l3) + l4;
                    const int oc = (l3 < 2) ? 4 * P + 2 * l3 : 4 * Q + 2 * l3 - 4;
                    double* op = G + (size_t)ld * oc + l4;
                    for (int r0 = 0; r0 < nr; r0 += 8 * kJbApplyTiles) {
                        double xa[kJbApplyTiles], xb[kJbApplyTiles], y0[kJbApplyTiles], y1[kJbApplyTiles];
#pragma unroll
                        for (int u = 0; u < kJbApplyTiles; ++u) {
                            const bool on = r0 + 8 * u < nr;
                            xa[u] = on ? ap0[r0 + 8 * u] : 0.0;
                            xb[u] = on ? ap1[r0 + 8 * u] : 0.0;
                            y0[u] = 0.0; y1[u] = 0.0;
                        }
#pragma unroll
                        for (int u = 0; u < kJbApplyTiles; ++u) dmma_884(y0[u], y1[u], xa[u], bj0);
#pragma unroll
                        for (int u = 0; u < kJbApplyTiles; ++u) dmma_884(y0[u], y1[u], xb[u], bj1);
                        // lanes l3 < 2 store their first column while lanes l3 >= 2 store their second one and vice versa: the
                        // columns of P and Q start in the same bank, so this keeps every store free of bank conflicts
#pragma unroll
                        for (int u = 0; u < kJbApplyTiles; ++u)
                            if (r0 + 8 * u < nr) {
                                op[r0 + 8 * u + (l3 < 2 ? 0 : ld)] = l3 < 2 ? y0[u] : y1[u];
                                op[r0 + 8 * u + (l3 < 2 ? ld : 0)] = l3 < 2 ? y1[u] : y0[u];
                            }
                    }
                }
                jb_publish(ver, P, Q, tstep + 1, lane);
                JB_T(3);
            }
        }
        if (!__syncthreads_or(rotated)) { ++sweep; break; }
        if (!__syncthreads_or(rotbig)) { ++sweep; break; }
    }
    return sweep;
}

// Conformer ensembles: the first diagonalisation of a molecule has no eigenvectors of its own to start from, but the
// converged ones of another conformer of the same molecule are a far better start than the unit matrix (any orthogonal
// Z is a valid start of the one-sided method).  The first molecule of every group of identical compositions (its
// "leader") is solved cold, its Z is copied to the followers, and they start warm.  thread per matrix element
__global__ void k_lead_copy(Dev d) {
    const int im = blockIdx.y, L = d.leader[im];
    if (L == im) return;
    const MolState& sl = d.st[L];
    if (!sl.zvalid || sl.fail) return;
    const MolDesc m = d.mol[im];
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)m.nao * m.nao) return;
    d.Zm[m.mat0 + t] = d.Zm[d.mol[L].mat0 + t];
    if (t == 0) d.st[im].zvalid = 1;
}

// phase: 0 all molecules, 1 leaders only, 2 followers only.  V: 0 = jb_sweeps, 1 = jb_sweeps2 (G scaled by 1/sigma)
// at most kJbMaxWarps warps (one per block pair of the largest order, 152 = 38 blocks); with the bound of 608 instead of 640
// threads ptxas keeps everything in registers (no stack), measured 0.8 % on the ensemble step (15.81 -> 15.60 ms per launch)
constexpr int kJbMaxWarps = 19;
template <int V>
__global__ void __launch_bounds__(32 * kJbMaxWarps) k_jacobi_blk(Dev d, int phase) {
#ifdef XTBB200_JB_TIMING
    long long jbt[6] = {0, 0, 0, 0, 0, 0}, jbt0 = clock64();
#endif
    const int im = blockIdx.x;
    const MolDesc m = d.mol[im];
    MolState& st = d.st[im];
    if (st.done) return;
    if (phase && (d.leader[im] == im) != (phase == 1)) return;
    const int n = m.nao;
    const JbDims dm = jb_dims(n);
    const int nr = dm.nr, ld = dm.ld, nb = dm.nb;
    extern __shared__ double smem[];
    double* G = smem;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    double* Msh = smem + (size_t)ld * nb * 4 + (size_t)warp * kJbScratch;
    double* Jsh = Msh + 72;
    __shared__ double red[32];
    __shared__ double s_sigma;
    // ver[b] = number of tournament steps column block b has completed.  Warps synchronise on the two blocks they are
    // about to rotate instead of a CTA-wide barrier per step, so the DMMA phases (Gram, apply) of some warps overlap
    // the latency-bound inner solves of others.
    __shared__ volatile int ver[64];
    const double* A = d.W + m.mat0;
    // Gershgorin bound -> shift that makes G positive definite.  A' is staged in the (still unused) shared memory of G by
    // all threads, kJbLoadBatch loads in flight per thread, and the row sums are formed from there in the order j = 0 .. n-1
    // (same sums as a loop over global memory per row thread).  One CTA per SM: nothing else hides the HBM latency of this
    // prologue -- but it is a small part of the ~570 us a matrix spends on its SM: batching the loads did not change the
    // launch time measurably (15.81 ms over 4096 matrices before and after)
    {
        const int total = n * n;
        for (int t0 = tid; t0 < total; t0 += nt * kJbLoadBatch) {
            double v[kJbLoadBatch];
#pragma unroll
            for (int u = 0; u < kJbLoadBatch; ++u) {
                const int t = t0 + u * nt;
                v[u] = t < total ? A[t] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < kJbLoadBatch; ++u) {
                const int t = t0 + u * nt;
                if (t < total) G[(t % n) + (size_t)ld * (t / n)] = v[u];
            }
        }
    }
    __syncthreads();
    double mx = 0.0;
    for (int i = tid; i < n; i += nt) {
        double s = 0;
        for (int j = 0; j < n; ++j) s += fabs(G[i + (size_t)ld * j]);
        mx = fmax(mx, s);
    }
    mx = warp_max(mx);
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    if (tid == 0) {
        double v = 0;
        for (int w = 0; w < nwarps; ++w) v = fmax(v, red[w]);
        s_sigma = 1.05 * v + 1.0;
    }
    __syncthreads();
    const double sigma = s_sigma;
    const double gscale = V == 1 ? 1.0 / sigma : 1.0;    // V = 1: columns of G of norm O(1), eigenvalue = sigma (norm - 1)
    const int ncolp = nb * 4;
    const bool warm = st.zvalid != 0;
    {
        // warm: G = T + sigma Z_prev; cold: the symmetrised A' + sigma I.  Both operands of kJbLoadBatch elements are
        // requested before the first one is used (SASS of the one-element-per-iteration form: two loads, then the
        // dependent DFMA and the store, 37 times at n = 144); same arithmetic per element, identical energies
        const double* T = d.T + m.mat0;
        const double* Zp = d.Zm + m.mat0;
        const double* p0 = warm ? T : A;
        const double* p1 = warm ? Zp : A;
        const int total = ld * ncolp;
        for (int t0 = tid; t0 < total; t0 += nt * kJbLoadBatch) {
            double x[kJbLoadBatch], y[kJbLoadBatch];
#pragma unroll
            for (int u = 0; u < kJbLoadBatch; ++u) {
                const int t = t0 + u * nt, i = t % ld, j = t / ld;
                const bool ok = i < n && j < n;     // (t >= total implies j >= ncolp >= n)
                x[u] = ok ? p0[i + (long long)n * j] : 0.0;
                y[u] = ok ? p1[warm ? i + (long long)n * j : j + (long long)n * i] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < kJbLoadBatch; ++u) {
                const int t = t0 + u * nt, i = t % ld, j = t / ld;
                if (t < total) {
                    double v = 0.0;
                    if (i < n && j < n) {
                        if (warm) v = fma(sigma, y[u], x[u]);
                        else v = fma(x[u] + y[u], 0.5, i == j ? sigma : 0.0);
                        v *= gscale;
                    }
                    G[t] = v;
                }
            }
        }
    }
    const double tol = d.jb_tol > 0.0 ? d.jb_tol : 8.0 * sqrt((double)n) * 2.220446049250313e-16;
    int sweep;
    if (V == 1) {
        double* Tsh = smem + (size_t)ld * nb * 4 + (size_t)nwarps * kJbScratch;
#ifdef XTBB200_JB_TIMING
        sweep = jb_sweeps2(G, Msh, Tsh, ver, n, tol, d.jb_inner, 40, d.jb_qtol, jbt, jbt0);
#else
        sweep = jb_sweeps2(G, Msh, Tsh, ver, n, tol, d.jb_inner, 40, d.jb_qtol);
#endif
    } else {
#ifdef XTBB200_JB_TIMING
        sweep = jb_sweeps(G, Msh, ver, n, tol, d.jb_inner, 40, d.jb_qtol, jbt, jbt0);
#else
        sweep = jb_sweeps(G, Msh, ver, n, tol, d.jb_inner, 40, d.jb_qtol);
#endif
    }
    // eigenvalues, ordering, normalised eigenvectors
    double* nrm = d.nrm + m.ao0;
    for (int j = tid; j < n; j += nt) {
        double s = 0;
        const double* g = G + (size_t)ld * j;
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
        emo[rk] = V == 1 ? sigma * (v - 1.0) : v - sigma;
        rank[j] = rk;
    }
    __syncthreads();
    for (int j = warp; j < n; j += nwarps) {      // warp per column: no index divisions, coalesced stores
        const double nj = nrm[j];
        double* zc = Z + (long long)n * rank[j];
        const double* gc = G + (size_t)ld * j;
        for (int r = lane; r < n; r += 32) zc[r] = gc[r] / nj;
    }
    if (tid == 0) { st.sweeps = sweep; st.sweeps_total += sweep; st.zvalid = 1; if (sweep >= 40) st.fail = 12; }
#ifdef XTBB200_JB_TIMING
    if (im == 0 && (tid == 0 || tid == 32 * 7))
        printf("jb timing warp %d sweeps %d: test+angles %lld gram %lld products(+sequential rounds) %lld apply %lld rounds %lld wait %lld\n", warp, sweep,
               jbt[0], jbt[1], jbt[2], jbt[3], jbt[4], jbt[5]);
#endif
}

}  // namespace xtbb200
