// Two-level block Jacobi eigensolver for matrices that do not fit into shared memory (crambin, water clusters).
//
// Same one-sided (Hestenes) scheme as k_jacobi_blk, one level up: the columns of G = (A' + sigma I) Z_prev live in
// HBM in blocks of 64; per tournament step every pair of blocks (an n x 128 panel W) is processed by three batched
// kernels, all on the FP64 tensor path (DMMA m8n8k4):
//   k_hj_gram   M = W^T W                      (four 64x64 tiles, K = n)
//   k_hj_inner  M J = J diag  (128 x 128, k_jacobi_blk's shared-memory sweeps; skipped when M is already diagonal to tol)
//   k_hj_apply  W <- W J                        (in place, row tile staged in shared memory)
// All pairs of a step are independent, so a step is three launches over (pairs x molecules); a sweep is nb-1 steps.
// The host reads one counter per sweep (number of pairs that still needed a rotation).
//
// Replaces dsyevd for large n (src/mctc/lapack/eigensolve.F90:218-246) and, applied to S, the Cholesky factorisation
// (src/scc_core.f90:411-413): X = V s^-1/2 is used instead of U^-1; both satisfy X^T S X = 1 and give the same
// density matrix.
#pragma once
#include "common.cuh"
#include "kernels_linalg.cuh"

namespace xtbb200 {

constexpr int HB = 64;            // outer block: columns
constexpr int HP = 2 * HB;        // panel columns
constexpr int kHjInnerWarps = 16; // 128 / 4 / 2 block pairs of the inner solver

__host__ __device__ inline int hj_nblocks(int n) {
    int nb = (n + HB - 1) / HB;
    if (nb & 1) nb++;
    return nb < 2 ? 2 : nb;
}

// 64x64 output tile: acc += op(A)(r0.., k) * op(B)(k, c0..), k < K; rows valid < mrows, cols valid < ncols.
// op(A)(r,k) = TA ? A[k + lda r] : A[r + lda k];  op(B)(k,c) = TB ? B[c + ldb k] : B[k + ldb c]
template <bool TA, bool TB>
__device__ __forceinline__ void gemm64_acc(double (&acc)[4][4][2], const double* __restrict__ A, long long lda,
                                           const double* __restrict__ B, long long ldb, int mrows, int ncols, int K,
                                           double (*As)[GLD], double (*Bs)[GLD]) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wr = (warp & 1) * 32, wc = (warp >> 1) * 32;
    double ra[GT * GK / 128], rb[GT * GK / 128];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int u = 0; u < GT * GK / 128; ++u) {
            const int t = tid + 128 * u;
            if (!TA) { const int r = t % GT, kk = t / GT, gk = k0 + kk; ra[u] = (r < mrows && gk < K) ? A[r + lda * gk] : 0.0; }
            else { const int kk = t % GK, r = t / GK, gk = k0 + kk; ra[u] = (r < mrows && gk < K) ? A[gk + lda * r] : 0.0; }
            if (!TB) { const int kk = t % GK, c = t / GK, gk = k0 + kk; rb[u] = (c < ncols && gk < K) ? B[gk + ldb * c] : 0.0; }
            else { const int c = t % GT, kk = t / GT, gk = k0 + kk; rb[u] = (c < ncols && gk < K) ? B[c + ldb * gk] : 0.0; }
        }
    };
    auto stage = [&]() {
#pragma unroll
        for (int u = 0; u < GT * GK / 128; ++u) {
            const int t = tid + 128 * u;
            if (!TA) As[t / GT][t % GT] = ra[u]; else As[t % GK][t / GK] = ra[u];
            if (!TB) Bs[t % GK][t / GK] = rb[u]; else Bs[t / GT][t % GT] = rb[u];
        }
    };
    if (K > 0) fetch(0);
    for (int k0 = 0; k0 < K; k0 += GK) {
        stage();
        __syncthreads();
        if (k0 + GK < K) fetch(k0 + GK);
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
}

// Gershgorin bound of the source matrix -> shift.  sigma = 2 g + 1 keeps the condition number of G below 3, so the
// Gram matrices of the panels (condition < 9) carry rotations accurate to a few eps.  One CTA per molecule.
__global__ void __launch_bounds__(256) k_hj_sigma(Dev d, const double* __restrict__ Asrc) {
    const int im = blockIdx.x;
    const MolDesc m = d.mol[im];
    if (d.st[im].done) return;
    const int n = m.nao;
    const double* A = Asrc + m.mat0;
    __shared__ double red[8];
    double mx = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double s = 0;
        for (int j = 0; j < n; ++j) s += fabs(A[i + (long long)n * j]);
        mx = fmax(mx, s);
    }
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        double v = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v = fmax(v, red[w]);
        d.hj_sigma[im] = 2.0 * v + 1.0;
    }
}

// G = (A + sigma I) Z_prev = T + sigma Z_prev (warm) or the symmetrised A + sigma I (cold); thread per element
__global__ void k_hj_load(Dev d, const double* __restrict__ Asrc, int allow_warm) {
    const int im = blockIdx.y;
    const MolDesc m = d.mol[im];
    if (d.st[im].done) return;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int n = m.nao;
    if (t >= (long long)n * n) return;
    const int i = (int)(t % n), j = (int)(t / n);
    const double sigma = d.hj_sigma[im];
    const double* A = Asrc + m.mat0;
    double v;
    if (allow_warm && d.st[im].zvalid) v = d.T[m.mat0 + t] + sigma * d.Zm[m.mat0 + t];
    else v = 0.5 * (A[t] + A[j + (long long)n * i]) + (i == j ? sigma : 0.0);
    d.hjG[m.mat0 + t] = v;
}

// M_p = W_p^T W_p for every block pair of this tournament step.  grid = (4 tiles, pairs, molecules), 128 threads
__global__ void __launch_bounds__(128) k_hj_gram(Dev d, int step) {
    const int im = blockIdx.z;
    const MolDesc m = d.mol[im];
    if (d.st[im].done || d.hj_conv[im]) return;
    const int n = m.nao, nb = hj_nblocks(n), pr = blockIdx.y;
    if (pr >= nb / 2) return;
    int P, Q;
    jb_block_pair(step, pr, nb, P, Q);
    const int ta = blockIdx.x & 1, tb = blockIdx.x >> 1;
    const int ca = HB * (ta ? Q : P), cb = HB * (tb ? Q : P);
    __shared__ double As[GK][GLD], Bs[GK][GLD];
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    const double* G = d.hjG + m.mat0;
    const int ma = max(0, min(HB, n - ca)), nbv = max(0, min(HB, n - cb));
    if (ma > 0 && nbv > 0)
        gemm64_acc<true, false>(acc, G + (long long)n * ca, n, G + (long long)n * cb, n, ma, nbv, n, As, Bs);
    double* M = d.hjM + m.hj0 + (long long)pr * HP * HP;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wr = (warp & 1) * 32, wc = (warp >> 1) * 32;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = ta * HB + wr + 8 * i + (lane >> 2), c = tb * HB + wc + 8 * j + 2 * (lane & 3);
            M[r + HP * c] = acc[i][j][0];
            M[r + HP * (c + 1)] = acc[i][j][1];
        }
}

// Eigenvectors J_p of the 128 x 128 Gram matrix of every block pair that still needs a rotation.
// grid = (pairs, molecules), 32 * kHjInnerWarps threads, dynamic shared memory jb_smem_bytes(128, kHjInnerWarps)
__global__ void __launch_bounds__(32 * kHjInnerWarps) k_hj_inner(Dev d, int step) {
    const int im = blockIdx.y;
    const MolDesc m = d.mol[im];
    if (d.st[im].done || d.hj_conv[im]) return;
    const int n = m.nao, nb = hj_nblocks(n), pr = blockIdx.x;
    if (pr >= nb / 2) return;
    int* flag = d.hj_flag + (m.hj0 / (HP * HP)) + pr;
    const double* M = d.hjM + m.hj0 + (long long)pr * HP * HP;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
    // does any pair of panel columns violate |M_ij| <= tol sqrt(M_ii M_jj) ?
    // the outer threshold sits well above what one inner solve + DMMA apply leaves behind (~1e-14), otherwise the
    // sweeps would chase rounding noise
    const double tol = d.jb_tol > 0.0 ? d.jb_tol : 32.0 * sqrt((double)n) * 2.220446049250313e-16;
    const double tol2 = tol * tol;
    int need = 0;
    for (int t = tid; t < HP * HP; t += nt) {
        const int i = t % HP, j = t / HP;
        if (i < j) { const double v = M[t]; need |= (v * v > tol2 * M[i * (HP + 1)] * M[j * (HP + 1)]) ? 1 : 0; }
    }
    need = __syncthreads_or(need);
    if (!need) { if (tid == 0) *flag = 0; return; }
    if (tid == 0) { *flag = 1; atomicAdd(d.hj_count + im, 1); }
    extern __shared__ double smem[];
    const JbDims dm = jb_dims(HP);
    double* G = smem;
    double* Msh = smem + (size_t)dm.ld * dm.nb * 4 + (size_t)warp * kJbScratch;
    __shared__ volatile int ver[64];
    __shared__ double s_nrm[HP];
    for (int t = tid; t < dm.ld * dm.nb * 4; t += nt) {
        const int i = t % dm.ld, j = t / dm.ld;
        G[t] = (i < HP && j < HP) ? 0.5 * (M[i + HP * j] + M[j + HP * i]) : 0.0;
    }
    __syncthreads();
    const double tol_inner = 2.0 * sqrt((double)HP) * 2.220446049250313e-16;
#ifdef XTBB200_JB_TIMING
    long long jbt[6] = {0, 0, 0, 0, 0, 0}, jbt0 = clock64();
    jb_sweeps(G, Msh, ver, HP, tol_inner, 1, 60, 0.0, jbt, jbt0);
#else
    jb_sweeps(G, Msh, ver, HP, tol_inner, 1, 60, 0.0);
#endif
    __syncthreads();
    for (int j = tid; j < HP; j += nt) {
        double s = 0;
        for (int r = 0; r < HP; ++r) { const double v = G[r + (size_t)dm.ld * j]; s += v * v; }
        s_nrm[j] = sqrt(s);
    }
    __syncthreads();
    double* J = d.hjJ + m.hj0 + (long long)pr * HP * HP;
    for (int t = tid; t < HP * HP; t += nt) {
        const int r = t % HP, j = t / HP;
        const double nj = s_nrm[j];
        J[t] = nj > 0.0 ? G[r + (size_t)dm.ld * j] / nj : (r == j ? 1.0 : 0.0);   // zero (padding) columns stay put
    }
}

// W_p <- W_p J_p in place.  grid = (row tiles, pairs, molecules), 128 threads, dynamic smem HP * GLD doubles
__global__ void __launch_bounds__(128) k_hj_apply(Dev d, int step) {
    const int im = blockIdx.z;
    const MolDesc m = d.mol[im];
    if (d.st[im].done || d.hj_conv[im]) return;
    const int n = m.nao, nb = hj_nblocks(n), pr = blockIdx.y;
    if (pr >= nb / 2) return;
    const int r0 = blockIdx.x * GT;
    if (r0 >= n) return;
    if (!d.hj_flag[(m.hj0 / (HP * HP)) + pr]) return;
    int P, Q;
    jb_block_pair(step, pr, nb, P, Q);
    extern __shared__ double smem[];
    double (*Af)[GLD] = reinterpret_cast<double (*)[GLD]>(smem);   // Af[k][r] : panel column k, row r of this tile
    __shared__ double Bs[GK][GLD];
    double* G = d.hjG + m.mat0;
    const double* J = d.hjJ + m.hj0 + (long long)pr * HP * HP;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wr = (warp & 1) * 32, wc = (warp >> 1) * 32;
    for (int t = tid; t < HP * GT; t += 128) {
        const int r = t % GT, k = t / GT;
        const int col = (k < HB) ? HB * P + k : HB * Q + (k - HB);
        Af[k][r] = (r0 + r < n && col < n) ? G[(r0 + r) + (long long)n * col] : 0.0;
    }
    __syncthreads();
    for (int tb = 0; tb < 2; ++tb) {
        double acc[4][4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        const double* B = J + (long long)HP * (tb * HB);   // B(k, c) = J[k + HP (tb*64 + c)]
        for (int k0 = 0; k0 < HP; k0 += GK) {
            for (int t = tid; t < GT * GK; t += 128) { const int kk = t % GK, c = t / GK; Bs[kk][c] = B[(k0 + kk) + (long long)HP * c]; }
            __syncthreads();
#pragma unroll
            for (int ks = 0; ks < GK; ks += 4) {
                double af[4], bf[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) af[i] = Af[k0 + ks + (lane & 3)][wr + 8 * i + (lane >> 2)];
#pragma unroll
                for (int j = 0; j < 4; ++j) bf[j] = Bs[ks + (lane & 3)][wc + 8 * j + (lane >> 2)];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) dmma_884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
            }
            __syncthreads();
        }
        const int cbase = HB * (tb ? Q : P);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int gr = r0 + wr + 8 * i + (lane >> 2), gc = cbase + wc + 8 * j + 2 * (lane & 3);
                if (gr < n) {
                    if (gc < n) G[gr + (long long)n * gc] = acc[i][j][0];
                    if (gc + 1 < n) G[gr + (long long)n * (gc + 1)] = acc[i][j][1];
                }
            }
    }
}

// per sweep: molecules without any rotated pair are converged
__global__ void k_hj_sweep_end(Dev d, int* still) {
    const int im = blockIdx.x * blockDim.x + threadIdx.x;
    if (im >= d.nmol) return;
    if (d.st[im].done || d.hj_conv[im]) return;
    d.st[im].sweeps++;
    if (d.hj_count[im] == 0) d.hj_conv[im] = 1;
    else atomicAdd(still, 1);
    d.hj_count[im] = 0;
}
__global__ void k_hj_begin(Dev d) {
    const int im = blockIdx.x * blockDim.x + threadIdx.x;
    if (im >= d.nmol) return;
    d.hj_conv[im] = 0; d.hj_count[im] = 0; d.st[im].sweeps = 0;
}

// column norms (warp per column) -> shifted eigenvalues
__global__ void k_hj_norms(Dev d) {
    const int im = blockIdx.y;
    const MolDesc m = d.mol[im];
    if (d.st[im].done) return;
    const int n = m.nao;
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (j >= n) return;
    const double* g = d.hjG + m.mat0 + (long long)n * j;
    double s = 0;
    for (int r = lane; r < n; r += 32) s += g[r] * g[r];
    s = warp_sum(s);
    if (lane == 0) d.nrm[m.ao0 + j] = sqrt(s);
}
// rank of every eigenvalue (thread per column) and sorted eigenvalues
__global__ void k_hj_rank(Dev d) {
    const int im = blockIdx.y;
    const MolDesc m = d.mol[im];
    if (d.st[im].done) return;
    const int n = m.nao;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const double* nrm = d.nrm + m.ao0;
    const double v = nrm[j];
    int rk = 0;
    for (int k = 0; k < n; ++k) { const double w = nrm[k]; rk += (w < v || (w == v && k < j)) ? 1 : 0; }
    d.emo[m.ao0 + rk] = v - d.hj_sigma[im];
    d.rank[m.ao0 + j] = rk;
}
// Z[:, rank[j]] = G[:, j] / |G[:, j]| ; thread per element
__global__ void k_hj_write(Dev d, int sweeps_fail) {
    const int im = blockIdx.y;
    const MolDesc m = d.mol[im];
    MolState& st = d.st[im];
    if (st.done) return;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int n = m.nao;
    if (t >= (long long)n * n) return;
    const int r = (int)(t % n), j = (int)(t / n);
    d.Zm[m.mat0 + r + (long long)n * d.rank[m.ao0 + j]] = d.hjG[m.mat0 + t] / d.nrm[m.ao0 + j];
    if (t == 0) { st.sweeps_total += st.sweeps; st.zvalid = 1; if (!d.hj_conv[im] || sweeps_fail) st.fail = 12; }
}

// X = V s^-1/2 from the eigen-decomposition of the overlap matrix (Zm = V, emo = s); thread per element
__global__ void k_hj_xscale(Dev d) {
    const int im = blockIdx.y;
    const MolDesc m = d.mol[im];
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int n = m.nao;
    if (t >= (long long)n * n) return;
    const int j = (int)(t / n);
    const double s = d.emo[m.ao0 + j];
    if (!(s > 0.0)) { if (t % n == 0) d.st[im].fail = 11; return; }
    d.X[m.mat0 + t] = d.Zm[m.mat0 + t] * rsqrt(s);
    if (t == 0) d.st[im].zvalid = 0;   // Zm holds overlap eigenvectors: not a start vector for the Fock matrix
}

}  // namespace xtbb200
