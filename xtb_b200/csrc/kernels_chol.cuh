// Blocked Cholesky factorisation of the overlap matrix and inverse of the factor for matrices that do not fit into
// shared memory (crambin, water clusters):  S = L L^T,  X = L^-T  (so that X^T S X = 1), all GEMM-shaped work on the
// FP64 tensor path (DMMA m8n8k4 through gemm64_acc).
//
// Replaces dpotrf (src/scc_core.f90:411-413, via src/mctc/lapack/trf.f90) and the triangular solves that the reference
// does with dsygst / dtrsm in every iteration (src/mctc/lapack/eigensolve.F90:230-244): here the inverse factor is
// formed once per geometry and the reduction to standard form is two GEMMs per iteration.
//
// Right-looking factorisation in 64-column panels, three launches per panel:
//   k_bchol_diag    L_kk = chol(A_kk) and D_k = L_kk^-1 in shared memory (one CTA per molecule)
//   k_bchol_panel   L_ik = A_ik D_k^T                       (one CTA per row block below the diagonal)
//   k_bchol_update  A_ij -= L_ik L_jk^T, i >= j > k         (one CTA per 64x64 tile of the trailing lower triangle)
// Inverse by block forward substitution: with Lt_ik = D_i L_ik (k_bchol_scale, one launch),
//   Linv_jj = D_j,   Linv_ij = - sum_{j <= k < i} Lt_ik Linv_kj   (k_bchol_inv_row: one launch per block row i, one CTA
//   per block column j < i),  written transposed: X = Linv^T (upper triangular).
// Storage: L overwrites the lower triangle of d.W (copy of S), the D_k blocks live in d.T, the result in d.X.
#pragma once
#include "common.cuh"
#include "kernels_hjacobi.cuh"

namespace xtbb200 {

constexpr int CB = 64;   // panel width = gemm64_acc tile
constexpr int kBlockedCholMin = 97;   // smaller matrices: one CTA per molecule in shared memory (k_chol_inv) is faster
constexpr size_t kBcholDiagSmem = 2 * CB * (CB + 1) * sizeof(double);

__device__ __forceinline__ void bchol_zero(double (&acc)[4][4][2]) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
}
// visit the elements of the 64x64 accumulator tile owned by this thread: f(row, col, value)
template <class F>
__device__ __forceinline__ void bchol_foreach(const double (&acc)[4][4][2], F f) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wr = (warp & 1) * 32, wc = (warp >> 1) * 32;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = wr + 8 * i + (lane >> 2), c = wc + 8 * j + 2 * (lane & 3);
            f(r, c, acc[i][j][0]);
            f(r, c + 1, acc[i][j][1]);
        }
}

// diagonal block k: Cholesky + inverse of the factor in shared memory.  grid = molecules, 256 threads,
// dynamic shared memory kBcholDiagSmem
__global__ void __launch_bounds__(256) k_bchol_diag(Dev d, int k) {
    const int im = blockIdx.x;
    const MolDesc m = d.mol[im];
    const int n = m.nao, c0 = CB * k;
    if (c0 >= n) return;
    const int b = min(CB, n - c0);
    extern __shared__ double smem[];   // 2 x 64 x 65 doubles (above the 48 KB static limit)
    double (*A)[CB + 1] = reinterpret_cast<double (*)[CB + 1]>(smem);
    double (*Li)[CB + 1] = reinterpret_cast<double (*)[CB + 1]>(smem + CB * (CB + 1));
    __shared__ int s_bad;
    double* W = d.W + m.mat0;
    const int tid = threadIdx.x, nt = blockDim.x;
    if (tid == 0) s_bad = 0;
    for (int t = tid; t < CB * CB; t += nt) {
        const int r = t % CB, c = t / CB;
        A[r][c] = (r < b && c < b) ? W[(c0 + r) + (long long)n * (c0 + c)] : (r == c ? 1.0 : 0.0);
        Li[r][c] = 0.0;
    }
    __syncthreads();
    for (int j = 0; j < b; ++j) {
        double ajj = A[j][j];
        if (!(ajj > 0.0)) { if (tid == 0) s_bad = 1; ajj = 1.0; }   // not positive definite: flagged below
        const double ljj = sqrt(ajj), rl = 1.0 / ljj;
        __syncthreads();
        for (int r = j + tid; r < b; r += nt) A[r][j] = (r == j) ? ljj : A[r][j] * rl;
        __syncthreads();
        const int mm = b - j - 1;
        for (int t = tid; t < mm * mm; t += nt) {
            const int r = j + 1 + t % mm, c = j + 1 + t / mm;
            if (r >= c) A[r][c] -= A[r][j] * A[c][j];
        }
        __syncthreads();
    }
    // inverse of the lower-triangular factor, one column per thread
    if (tid < b) {
        const int c = tid;
        Li[c][c] = 1.0 / A[c][c];
        for (int r = c + 1; r < b; ++r) {
            double s = 0.0;
            for (int kk = c; kk < r; ++kk) s += A[r][kk] * Li[kk][c];
            Li[r][c] = -s / A[r][r];
        }
    }
    __syncthreads();
    double* Dk = d.cholD + (long long)im * d.chol_stride + (long long)k * CB * CB;
    for (int t = tid; t < CB * CB; t += nt) {
        const int r = t % CB, c = t / CB;
        if (r < b && c < b) W[(c0 + r) + (long long)n * (c0 + c)] = (r >= c) ? A[r][c] : 0.0;
        Dk[t] = (r < b && c < b) ? Li[r][c] : 0.0;   // column-major 64 x 64, zero padded
    }
    if (tid == 0 && s_bad) d.st[im].fail = 11;
}

// panel below the diagonal block: L_ik = A_ik D_k^T.  grid = (row blocks below k, molecules), 128 threads
__global__ void __launch_bounds__(128) k_bchol_panel(Dev d, int k) {
    const MolDesc m = d.mol[blockIdx.y];
    const int n = m.nao, c0 = CB * k, r0 = CB * (k + 1 + (int)blockIdx.x);
    if (r0 >= n || c0 >= n) return;
    const int b = min(CB, n - c0), mr = min(CB, n - r0);
    __shared__ double As[GK][GLD], Bs[GK][GLD];
    double* W = d.W + m.mat0;
    const double* Dk = d.cholD + (long long)blockIdx.y * d.chol_stride + (long long)k * CB * CB;
    double acc[4][4][2];
    bchol_zero(acc);
    // op(B)(kk, c) = D_k(c, kk) = Dk[c + 64 kk]
    gemm64_acc<false, true>(acc, W + r0 + (long long)n * c0, n, Dk, CB, mr, b, b, As, Bs);
    bchol_foreach(acc, [&](int r, int c, double v) { if (r < mr && c < b) W[(r0 + r) + (long long)n * (c0 + c)] = v; });
}

// trailing update: A_ij -= L_ik L_jk^T for the lower-triangle tiles i >= j > k.
// grid = (tiles of the trailing triangle, molecules), 128 threads
__global__ void __launch_bounds__(128) k_bchol_update(Dev d, int k) {
    const MolDesc m = d.mol[blockIdx.y];
    const int n = m.nao, c0 = CB * k;
    const int nb = (n + CB - 1) / CB, mt = nb - k - 1;
    const long long t = blockIdx.x;
    if (mt <= 0 || t >= (long long)mt * (mt + 1) / 2) return;
    int ti = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while ((long long)ti * (ti + 1) / 2 > t) --ti;
    while ((long long)(ti + 1) * (ti + 2) / 2 <= t) ++ti;
    const int tj = (int)(t - (long long)ti * (ti + 1) / 2);
    const int r0 = CB * (k + 1 + ti), q0 = CB * (k + 1 + tj);
    const int b = min(CB, n - c0), mr = min(CB, n - r0), mc = min(CB, n - q0);
    __shared__ double As[GK][GLD], Bs[GK][GLD];
    double* W = d.W + m.mat0;
    double acc[4][4][2];
    bchol_zero(acc);
    // op(B)(kk, c) = L_jk(c, kk) = W[(q0 + c) + n (c0 + kk)]
    gemm64_acc<false, true>(acc, W + r0 + (long long)n * c0, n, W + q0 + (long long)n * c0, n, mr, mc, b, As, Bs);
    bchol_foreach(acc, [&](int r, int c, double v) { if (r < mr && c < mc) W[(r0 + r) + (long long)n * (q0 + c)] -= v; });
}

// Lt_ik = D_i L_ik for every tile below the diagonal (in place).  grid = (tiles i > k, molecules), 128 threads
__global__ void __launch_bounds__(128) k_bchol_scale(Dev d) {
    const MolDesc m = d.mol[blockIdx.y];
    const int n = m.nao, nb = (n + CB - 1) / CB;
    const long long t = blockIdx.x;
    if (t >= (long long)nb * (nb - 1) / 2) return;
    int ti = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while ((long long)ti * (ti + 1) / 2 > t) --ti;
    while ((long long)(ti + 1) * (ti + 2) / 2 <= t) ++ti;
    const int tk = (int)(t - (long long)ti * (ti + 1) / 2);
    const int i = ti + 1, r0 = CB * i, c0 = CB * tk;
    const int mr = min(CB, n - r0);
    __shared__ double As[GK][GLD], Bs[GK][GLD];
    double* W = d.W + m.mat0;
    const double* Di = d.cholD + (long long)blockIdx.y * d.chol_stride + (long long)i * CB * CB;
    double acc[4][4][2];
    bchol_zero(acc);
    gemm64_acc<false, false>(acc, Di, CB, W + r0 + (long long)n * c0, n, mr, CB, mr, As, Bs);
    bchol_foreach(acc, [&](int r, int c, double v) { if (r < mr) W[(r0 + r) + (long long)n * (c0 + c)] = v; });
}

// diagonal blocks of X = Linv^T: X_jj = D_j^T.  grid = (block columns, molecules), 256 threads (X zeroed before)
__global__ void k_bchol_inv_diag(Dev d) {
    const MolDesc m = d.mol[blockIdx.y];
    const int n = m.nao, j = blockIdx.x, c0 = CB * j;
    if (c0 >= n) return;
    const int b = min(CB, n - c0);
    const double* Dj = d.cholD + (long long)blockIdx.y * d.chol_stride + (long long)j * CB * CB;
    double* X = d.X + m.mat0;
    for (int t = threadIdx.x; t < CB * CB; t += blockDim.x) {
        const int r = t % CB, c = t / CB;
        if (r < b && c < b) X[(c0 + c) + (long long)n * (c0 + r)] = Dj[r + CB * c];   // X(c0+c, c0+r) = D_j(r, c)
    }
}

// block row i of Linv: Linv_ij = - sum_{k=j}^{i-1} Lt_ik Linv_kj, stored as X_ji = Linv_ij^T.
// grid = (block columns j < i, molecules), 128 threads
__global__ void __launch_bounds__(128) k_bchol_inv_row(Dev d, int i) {
    const MolDesc m = d.mol[blockIdx.y];
    const int n = m.nao, j = blockIdx.x, r0 = CB * i, c0 = CB * j;
    if (j >= i || r0 >= n) return;
    const int mr = min(CB, n - r0);
    __shared__ double As[GK][GLD], Bs[GK][GLD];
    const double* W = d.W + m.mat0;
    double* X = d.X + m.mat0;
    double acc[4][4][2];
    bchol_zero(acc);
    // A(r, kk) = Lt(r0 + r, c0 + kk);  op(B)(kk, c) = Linv(c0 + kk, c0 + c) = X[(c0 + c) + n (c0 + kk)]
    gemm64_acc<false, true>(acc, W + r0 + (long long)n * c0, n, X + c0 + (long long)n * c0, n, mr, CB, r0 - c0, As, Bs);
    bchol_foreach(acc, [&](int r, int c, double v) { if (r < mr) X[(c0 + c) + (long long)n * (r0 + r)] = -v; });
}

}  // namespace xtbb200
