// Symmetric eigensolver for matrices that do not fit into shared memory (crambin, peptides, water clusters):
// blocked Householder tridiagonalisation, divide and conquer on the tridiagonal matrix, blocked back-transformation.
//
//   A = Q T Q^T      k_sytrd_panel (persistent cooperative kernel, one launch per panel of EB columns: column update,
//                    reflector, w = tau (A v - V W^T v - W V^T v) with the trailing-matrix product as warp-per-column
//                    dot products -- the HBM-bound part, 8 (n-j)^2 bytes per column) + rank-2*EB trailing update on the
//                    FP64 tensor path (k_eg_gemm, DMMA m8n8k4); once the trailing matrix fits into the shared memory
//                    of all SMs together (order <= ~1920) k_sytrd_smem finishes the reduction: columns resident in
//                    shared memory, unblocked rank-2 updates, one grid barrier per column
//   T = Qt L Qt^T    Cuppen divide and conquer down to 1x1 / 2x2 leaves, one batch of kernels per tree level:
//                    rank sort, deflation scan, Givens rotations, secular equation (warp per root), Loewner update
//                    vector, eigenvector update as two DMMA GEMMs per node (top / bottom rows, columns grouped by
//                    kind as in LAPACK's dlaed2/3), batched over the level (scalar routines: eig_dc_core.h)
//   Z = Q Qt         block reflectors I - V T V^T over two panels; T factors from the Gram matrices V^T V (the V^T v
//                    products of the panel kernel, one batched GEMM for the resident range); three DMMA GEMMs per block
//
// The same stages batched over an ensemble of small matrices of one order (k_sytrd_cta, blockIdx.z = matrix) are kept
// as a cross-check of the block Jacobi solver (solver mode 4).
//
// Replaces dsyevd('V') inside TEigenSolver%fact_solve (src/mctc/lapack/eigensolve.F90:218-246) for large n; the
// reduction to standard form (dsygst) and the back-substitution (dtrsm) are the X^T H X and X Z products of engine.cu.
#pragma once
#include <cuda.h>   // CUtensorMap (types only: the encoder is looked up through the runtime, no link against libcuda)

#include "common.cuh"
#include "eig_dc_core.h"
#include "kernels_linalg.cuh"
#include "sytrd_sym_grid.h"

namespace xtbb200 {

constexpr int EB = 64;             // panel width of the tridiagonalisation
constexpr int NB = 2 * EB;         // block reflector width of the back-transformation (two panels)
constexpr int kSytrdThreads = 1024;
constexpr int kBtSplit = 4;        // largest split of the long K dimension of W = V^T Z (small matrices: too few tiles otherwise)

// ---------------------------------------------------------------------------------------------------------------
// general GEMM  C = alpha op(A) op(B) + beta C  on 64x64 tiles (128 threads, 4 warps of 32x32, DMMA m8n8k4)
// ---------------------------------------------------------------------------------------------------------------
struct GemmTask {
    const double* A;
    const double* B;
    double* C;
    int lda, ldb, ldc;
    int M, N, K;
    double alpha, beta;
    // two-segment K range: k >= ksegA reads K-index (k - ksegA + offA2) of A (ksegA = 0: plain); same for B.  Lets
    // [V | W] [W | V]^T run as one product without copying the panels.
    int ksegA, offA2, ksegB, offB2;
    int ksplit;           // > 1: blockIdx.z selects a K chunk, partial products go to C + blockIdx.z * csplit
    long long csplit;
    int lower;            // != 0: only the tiles on and below the diagonal of a square C are computed
};
inline GemmTask gemm_task(const double* A, int lda, const double* B, int ldb, double* C, int ldc, int M, int N, int K,
                          double alpha, double beta) {
    GemmTask t;
    t.A = A; t.B = B; t.C = C; t.lda = lda; t.ldb = ldb; t.ldc = ldc; t.M = M; t.N = N; t.K = K; t.alpha = alpha; t.beta = beta;
    t.ksegA = 0; t.offA2 = 0; t.ksegB = 0; t.offB2 = 0; t.ksplit = 1; t.csplit = 0; t.lower = 0;
    return t;
}

__device__ __forceinline__ GemmTask gemm_task_dev(const double* A, int lda, const double* B, int ldb, double* C, int ldc, int M,
                                                  int N, int K) {
    GemmTask t;
    t.A = A; t.B = B; t.C = C; t.lda = lda; t.ldb = ldb; t.ldc = ldc; t.M = M; t.N = N; t.K = K; t.alpha = 1.0; t.beta = 0.0;
    t.ksegA = 0; t.offA2 = 0; t.ksegB = 0; t.offB2 = 0; t.ksplit = 1; t.csplit = 0; t.lower = 0;
    return t;
}

// C = alpha acc + beta C for a warp's 32 x 32 accumulator block (row gr0 + 8 i, columns gc0 + 8 j + {0, 1}).  With beta != 0
// the old values are fetched sixteen at a time before the first store, so that their latencies overlap (a load after a
// store to the same array cannot be moved up by the compiler: one L2 / HBM round trip per element otherwise).
__device__ __forceinline__ void gemm_epilogue(const double (&acc)[4][4][2], double* __restrict__ C, int ldc, int M, int N, int gr0, int gc0,
                                              double alpha, double beta) {
    if (beta == 0.0) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int gr = gr0 + 8 * i, gc = gc0 + 8 * j;
                if (gr < M) {
#pragma unroll
                    for (int h = 0; h < 2; ++h)
                        if (gc + h < N) C[gr + (long long)ldc * (gc + h)] = alpha * acc[i][j][h];
                }
            }
        return;
    }
#pragma unroll
    for (int ih = 0; ih < 4; ih += 2) {
        double old[2][4][2];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int gr = gr0 + 8 * (ih + i), gc = gc0 + 8 * j + h;
                    old[i][j][h] = (gr < M && gc < N) ? C[gr + (long long)ldc * gc] : 0.0;
                }
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int gr = gr0 + 8 * (ih + i), gc = gc0 + 8 * j + h;
                    if (gr < M && gc < N) C[gr + (long long)ldc * gc] = alpha * acc[ih + i][j][h] + beta * old[i][j][h];
                }
    }
}

// tasks != null: task = tasks[blockIdx.y + zstride * blockIdx.z] (zstride > 0: batch of matrices along z, no K split)
template <bool TA, bool TB>
__global__ void __launch_bounds__(128, 3) k_eg_gemm(GemmTask t0, const GemmTask* __restrict__ tasks, int zstride = 0,
                                                  const int* __restrict__ skip = nullptr) {
    if (skip && skip[blockIdx.z]) return;
    const GemmTask t = tasks ? tasks[blockIdx.y + (long long)zstride * blockIdx.z] : t0;
    const int tm = (t.M + GT - 1) / GT, tn = (t.N + GT - 1) / GT;
    if ((int)blockIdx.x >= tm * tn) return;
    const int r0 = (blockIdx.x % tm) * GT, c0 = (blockIdx.x / tm) * GT;
    if (t.lower && c0 > r0) return;
    int kb = 0, ke = t.K;
    double* __restrict__ C = t.C + r0 + (long long)t.ldc * c0;
    if (t.ksplit > 1) {
        const int ch = (((t.K + t.ksplit - 1) / t.ksplit + GK - 1) / GK) * GK;
        kb = min(t.K, (int)blockIdx.z * ch);
        ke = min(t.K, kb + ch);
        C += (long long)blockIdx.z * t.csplit;
    }
    // tile-local operands: A rows / B columns start at the tile, K index 0 = first K index of the task
    const double* __restrict__ A = TA ? t.A + (long long)t.lda * r0 : t.A + r0;
    const double* __restrict__ B = TB ? t.B + c0 : t.B + (long long)t.ldb * c0;
    const int lda = t.lda, ldb = t.ldb, ldc = t.ldc;
    const int M = t.M - r0, N = t.N - c0;   // valid rows / columns of this tile (may exceed GT)
    const int sgA = t.ksegA ? t.ksegA : 0x7fffffff, oA = t.offA2 - t.ksegA;
    const int sgB = t.ksegB ? t.ksegB : 0x7fffffff, oB = t.offB2 - t.ksegB;
    // operand tiles: [k][row] with pitch GLD when the k index is the slow one in global memory, [row][k] with pitch GK + 4 when it
    // is the fast one (TA / !TB): in both layouts the staging stores and the m8n8k4 fragment loads are free of bank conflicts
    constexpr int PK = GK + 4;
    __shared__ double As[TA ? GT * PK : GK * GLD], Bs[TB ? GK * GLD : GT * PK];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wr = (warp & 1) * 32, wc = (warp >> 1) * 32;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    double ra[GT * GK / 128], rb[GT * GK / 128];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int u = 0; u < GT * GK / 128; ++u) {
            const int tt = tid + 128 * u;
            if (!TA) {
                const int r = tt % GT, gk = k0 + tt / GT, ka = gk >= sgA ? gk + oA : gk;
                ra[u] = (r < M && gk < ke) ? A[r + (long long)lda * ka] : 0.0;
            } else {
                const int gk = k0 + tt % GK, r = tt / GK, ka = gk >= sgA ? gk + oA : gk;
                ra[u] = (r < M && gk < ke) ? A[ka + (long long)lda * r] : 0.0;
            }
            if (!TB) {
                const int gk = k0 + tt % GK, c = tt / GK, kq = gk >= sgB ? gk + oB : gk;
                rb[u] = (c < N && gk < ke) ? B[kq + (long long)ldb * c] : 0.0;
            } else {
                const int c = tt % GT, gk = k0 + tt / GT, kq = gk >= sgB ? gk + oB : gk;
                rb[u] = (c < N && gk < ke) ? B[c + (long long)ldb * kq] : 0.0;
            }
        }
    };
    auto stage = [&]() {
#pragma unroll
        for (int u = 0; u < GT * GK / 128; ++u) {
            const int tt = tid + 128 * u;
            if (!TA) As[(tt / GT) * GLD + tt % GT] = ra[u]; else As[(tt / GK) * PK + tt % GK] = ra[u];
            if (!TB) Bs[(tt / GK) * PK + tt % GK] = rb[u]; else Bs[(tt / GT) * GLD + tt % GT] = rb[u];
        }
    };
    if (ke > kb) fetch(kb);
    for (int k0 = kb; k0 < ke; k0 += GK) {
        stage();
        __syncthreads();
        if (k0 + GK < ke) fetch(k0 + GK);
#pragma unroll
        for (int ks = 0; ks < GK; ks += 4) {
            double af[4], bf[4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
                af[i] = TA ? As[(wr + 8 * i + (lane >> 2)) * PK + ks + (lane & 3)] : As[(ks + (lane & 3)) * GLD + wr + 8 * i + (lane >> 2)];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                bf[j] = TB ? Bs[(ks + (lane & 3)) * GLD + wc + 8 * j + (lane >> 2)] : Bs[(wc + 8 * j + (lane >> 2)) * PK + ks + (lane & 3)];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma_884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        __syncthreads();
    }
    gemm_epilogue(acc, C, ldc, M, N, wr + (lane >> 2), wc + 2 * (lane & 3), t.alpha, t.beta);
}

// ---------------------------------------------------------------------------------------------------------------
// The same GEMM with the operand tiles brought in by the bulk asynchronous copy engine (TMA, cp.async.bulk: SASS UBLKCP)
// through a ring of kTmaStages shared-memory stages guarded by mbarriers: one producer warp issues the copies of k-tile
// i + kTmaStages - 1 while the four consumer warps run the DMMAs of k-tile i; no register staging, no __syncthreads in
// the main loop.  Every copy is one line of the tile that is contiguous in global memory (64 rows of a column: 512 B, or
// 16 k-values of a row: 128 B) and lands at a padded pitch (68 / 20 doubles), so the fragment loads stay free of bank
// conflicts; the hardware swizzles of the tensor-map form of TMA do not fit the 8-byte m8n8k4 fragments.
// Requirements (checked by EighWork::gemm, which otherwise launches k_eg_gemm): base pointers 16-byte aligned, lda / ldb
// even, M and N even, K and the segment boundaries multiples of GK.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kTmaStages = 4;
constexpr int kTmaPitchK = GK + 4;      // lines of 16 k-values
template <bool TA, bool TB>
constexpr size_t eg_tma_smem() {
    return (size_t)kTmaStages * ((TA ? GT * kTmaPitchK : GK * GLD) + (TB ? GK * GLD : GT * kTmaPitchK)) * sizeof(double) +
           2 * kTmaStages * sizeof(unsigned long long);
}
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(b)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// one box of a 2-D tensor map (dimension 0 = rows of the column-major matrix, dimension 1 = columns) into shared memory:
// SASS UTMALDG.2D; lines outside the matrix arrive as zeros and count towards the transaction bytes like the others
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int row0, int col0, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_u32(dst)),
                 "l"(reinterpret_cast<unsigned long long>(map)), "r"(row0), "r"(col0), "r"(smem_u32(bar))
                 : "memory");
}

template <bool TA, bool TB>
__global__ void __launch_bounds__(160) k_eg_gemm_tma(GemmTask t) {
    const int tm = (t.M + GT - 1) / GT, tn = (t.N + GT - 1) / GT;
    if ((int)blockIdx.x >= tm * tn) return;
    const int r0 = (blockIdx.x % tm) * GT, c0 = (blockIdx.x / tm) * GT;
    if (t.lower && c0 > r0) return;
    int kb = 0, ke = t.K;
    double* __restrict__ C = t.C + r0 + (long long)t.ldc * c0;
    if (t.ksplit > 1) {
        const int ch = (((t.K + t.ksplit - 1) / t.ksplit + GK - 1) / GK) * GK;
        kb = min(t.K, (int)blockIdx.z * ch);
        ke = min(t.K, kb + ch);
        C += (long long)blockIdx.z * t.csplit;
    }
    const double* __restrict__ A = TA ? t.A + (long long)t.lda * r0 : t.A + r0;
    const double* __restrict__ B = TB ? t.B + c0 : t.B + (long long)t.ldb * c0;
    const int lda = t.lda, ldb = t.ldb, ldc = t.ldc;
    const int M = t.M - r0, N = t.N - c0;
    const int rows = min(M, GT), cols = min(N, GT);
    const int sgA = t.ksegA ? t.ksegA : 0x7fffffff, oA = t.offA2 - t.ksegA;
    const int sgB = t.ksegB ? t.ksegB : 0x7fffffff, oB = t.offB2 - t.ksegB;
    constexpr int ASZ = TA ? GT * kTmaPitchK : GK * GLD, BSZ = TB ? GK * GLD : GT * kTmaPitchK;
    extern __shared__ __align__(128) unsigned char eg_smem[];
    double* Abuf = reinterpret_cast<double*>(eg_smem);
    double* Bbuf = Abuf + kTmaStages * ASZ;
    unsigned long long* full = reinterpret_cast<unsigned long long*>(Bbuf + kTmaStages * BSZ);
    unsigned long long* empty = full + kTmaStages;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < kTmaStages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int nk = (ke - kb) / GK;
    if (warp == 4) {
        // ---- producer warp
        const unsigned bytes = (unsigned)(rows + cols) * GK * sizeof(double);
        for (int it = 0; it < nk; ++it) {
            const int s = it % kTmaStages, k0 = kb + it * GK;
            if (it >= kTmaStages) mbar_wait(empty + s, ((it / kTmaStages) - 1) & 1);
            if (lane == 0) mbar_expect_tx(full + s, bytes);
            __syncwarp();
            const int ka = k0 >= sgA ? k0 + oA : k0, kq = k0 >= sgB ? k0 + oB : k0;
            double* As = Abuf + s * ASZ;
            double* Bs = Bbuf + s * BSZ;
            if (!TA) { if (lane < GK) bulk_g2s(As + lane * GLD, A + (long long)lda * (ka + lane), rows * 8, full + s); }
            else { for (int r = lane; r < rows; r += 32) bulk_g2s(As + r * kTmaPitchK, A + ka + (long long)lda * r, GK * 8, full + s); }
            if (TB) { if (lane < GK) bulk_g2s(Bs + lane * GLD, B + (long long)ldb * (kq + lane), cols * 8, full + s); }
            else { for (int c = lane; c < cols; c += 32) bulk_g2s(Bs + c * kTmaPitchK, B + kq + (long long)ldb * c, GK * 8, full + s); }
        }
        return;
    }
    // ---- consumer warps
    const int wr = (warp & 1) * 32, wc = (warp >> 1) * 32;
    const int l4 = lane >> 2, l3 = lane & 3;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int it = 0; it < nk; ++it) {
        const int s = it % kTmaStages;
        mbar_wait(full + s, (it / kTmaStages) & 1);
        const double* As = Abuf + s * ASZ;
        const double* Bs = Bbuf + s * BSZ;
#pragma unroll
        for (int ks = 0; ks < GK; ks += 4) {
            double af[4], bf[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) af[i] = TA ? As[(wr + 8 * i + l4) * kTmaPitchK + ks + l3] : As[(ks + l3) * GLD + wr + 8 * i + l4];
#pragma unroll
            for (int j = 0; j < 4; ++j) bf[j] = TB ? Bs[(ks + l3) * GLD + wc + 8 * j + l4] : Bs[(wc + 8 * j + l4) * kTmaPitchK + ks + l3];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma_884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + s);
    }
    gemm_epilogue(acc, C, ldc, M, N, wr + l4, wc + 2 * l3, t.alpha, t.beta);
}

// ---------------------------------------------------------------------------------------------------------------
// tridiagonalisation
// ---------------------------------------------------------------------------------------------------------------
struct SytrdArgs {
    double* A;          // n x n, column-major, both triangles valid on entry of every panel
    int n, j0, ncols;   // panel columns [j0, j0 + ncols)
    double* V;          // n x EB reflectors of the panel (explicit, zero above the unit entry)
    double* W;          // n x EB
    const double* Vprev;// reflectors of the previous panel when it belongs to the same block reflector, else null
    double* d;          // diagonal of T
    double* e;          // sub-diagonal of T
    double* tau;
    double* y;          // n: trailing-matrix product A v
    double* part_n;     // [grid] partial sums of squares
    double* part_vy;    // [grid] partial v^T y
    double* vtv;        // [EB]  V^T v
    double* wtv;        // [EB]  W^T v
    double* G;          // NB x NB Gram matrix of the block reflector: G[k + NB i] = v_k^T v_i (k < i), input of the T factor
    int gq;             // offset of this panel inside the block reflector (0 or EB)
    unsigned* bar;      // grid barrier counter (zeroed before the launch)
    long long* prof;    // optional: cycles of CTA 0 per phase section (diagnostics), else null
    // symmetric form of the trailing-matrix product (sym != 0): only the lower triangle is read, as 128 x 32 tiles streamed
    // through a shared-memory ring by the bulk asynchronous copy engine; partial sums of y per tile / per CTA segment
    int sym;            // 0: warp-per-column dot products over both triangles
    int nst;            // stages of the tile ring (2 .. kSyStagesMax)
    double* colbuf;     // [tile][kSyCW]  column sums of every tile (the transposed product)
    double* rowbuf;     // [(row chunk - first row chunk) + CTA][kSyRH]  row sums of a CTA's consecutive tiles of one row chunk
};
#define SY_T(k) do { if (a.prof && cta == 0 && tid == 0) { const long long t_ = clock64(); a.prof[k] += t_ - tprof; tprof = t_; } } while (0)

// barrier over all CTAs of a cooperative grid; bar is zeroed before the launch, target is the caller's running count
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned nblocks, unsigned& target) {
    // arrive with a release reduction (cumulative over the CTA's writes ordered before it by the CTA barrier), wait with
    // acquire loads: one L2 round trip less than fence + atomic + fence
    __syncthreads();
    if (threadIdx.x == 0) {
        target += nblocks;
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
        unsigned x;
        do {
            asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(x) : "l"(bar) : "memory");
        } while (x < target);
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------------
// EEQ guess of one large molecule (more than 320 atoms): the Gaussian elimination with partial pivoting of k_eeq_pivot /
// k_eeq_elim (kernels_geom.cuh; replaces dsysv, src/eeq_model.f90:858-862) as ONE cooperative launch with two grid barriers
// per column instead of two launches per column (6000 launches for 3000 atoms).  Same pivot rule, same arithmetic per
// element (f = A_ik / A_kk as A_ik * (1 / A_kk); A_ij -= f A_kj), hence the same charges bit for bit.  Every CTA finds the
// pivot itself (the column is read by all); the row swap and the elimination are spread over the grid, a warp per column.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kEeqRowMax = 256;   // columns of a CTA's share: (natoms + 2) / grid + 1 must fit
__global__ void __launch_bounds__(1024, 1) k_eeq_lu_coop(Dev d, unsigned* bar) {
    const MolDesc m = d.mol[0];
    const int n1 = m.nat + 1, nc = n1 + 1;
    double* A = d.eeqA + d.eeq0[0];
    const int G = gridDim.x, cta = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    __shared__ int s_piv;
    __shared__ double s_red[32];
    __shared__ int s_idx[32];
    __shared__ double s_row[kEeqRowMax];
    unsigned target = 0;
    for (int k = 0; k < n1; ++k) {
        double best = -1.0; int bi = k;
        for (int i = k + tid; i < n1; i += blockDim.x) {
            const double v = fabs(__ldcg(A + i + (long long)n1 * k));
            if (v > best) { best = v; bi = i; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (lane == 0) { s_red[tid >> 5] = best; s_idx[tid >> 5] = bi; }
        __syncthreads();
        if (tid == 0) {
            const int nw = (blockDim.x + 31) >> 5;
            double b = s_red[0]; int ix = s_idx[0];
            for (int w = 1; w < nw; ++w) if (s_red[w] > b || (s_red[w] == b && s_idx[w] < ix)) { b = s_red[w]; ix = s_idx[w]; }
            s_piv = ix;
            if (b == 0.0 && cta == 0) d.st[0].fail = 10;
        }
        __syncthreads();
        const int p = s_piv;
        if (p != k)
            for (int j = k + cta * (int)blockDim.x + tid; j < nc; j += G * (int)blockDim.x) {
                const double t = __ldcg(A + k + (long long)n1 * j);
                A[k + (long long)n1 * j] = __ldcg(A + p + (long long)n1 * j);
                A[p + (long long)n1 * j] = t;
            }
        grid_barrier(bar, G, target);
        // elimination: this CTA's share of the trailing columns, threads along the rows (multiplier in a register, the columns'
        // updates independent: many loads in flight); the pivot row segment is staged in shared memory
        const double pinv = 1.0 / __ldcg(A + k + (long long)n1 * k);
        const int ntc = nc - k - 1;
        const int jb = k + 1 + (int)(((long long)ntc * cta) / G), je = k + 1 + (int)(((long long)ntc * (cta + 1)) / G);
        for (int j = jb + tid; j < je; j += blockDim.x) s_row[j - jb] = __ldcg(A + k + (long long)n1 * j);
        __syncthreads();
        for (int i = k + 1 + tid; i < n1; i += blockDim.x) {
            const double f = __ldcg(A + i + (long long)n1 * k) * pinv;
            for (int j = jb; j < je; ++j) {
                double* q = A + i + (long long)n1 * j;
                *q = __ldcg(q) - f * s_row[j - jb];
            }
        }
        grid_barrier(bar, G, target);
    }
}

// deterministic block sum, result in all threads; second stage by shuffles in every warp (red: 32 doubles)
__device__ __forceinline__ double block_sum_sh(double v, double* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    return warp_sum(lane < nw ? red[lane] : 0.0);
}

// Block-wide sum of v into part[blockIdx.x] followed by a barrier over all CTAs of the (cooperative) grid.  red: 32 doubles.
__device__ __forceinline__ void grid_sum_barrier(double v, double* red, double* part, unsigned* bar, unsigned nblocks, unsigned& target) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = warp_sum(v);
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (warp == 0) {
        double s = lane < nw ? red[lane] : 0.0;
        s = warp_sum(s);
        if (lane == 0) {
            part[blockIdx.x] = s;
            target += nblocks;
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
            unsigned x;
            do {
                asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(x) : "l"(bar) : "memory");
            } while (x < target);
        }
    }
    __syncthreads();
}


constexpr int kSyMaxChunks = 96;                    // row chunks of the largest order the symmetric form takes
constexpr int kSyGT = kSytrdThreads / kSyGroups;   // threads of a group of warps of the symmetric panel product
static_assert(kSyRH == 4 * kSyCW && kSyRH == 4 * 32 && kSyGT / 32 * 4 == kSyCW , "tile mapping of k_sytrd_panel");

constexpr int kSyRows = 64;                         // rows per CTA pass of the panel update
constexpr int kSySlices = kSytrdThreads / kSyRows;  // threads per row, each with a slice of the panel columns
constexpr int kSyKmax = (EB + kSySlices - 1) / kSySlices;

// dynamic shared memory: sv[n] | sh_part[(n / grid + 6) * 8]; symmetric form: sv[npad] (absolute row index, zero outside the
// trailing range) | sh_part | tile ring [nst][kSyTile] (128-byte aligned)
inline size_t sytrd_sym_ring_offset(int n, int grid) {   // in doubles
    const size_t npad = (size_t)(n + kSyRH - 1) / kSyRH * kSyRH;
    return (npad + (size_t)(n / grid + 6) * 8 + 15) / 16 * 16;
}
__global__ void __launch_bounds__(kSytrdThreads, 1) k_sytrd_panel(SytrdArgs a, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(128) double sm_dyn[];
    __shared__ double sh_vtv[EB], sh_wtv[EB], sh_vrow[EB], sh_wrow[EB], red[32], sh_sc[4];
    __shared__ __align__(16) double sh_r12[2][kSySlices][kSyRows];
    __shared__ double sh_r3[kSySlices][kSyRows];
    double (*sh_r1)[kSyRows] = sh_r12[0];
    double (*sh_r2)[kSyRows] = sh_r12[1];
    __shared__ unsigned long long sy_full[kSyGroups];
    // tables of the tile grid of the product in flight, for the gather of its sums in the next column update: tiles before
    // every row chunk; first owner CTA and number of owner CTAs of three row chunks (the two this CTA's rows of the next
    // update lie in, and the one of row j + 1)
    __shared__ int sy_pre[kSyMaxChunks], sy_plo[3], sy_nseg[3];
    const int n = a.n, G = gridDim.x, cta = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
    const int rl = tid & (kSyRows - 1), ks = tid / kSyRows;
    const long long ld = n;
    double* __restrict__ A = a.A;
    double* __restrict__ V = a.V;
    double* __restrict__ W = a.W;
    const bool sym = a.sym != 0;
    const int npad = (n + kSyRH - 1) / kSyRH * kSyRH;
    double* svabs = sm_dyn;                                   // symmetric form: v by absolute row
    double* sv = sm_dyn;                                      // v by row relative to j + 1 (set per column in the symmetric form)
    double* sh_part = sm_dyn + (sym ? npad : n);
    double* ring = sm_dyn + (sym ? ((size_t)npad + (size_t)(n / G + 6) * 8 + 15) / 16 * 16 : 0);
    // symmetric form: the CTA works as kSyGroups groups of 8 warps; group q takes every kSyGroups-th tile of the CTA's range
    // into ring stage q (one tensor-map box per tile, issued by the group's first thread, completion on sy_full[q])
    const int grp = tid / kSyGT, gw = warp & (kSyGT / 32 - 1);
    unsigned guse = 0;                                        // tiles this group has taken since the launch (mbarrier phase)
    if (sym) {
        for (int t = tid; t < npad; t += nthr) svabs[t] = 0.0;
        if (tid == 0) {
            for (int s = 0; s < kSyGroups; ++s) mbar_init(sy_full + s, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }
    auto issue_tile = [&](int R, int S) {                     // one thread of the group
        mbar_expect_tx(sy_full + grp, (unsigned)(kSyTile * sizeof(double)));
        tma_load_2d(ring + (size_t)grp * kSyTile, &tmap, R * kSyRH, S * kSyCW, sy_full + grp);
    };
    unsigned target = 0;
    double taup = 0.0;   // tau of the previous column
    long long tprof = clock64();
    for (int i = 0;; ++i) {
        const int j = a.j0 + i;
        const bool fin = i > 0, upd = i < a.ncols;
        const int ip = i - 1, jp = j - 1;
        const SymGrid gy = (sym && fin) ? sym_grid(j, n) : SymGrid();   // tile grid of the product with column j-1's reflector
        // ---------------- phase A: finish w of column j-1, update column j, partial norm.  A CTA pass covers kSyRows
        // rows; the panel columns of a row are shared out over kSySlices threads so that all loads of a pass are in
        // flight together (the panel sits in L2; a serial loop per row would pay its latency 2*ip times).
        double sq = 0.0;
        for (int base = j + cta * kSyRows, pass = 0; base < n || pass == 0; base += G * kSyRows, ++pass) {
            const int r = base + rl;
            const bool rv = r < n;
            double vk[kSyKmax], wk[kSyKmax], vr = 0.0, yr = 0.0, ar = 0.0, yg = 0.0, yjg = 0.0;
            if (sym && fin) {   // y = A v of column j-1 from the tile sums: this thread's slice of the terms of row r (and of row j)
                auto gather = [&](int row, int e, int slice) {   // e: which of the three tabulated row chunks
                    const int R = row / kSyRH, S = row / kSyCW, plo = sy_plo[e], nseg = sy_nseg[e], nit = nseg + gy.nR - R;
                    const double* rowp = a.rowbuf + (size_t)(R - gy.Rmin + plo) * kSyRH + (row - R * kSyRH);
                    const double* colp = a.colbuf + (size_t)(S - gy.Smin) * kSyCW + (row - S * kSyCW);
                    double x[kSyGatherTerms], sum = 0.0;
#pragma unroll
                    for (int u = 0; u < kSyGatherTerms; ++u) {
                        const int t = slice + u * kSySlices;
                        x[u] = t < nseg ? __ldcg(rowp + (size_t)t * kSyRH)
                                        : (t < nit ? __ldcg(colp + (size_t)sy_pre[R + (t - nseg) - gy.Rmin] * kSyCW) : 0.0);
                    }
#pragma unroll
                    for (int u = 0; u < kSyGatherTerms; ++u) sum += x[u];
                    return sum;
                };
                if (rv) yg = gather(r, r / kSyRH - (j + cta * kSyRows) / kSyRH, ks);
                if (warp == 0 && lane < kSySlices) yjg = gather(j, 2, lane);
            }
            if (fin) {
#pragma unroll
                for (int u = 0; u < kSyKmax; ++u) {
                    const int k = ks + u * kSySlices;
                    const bool ok = rv && k < ip;
                    vk[u] = ok ? __ldcg(V + r + ld * k) : 0.0;
                    wk[u] = ok ? __ldcg(W + r + ld * k) : 0.0;
                }
                if (ks == 0 && rv) { vr = __ldcg(V + r + ld * ip); if (!sym) yr = __ldcg(a.y + r); }
            }
            if (ks == 0 && rv && upd) ar = __ldcg(A + r + ld * j);
            SY_T(8);
            double alpha2 = 0.0;
            if (fin) {
                if (pass == 0) {
                    double vy = 0.0, yj = 0.0;
                    if (warp == 0) {
                        for (int c = lane; c < G; c += 32) vy += __ldcg(a.part_vy + c);
                        if (sym) {
#pragma unroll
                            for (int u = 0; u < kSySlices; ++u) yj += __shfl_sync(0xffffffffu, yjg, u);
                        } else if (lane == 0) yj = __ldcg(a.y + j);
                    }
                    for (int k = tid; k < ip; k += nthr) {
                        sh_vtv[k] = __ldcg(a.vtv + k); sh_wtv[k] = __ldcg(a.wtv + k);
                        sh_vrow[k] = __ldcg(V + j + ld * k); sh_wrow[k] = __ldcg(W + j + ld * k);
                    }
                    __syncthreads();
                    SY_T(9);
                    if (warp == 0) {
                        double vw = 0.0, s = 0.0;
                        for (int k = lane; k < ip; k += 32) { vw += sh_vtv[k] * sh_wtv[k]; s += sh_vrow[k] * sh_wtv[k] + sh_wrow[k] * sh_vtv[k]; }
                        vy = warp_sum(vy); vw = warp_sum(vw); s = warp_sum(s);
                        if (lane == 0) {
                            const double al2 = -0.5 * taup * taup * (vy - 2.0 * vw);
                            sh_sc[0] = al2;
                            sh_vrow[ip] = 1.0;                          // V[j, ip]: the unit entry of v
                            sh_wrow[ip] = taup * (yj - s) + al2;         // W[j, ip]
                        }
                    }
                    __syncthreads();
                    SY_T(0);
                }
                alpha2 = sh_sc[0];
                double s1 = 0.0, s2 = 0.0;
#pragma unroll
                for (int u = 0; u < kSyKmax; ++u) {
                    const int k = ks + u * kSySlices;
                    if (k < ip) {
                        s1 += vk[u] * sh_wtv[k] + wk[u] * sh_vtv[k];
                        s2 += vk[u] * sh_wrow[k] + wk[u] * sh_vrow[k];
                    }
                }
                sh_r1[ks][rl] = s1; sh_r2[ks][rl] = s2;
                if (sym) sh_r3[ks][rl] = yg;
                __syncthreads();
            }
            if (ks == 0 && rv) {
                double s2 = 0.0;
                if (fin) {
                    double s1 = 0.0;
#pragma unroll
                    for (int u = 0; u < kSySlices; ++u) { s1 += sh_r1[u][rl]; s2 += sh_r2[u][rl]; }
                    if (sym) {
#pragma unroll
                        for (int u = 0; u < kSySlices; ++u) yr += sh_r3[u][rl];
                    }
                    const double wr = taup * (yr - s1) + alpha2 * vr;
                    W[r + ld * ip] = wr;
                    if (r > j) A[r + ld * jp] = vr;             // reflector storage below the sub-diagonal
                    s2 += vr * sh_wrow[ip] + wr * sh_vrow[ip];
                }
                if (upd) {
                    const double av = ar - s2;
                    A[r + ld * j] = av;
                    if (r == j) a.d[j] = av;
                    else if (r >= j + 2) sq += av * av;
                }
            }
            if (fin && base + G * kSyRows < n) __syncthreads();   // sh_r1/sh_r2 are reused by the next pass
        }
        if (!upd || j == n - 1) break;
        SY_T(1);
        // symmetric form: this CTA's tiles of the product with column j's reflector; the trailing matrix does not change
        // inside a panel, so the first copies are issued before the barrier and land while it is being crossed
        const SymGrid sg = sym ? sym_grid(j + 1, n) : SymGrid();
        int tcur = 0, tend = 0, cR = 0, cS = 0;   // tiles [tcur, tend) of this CTA, the first one is (cR, cS)
        if (sym) {
            if (tid < sg.nR - sg.Rmin) sy_pre[tid] = sym_pre(sg, sg.Rmin + tid);
            if (tid >= 96 && tid < 99) {
                const int e = tid - 96, R = e == 2 ? sg.Rmin : (j + 1 + cta * kSyRows) / kSyRH + e;
                if (R < sg.nR) {
                    const int tb = sym_pre(sg, R), plo = sym_owner(sg, tb, G);
                    sy_plo[e] = plo; sy_nseg[e] = sym_owner(sg, tb + sym_cnt(sg, R) - 1, G) - plo + 1;
                }
            }
            tcur = sym_start(sg, cta, G); tend = sym_start(sg, cta + 1, G);
            if (tcur < tend) sym_locate(sg, tcur, cR, cS);
            if ((tid & (kSyGT - 1)) == 0 && tcur + grp < tend) {
                int R = cR, S = cS;
                for (int u = 0; u < grp; ++u) sym_next(sg, R, S);
                issue_tile(R, S);
            }
        }
        grid_sum_barrier(sq, red, a.part_n, a.bar, G, target);
        SY_T(3);
        // ---------------- phase B: reflector, trailing-matrix product, V^T v, W^T v
        const int m = n - j - 1;
        if (sym) {
            sv = svabs + j + 1;
            if (tid == 0) svabs[j] = 0.0;   // the unit entry of the previous reflector
        }
        {   // raw column and the norm partials are loaded together; the column is scaled once the reflector is known
            double s = 0.0, alpha = 0.0;
            if (warp == 0) {
                for (int c = lane; c < G; c += 32) s += __ldcg(a.part_n + c);
                if (lane == 0) alpha = __ldcg(A + (j + 1) + ld * j);
            }
            for (int t = tid; t < m; t += nthr) sv[t] = __ldcg(A + (j + 1 + t) + ld * j);
            if (warp == 0) {
                s = warp_sum(s);
                if (lane == 0) {
                    double beta = alpha, tau = 0.0, scal = 0.0;
                    if (s > 0.0) {
                        beta = -copysign(sqrt(alpha * alpha + s), alpha);
                        tau = (beta - alpha) / beta;
                        scal = 1.0 / (alpha - beta);
                    }
                    sh_sc[1] = tau; sh_sc[2] = scal;
                    if (cta == 0) { a.e[j] = beta; a.tau[j] = tau; }
                }
            }
        }
        __syncthreads();
        const double tau = sh_sc[1], scal = sh_sc[2];
        for (int t = tid; t < m; t += nthr) sv[t] = (t == 0) ? 1.0 : sv[t] * scal;
        __syncthreads();
        for (long long t = (long long)cta * nthr + tid; t < m; t += (long long)G * nthr) V[(j + 1 + t) + ld * i] = sv[t];
        SY_T(4);
        // jobs of this CTA: its share of the trailing columns, plus V / W (/ previous panel V) columns dealt round-robin
        const int cb = j + 1 + (int)(((long long)m * cta) / G), ce = j + 1 + (int)(((long long)m * (cta + 1)) / G);
        const int ncA = sym ? 0 : ce - cb;
        // two-triangle form: V and W column k = cta, cta + G, ... as a pair; symmetric form (no trailing columns here, the CTAs
        // are otherwise idle): the 2 i columns singly, f = cta, cta + G, ... (f even: V column f / 2, odd: W), up to 32 warps each
        int nex = 0, nex2 = 0, nvw = 0;
        if (sym) { for (int f = cta; f < 2 * i; f += G) ++nvw; }
        else {
            for (int k = cta; k < i; k += G) ++nex;
            nvw = 2 * nex;
            if (a.Vprev) for (int k = cta; k < EB; k += G) ++nex2;       // symmetric form: Gram matrices by one GEMM afterwards
        }
        const int nloc = ncA + nvw + nex2;
        int Q = 1;
        if (nloc > 0) {
            Q = (4 * nwarp + nloc - 1) / nloc;
            Q = max(1, min(sym ? 32 : 8, min(Q, m / 128)));
            if (sym) Q = min(Q, max(1, nwarp / nloc));
        }
        const int njobs = nloc * Q;
        for (int job = warp; job < njobs; job += nwarp) {
            const int lc = job / Q, q = job - lc * Q;
            const double* col;
            if (lc < ncA) col = A + (j + 1) + ld * (cb + lc);
            else if (lc < ncA + nvw) {
                const int x = lc - ncA, f = sym ? cta + x * G : 0, k = sym ? f >> 1 : cta + (x >> 1) * G;
                col = (((sym ? f : x) & 1) ? W : V) + (j + 1) + ld * k;
            } else col = a.Vprev + (j + 1) + ld * (cta + (lc - ncA - nvw) * G);
            const int tb = (int)(((long long)m * q) / Q), te = (int)(((long long)m * (q + 1)) / Q);
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            int t = tb + lane;
            for (; t + 224 < te; t += 256) {   // eight loads in flight per lane: the L2 / HBM latency is the limit here
                const double a0 = __ldcg(col + t), a1 = __ldcg(col + t + 32), a2 = __ldcg(col + t + 64), a3 = __ldcg(col + t + 96);
                const double a4 = __ldcg(col + t + 128), a5 = __ldcg(col + t + 160), a6 = __ldcg(col + t + 192), a7 = __ldcg(col + t + 224);
                s0 += a0 * sv[t]; s1 += a1 * sv[t + 32]; s2 += a2 * sv[t + 64]; s3 += a3 * sv[t + 96];
                s0 += a4 * sv[t + 128]; s1 += a5 * sv[t + 160]; s2 += a6 * sv[t + 192]; s3 += a7 * sv[t + 224];
            }
            for (; t + 96 < te; t += 128) {
                s0 += __ldcg(col + t) * sv[t];
                s1 += __ldcg(col + t + 32) * sv[t + 32];
                s2 += __ldcg(col + t + 64) * sv[t + 64];
                s3 += __ldcg(col + t + 96) * sv[t + 96];
            }
            for (; t < te; t += 32) s0 += __ldcg(col + t) * sv[t];
            const double s = warp_sum((s0 + s1) + (s2 + s3));
            if (lane == 0) sh_part[job] = s;
        }
        __syncthreads();
        SY_T(5);
        double vyl = 0.0;
        for (int lc = tid; lc < nloc; lc += nthr) {
            double s = 0.0;
            for (int q = 0; q < Q; ++q) s += sh_part[lc * Q + q];
            if (lc < ncA) {
                const int c = cb + lc;
                a.y[c] = s;
                vyl += sv[c - (j + 1)] * s;
            } else if (lc < ncA + nvw) {
                const int x = lc - ncA, f = sym ? cta + x * G : 0, k = sym ? f >> 1 : cta + (x >> 1) * G;
                if ((sym ? f : x) & 1) a.wtv[k] = s;
                else { a.vtv[k] = s; if (!sym) a.G[(a.gq + k) + NB * (a.gq + i)] = s; }
            } else a.G[(cta + (lc - ncA - nvw) * G) + NB * (a.gq + i)] = s;
        }
        taup = tau;
        SY_T(6);
        if (sym) {
            // ---- lower-triangle tiles: y[r] += T[r, c] v[c] (row sums) and y[c] += T[r, c] v[r] (column sums); diagonal tiles
            // count A[r, r] once and nothing above it.  Inside a group: lane = 2 + 2 rows (half a tile apart), warp = 4 consecutive columns
            // of the tile; every element is read once.  The row sums stay in registers over the tiles of a row chunk and are
            // added up over the CTA when the chunk ends; the column sums of a tile are complete inside the warp.
            double rowacc[4] = {0.0, 0.0, 0.0, 0.0};
            double* scr = &sh_r12[0][0][0];                              // [16 warps][kSyRH]
            auto flush_rows = [&](int R) {
                double tot = 0.0;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if ((warp >> 4) == h) {
                        double* q = scr + (warp & 15) * kSyRH + 2 * lane;
                        *reinterpret_cast<double2*>(q) = make_double2(rowacc[0], rowacc[1]);
                        *reinterpret_cast<double2*>(q + kSyRH / 2) = make_double2(rowacc[2], rowacc[3]);
                    }
                    __syncthreads();
                    if (tid < kSyRH) {
#pragma unroll
                        for (int u = 0; u < 16; ++u) tot += scr[u * kSyRH + tid];
                    }
                    __syncthreads();
                }
                if (tid < kSyRH) {
                    a.rowbuf[(size_t)(R - sg.Rmin + cta) * kSyRH + tid] = tot;
                    vyl += svabs[R * kSyRH + tid] * tot;
                }
                rowacc[0] = rowacc[1] = rowacc[2] = rowacc[3] = 0.0;
            };
            const int ntile = tend - tcur;
            int kR = cR, kS = cS;
            for (int k = 0; k < ntile; ++k) {
                if ((k & (kSyGroups - 1)) == grp) {
                    mbar_wait(sy_full + grp, guse & 1);
                    ++guse;
                    const double* T = ring + (size_t)grp * kSyTile;
                    // rows r0, r0 + 1, r0 + 64, r0 + 65 (16-byte loads of consecutive lanes are contiguous: no bank conflicts)
                    const int r0 = kR * kSyRH + 2 * lane, c0 = kS * kSyCW + 4 * gw;   // first row / column of this thread
                    const bool diag = (kS >> 2) == kR;
                    double vr[4], vc[4], colp[4];
                    {
                        const double2 p0 = *reinterpret_cast<const double2*>(svabs + r0), p1 = *reinterpret_cast<const double2*>(svabs + r0 + kSyRH / 2);
                        const double2 q0 = *reinterpret_cast<const double2*>(svabs + c0), q1 = *reinterpret_cast<const double2*>(svabs + c0 + 2);
                        vr[0] = p0.x; vr[1] = p0.y; vr[2] = p1.x; vr[3] = p1.y;
                        vc[0] = q0.x; vc[1] = q0.y; vc[2] = q1.x; vc[3] = q1.y;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const double* q = T + (4 * gw + u) * kSyRH + 2 * lane;
                        const double2 xa = *reinterpret_cast<const double2*>(q), xb = *reinterpret_cast<const double2*>(q + kSyRH / 2);
                        double x[4] = {xa.x, xa.y, xb.x, xb.y};
                        if (diag) {
                            double cp = 0.0;
#pragma unroll
                            for (int h = 0; h < 4; ++h) {
                                const int d = (r0 + (h & 1) + (h >> 1) * (kSyRH / 2)) - (c0 + u);
                                rowacc[h] += (d >= 0 ? x[h] : 0.0) * vc[u];
                                cp += (d > 0 ? x[h] : 0.0) * vr[h];
                            }
                            colp[u] = cp;
                        } else {
#pragma unroll
                            for (int h = 0; h < 4; ++h) rowacc[h] += x[h] * vc[u];
                            colp[u] = (x[0] * vr[0] + x[1] * vr[1]) + (x[2] * vr[2] + x[3] * vr[3]);
                        }
                    }
                    // four column sums over the 32 lanes: halve the number of values per lane twice, then three plain steps
                    double cs;
                    {
                        const bool h4 = lane & 16, h3 = lane & 8;
                        const double b0 = (h4 ? colp[2] : colp[0]) + __shfl_xor_sync(0xffffffffu, h4 ? colp[0] : colp[2], 16);
                        const double b1 = (h4 ? colp[3] : colp[1]) + __shfl_xor_sync(0xffffffffu, h4 ? colp[1] : colp[3], 16);
                        cs = (h3 ? b1 : b0) + __shfl_xor_sync(0xffffffffu, h3 ? b0 : b1, 8);
                        cs += __shfl_xor_sync(0xffffffffu, cs, 4);
                        cs += __shfl_xor_sync(0xffffffffu, cs, 2);
                        cs += __shfl_xor_sync(0xffffffffu, cs, 1);
                    }
                    if ((lane & 7) == 0) {                                // lanes 0, 8, 16, 24: columns 0, 1, 2, 3 of the warp
                        const int c = c0 + (lane >> 3);
                        if (c < n) {
                            a.colbuf[(size_t)(tcur + k) * kSyCW + 4 * gw + (lane >> 3)] = cs;
                            vyl += svabs[c] * cs;
                        }
                    }
                    asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(kSyGT) : "memory");   // the group has read its stage
                    if ((tid & (kSyGT - 1)) == 0 && k + kSyGroups < ntile) {
                        int R = kR, S = kS;
#pragma unroll
                        for (int u = 0; u < kSyGroups; ++u) sym_next(sg, R, S);
                        issue_tile(R, S);
                    }
                }
                const int Rdone = kR;
                sym_next(sg, kR, kS);
                if (kR != Rdone || k + 1 == ntile) flush_rows(Rdone);
            }
        }
        grid_sum_barrier(vyl, red, a.part_vy, a.bar, G, target);
        SY_T(7);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// tridiagonalisation with the trailing matrix resident in shared memory (order up to ~1860 on 148 SMs)
// ---------------------------------------------------------------------------------------------------------------
// The columns of the (trailing) matrix are dealt cyclically to the CTAs of a cooperative grid and stay in their shared
// memory for the whole reduction, so the unblocked two-sided update A -= v w^T + w v^T (dsytd2) costs no HBM or L2
// traffic at all.  One grid barrier per column: before it every CTA publishes its entries of y = A v and the owner of
// the next column publishes that column as it stands; after it every CTA forms w, updates its columns and forms the
// NEXT reflector itself from the published column (redundantly, O(m) work), so no second exchange is needed.  The
// panel kernel above is bound by L2 round trips below this order (~12 us per column); this form needs ~6.
struct SytrdSmemArgs {
    double* A;          // n x n; the reduction covers A[j0:, j0:]; reflectors are stored below the sub-diagonal
    int n, j0;          // j0 must be a multiple of NB
    double *d, *e, *tau;
    double* cbuf;       // [2][n] next column as published by its owner (double buffered by the parity of the column)
    double* ybuf;       // [2][n] y = A v
    double* Gf;         // Gram matrices of the block reflectors (as written by the panel kernel)
    unsigned* bar;
    // cluster form (k_sytrd_smem<RR, true>): a batch of matrices of one order, one thread-block cluster of `clu` CTAs per
    // matrix (blockIdx.x / clu); matrices ms apart, d / e / tau vs apart, cbuf / ybuf 2 n apart; skip[matrix] != 0: left out
    int clu;
    long long ms, vs;
    const int* skip;
    int lower_only;     // != 0: only the lower triangle of A[j0:, j0:] is valid on entry
};
// barrier over the CTAs of a thread-block cluster (hardware barrier; release / acquire at cluster scope also orders the
// global-memory exchange buffers, which are read back with ld.cg)
__device__ __forceinline__ void cluster_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
constexpr int kResGridMin = 148;   // smallest grid of the resident kernel (see EighWork::solve): fewer CTAs were slower at
                                   // every order tried (n = 552: 3.09 ms with 148 CTAs, 3.71 with 32, 4.84 with 12)
constexpr int kResRows = 2;   // rows of a column per thread: the resident kernel takes orders up to kResRows * kSytrdThreads
inline size_t sytrd_smem_resident_bytes(int m0, int grid) {
    const size_t slots = (size_t)(m0 + grid - 1) / grid;
    return (slots * m0 + 2 * (size_t)m0 + 128) * sizeof(double);
}

template <int RR, bool CLU = false>   // RR = rows of a column per thread (the CTA has at least m0 / RR threads)
__global__ void __launch_bounds__(kSytrdThreads, 1) k_sytrd_smem(SytrdSmemArgs a) {
    extern __shared__ double sm_dyn[];
    __shared__ double red[32], sh_d, sh_alpha;
    int G = gridDim.x, cta = blockIdx.x;
    if (CLU) {   // one cluster per matrix of the batch: the cluster's CTAs take the role of the cooperative grid
        G = a.clu;
        cta = blockIdx.x % G;
        const long long ib = blockIdx.x / G;
        if (a.skip && a.skip[ib]) return;   // the same for every CTA of the cluster
        a.A += ib * a.ms; a.d += ib * a.vs; a.e += ib * a.vs; a.tau += ib * a.vs;
        a.cbuf += ib * 2 * a.n; a.ybuf += ib * 2 * a.n;
    }
    const int n = a.n, j0 = a.j0, m0 = n - j0, tid = threadIdx.x, nthr = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
    const long long ld = n;
    const int maxslot = (m0 + G - 1) / G;
    const int nslot = cta < m0 ? (m0 - cta + G - 1) / G : 0;   // owned columns c = cta + s G
    double* cols = sm_dyn;
    double* sv = sm_dyn + (size_t)maxslot * m0;
    double* sw = sv + m0;
    double* part = sw + m0;                                    // 128 doubles
    for (int s = 0; s < nslot; ++s) {
        const int c = cta + s * G;
        const double* src = a.A + j0 + ld * (j0 + c);
        if (a.lower_only)   // the panel kernel's updates kept only the lower triangle current: rows above the diagonal from the row
            for (int r = tid; r < m0; r += nthr) cols[(size_t)s * m0 + r] = r >= c ? src[r] : a.A[(j0 + c) + ld * (j0 + r)];
        else
            for (int r = tid; r < m0; r += nthr) cols[(size_t)s * m0 + r] = src[r];
    }
    __syncthreads();
    unsigned target = 0;
    double tau = 0.0;
    // Reflector of column jn from its up-to-date entries cn[k] (row jn + tid + k nthr) held in registers by every CTA:
    // v -> sv, tau; CTA jn % G stores d, e, tau and the reflector.  Ends with a CTA barrier.
    auto form_reflector = [&](int jn, const double (&cn)[RR]) {
        double sq = 0.0;
#pragma unroll
        for (int k = 0; k < RR; ++k) {
            const int r = jn + tid + k * nthr;
            if (r == jn) sh_d = cn[k];
            else if (r == jn + 1) sh_alpha = cn[k];
            else if (r < m0) sq += cn[k] * cn[k];
        }
        sq = block_sum_sh(sq, red);
        const bool writer = cta == jn % G;
        if (jn == m0 - 1) {
            if (writer && tid == 0) a.d[j0 + jn] = sh_d;
            tau = 0.0;
            return;
        }
        const double alpha = sh_alpha;
        double beta = alpha, scal = 0.0;
        tau = 0.0;
        if (sq > 0.0) {
            beta = -copysign(sqrt(alpha * alpha + sq), alpha);
            tau = (beta - alpha) / beta;
            scal = 1.0 / (alpha - beta);
        }
        if (writer && tid == 0) { a.d[j0 + jn] = sh_d; a.e[j0 + jn] = beta; a.tau[j0 + jn] = tau; }
#pragma unroll
        for (int k = 0; k < RR; ++k) {
            const int r = jn + tid + k * nthr;
            if (r > jn && r < m0) {
                const double v = (r == jn + 1) ? 1.0 : cn[k] * scal;
                sv[r] = v;
                if (writer && r >= jn + 2) a.A[(j0 + r) + ld * (j0 + jn)] = v;
            }
        }
        __syncthreads();
    };
    {   // first column: published by its owner, reflector formed by everybody
        if (cta == 0) for (int r = tid; r < m0; r += nthr) a.cbuf[ld + r] = cols[r];
        if (CLU) cluster_barrier(); else grid_barrier(a.bar, G, target);
        double cn[RR];
#pragma unroll
        for (int k = 0; k < RR; ++k) { const int r = tid + k * nthr; cn[k] = r < m0 ? __ldcg(a.cbuf + ld + r) : 0.0; }
        form_reflector(0, cn);
    }
    for (int jj = 0; jj < m0 - 1; ++jj) {
        const int j = j0 + jj, mm = m0 - jj - 1;
        double* ybuf = a.ybuf + (jj & 1) * ld;
        double* cbuf = a.cbuf + (jj & 1) * ld;
        // ---- y = A v on the owned columns (the Gram matrices of the block reflectors follow from one batched GEMM afterwards)
        const int s0 = (jj + 1 - cta + G - 1) / G;            // first owned slot with column > jj (cta + s G >= jj + 1)
        const int sfirst = cta >= jj + 1 ? 0 : s0;
        const int nact = nslot - sfirst;
        const int njobcols = max(nact, 0);
        int Q = 1;
        if (njobcols > 0) Q = max(1, min(8, min((nwarp + njobcols - 1) / njobcols, max(mm / 64, 1))));
        const int njobs = njobcols * Q;
        for (int job = warp; job < njobs; job += nwarp) {
            const int lc = job / Q, q = job - lc * Q;
            const int tb = jj + 1 + (int)(((long long)mm * q) / Q), te = jj + 1 + (int)(((long long)mm * (q + 1)) / Q);
            double s = 0.0;
            {
                const double* col = cols + (size_t)(sfirst + lc) * m0;
                double s1 = 0.0;
                int r = tb + lane;
                for (; r + 32 < te; r += 64) { s += col[r] * sv[r]; s1 += col[r + 32] * sv[r + 32]; }
                if (r < te) s += col[r] * sv[r];
                s += s1;
            }
            s = warp_sum(s);
            if (lane == 0) part[job] = s;
        }
        if ((jj + 1) % G == cta) {                            // the next column as it stands (its update comes after the barrier)
            const double* col = cols + (size_t)((jj + 1) / G) * m0;
            for (int r = jj + 1 + tid; r < m0; r += nthr) cbuf[r] = col[r];
        }
        __syncthreads();
        for (int lc = tid; lc < njobcols; lc += nthr) {
            double s = 0.0;
            for (int q = 0; q < Q; ++q) s += part[lc * Q + q];
            ybuf[cta + (sfirst + lc) * G] = s;
        }
        if (CLU) cluster_barrier(); else grid_barrier(a.bar, G, target);
        // ---- w = tau y - tau^2/2 (v^T y) v;  A -= v w^T + w v^T on the owned columns and on the published next column
        double vy = 0.0, cn[RR];
#pragma unroll
        for (int k = 0; k < RR; ++k) {
            const int r = jj + 1 + tid + k * nthr;
            cn[k] = 0.0;
            if (r < m0) { const double y = __ldcg(ybuf + r); cn[k] = __ldcg(cbuf + r); sw[r] = y; vy += sv[r] * y; }
        }
        vy = block_sum_sh(vy, red);
        const double coef = -0.5 * tau * tau * vy;
#pragma unroll
        for (int k = 0; k < RR; ++k) {
            const int r = jj + 1 + tid + k * nthr;
            if (r < m0) sw[r] = tau * sw[r] + coef * sv[r];
        }
        __syncthreads();
        {   // thread = (row, slot parity): rows fastest so that the shared-memory accesses are conflict free
            double vr[RR], wr[RR];
#pragma unroll
            for (int k = 0; k < RR; ++k) {
                const int r = jj + 1 + tid + k * nthr;
                vr[k] = r < m0 ? sv[r] : 0.0;
                wr[k] = r < m0 ? sw[r] : 0.0;
            }
            for (int s = 0; s < nact; ++s) {
                const int c = cta + (sfirst + s) * G;
                const double wc = sw[c], vc = sv[c];
                double* col = cols + (size_t)(sfirst + s) * m0;
#pragma unroll
                for (int k = 0; k < RR; ++k) {
                    const int r = jj + 1 + tid + k * nthr;
                    if (r < m0) col[r] -= vr[k] * wc + wr[k] * vc;
                }
            }
        }
        const double wn = sw[jj + 1], vn = sv[jj + 1];
#pragma unroll
        for (int k = 0; k < RR; ++k) {
            const int r = jj + 1 + tid + k * nthr;
            if (r < m0) cn[k] -= sv[r] * wn + sw[r] * vn;
        }
        form_reflector(jj + 1, cn);
    }
}

// Gram matrices G_b = V_b^T V_b of the block reflectors whose columns were reduced by the resident kernel: explicit V of all
// those blocks side by side (Vall: n x (n - jres), column jb + k - jres), then one task per block for k_eg_gemm<true,false>
__global__ void k_gram_formv(const double* __restrict__ A, int n, int jres, double* __restrict__ Vall) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int nc = n - 1 - jres;                              // reflectors jres .. n-2
    if (t >= (long long)n * nc) return;
    const int r = (int)(t % n), c = jres + (int)(t / n), piv = c + 1;
    Vall[t] = r < piv ? 0.0 : (r == piv ? 1.0 : A[r + (long long)n * c]);
}
__global__ void k_gram_tasks(GemmTask* __restrict__ tasks, int n, int jres, double* Vall, double* Gf) {
    const int b = jres / NB + blockIdx.x * blockDim.x + threadIdx.x, jb = b * NB;
    if (jb >= n - 1) return;
    const int kb = min(NB, n - 1 - jb), r0 = jb + 1;
    double* V = Vall + (long long)n * (jb - jres) + r0;
    tasks[b - jres / NB] = gemm_task_dev(V, n, V, n, Gf + (long long)b * NB * NB, NB, kb, kb, n - r0);
}

// ---------------------------------------------------------------------------------------------------------------
// back-transformation  Z <- (I - V T V^T) Z  per block reflector of NB columns
// ---------------------------------------------------------------------------------------------------------------
// explicit V of the block [jb, jb+kb): Vb[r + n k], rows absolute
__global__ void k_bt_formv(const double* __restrict__ A, int n, int jb, int kb, double* __restrict__ Vb) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)n * kb) return;
    const int r = (int)(t % n), k = (int)(t / n), piv = jb + k + 1;
    Vb[t] = r < piv ? 0.0 : (r == piv ? 1.0 : A[r + (long long)n * (jb + k)]);
}
// T factors of all blocks (one CTA per block): T(0:i,i) = -tau_i T(0:i,0:i) G(0:i,i), T(i,i) = tau_i
// Row r of T depends on nothing but itself and G; kTfParts threads share a row and split the k range of its dot products
// (the column-by-column recurrence is serial in i: 128 steps of <= 128 / kTfParts terms instead of <= 128).
constexpr int kTfParts = 8;
__global__ void __launch_bounds__(NB * kTfParts) k_bt_tfactor(const double* __restrict__ Gall, const double* __restrict__ tau, int n,
                                                              double* __restrict__ Tall) {
    extern __shared__ double sm_dyn[];
    double (*T)[NB + 1] = (double (*)[NB + 1])sm_dyn;                   // NB x (NB + 1)
    double (*part)[NB] = (double (*)[NB])(sm_dyn + NB * (NB + 1));     // kTfParts x NB
    const int b = blockIdx.x, jb = b * NB, kb = min(NB, n - 1 - jb), r = threadIdx.x % NB, pt = threadIdx.x / NB;
    const double* G = Gall + (long long)b * NB * NB;
    for (int c = pt; c < NB; c += kTfParts) T[r][c] = 0.0;
    __syncthreads();
    for (int i = 0; i < kb; ++i) {
        double s = 0.0;
        if (r < i)
            for (int k = r + pt; k < i; k += kTfParts) s += T[r][k] * G[k + NB * i];
        part[pt][r] = s;
        __syncthreads();
        if (pt == 0) {
            if (r < i) {
                double t = 0.0;
#pragma unroll
                for (int q = 0; q < kTfParts; ++q) t += part[q][r];
                T[r][i] = -tau[jb + i] * t;
            } else if (r == i) T[r][i] = tau[jb + i];
        }
        __syncthreads();
    }
    double* To = Tall + (long long)b * NB * NB;
    for (int c = pt; c < NB; c += kTfParts) To[r + NB * c] = T[r][c];
}
// W = sum_z Wz (partial products of the K-split W = V^T Z); thread per element
__global__ void k_bt_sum(double* __restrict__ Wz, long long zstride, int nz, long long count) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    double w = Wz[t];
    for (int z = 1; z < nz; ++z) w += Wz[(long long)z * zstride + t];
    Wz[t] = w;
}

// ---------------------------------------------------------------------------------------------------------------
// divide and conquer on the tridiagonal matrix; blockIdx.y (or .z) = node of the current tree level
// ---------------------------------------------------------------------------------------------------------------
struct DcArgs {
    int n, depth;
    const double* d;      // diagonal of T
    const double* e;      // sub-diagonal of T (coupling of the torn blocks)
    // batch of matrices of the same order along blockIdx.z (0 strides: single matrix): vectors are vs apart, matrices
    // ms, skip[matrix] != 0 leaves a matrix out
    long long vs, ms;
    const int* skip;
    double *Qo, *Qn;      // eigenvector blocks of the children / of the merged nodes
    double *U, *Qg, *Ug;  // delta matrix; non-deflated columns and update matrix in grouped order
    double *lam, *lam2, *z, *zhat, *dk, *zk, *rc, *rs;
    int *perm, *cols, *pos, *tcnt, *ctype, *ra, *rb, *kcnt, *nrot, *fail;
    GemmTask* tasks;
};
// select the matrix of the batch this CTA works on; false: skipped
__device__ __forceinline__ bool dc_select(DcArgs& a) {
    const long long ib = blockIdx.z;
    if (a.skip && a.skip[ib]) return false;
    const long long v = ib * a.vs, m = ib * a.ms;
    a.d += v; a.e += v; a.Qo += m; a.Qn += m; a.U += m; a.Qg += m; a.Ug += m;
    a.lam += v; a.lam2 += v; a.z += v; a.zhat += v; a.dk += v; a.zk += v; a.rc += v; a.rs += v;
    a.perm += v; a.cols += v; a.pos += v; a.tcnt += 3 * v; a.ctype += v; a.ra += v; a.rb += v; a.kcnt += v; a.nrot += v;
    a.tasks += 2 * v;
    if (a.vs) a.fail += ib;
    return true;
}
struct DcNode { int lo, mid, hi; };
__device__ __forceinline__ DcNode dc_node(const DcArgs& a, int t) {
    DcNode nd;
    nd.lo = eigcore::tree_bound(a.n, a.depth, t);
    nd.hi = eigcore::tree_bound(a.n, a.depth, t + 1);
    nd.mid = eigcore::tree_bound(a.n, a.depth + 1, 2 * t + 1);
    return nd;
}
struct WarpPar {
    __device__ __forceinline__ int first() const { return threadIdx.x & 31; }
    __device__ __forceinline__ int step() const { return 32; }
    __device__ __forceinline__ double sum(double x) const { return warp_sum(x); }
};

// leaves of the tearing tree: 1x1 and 2x2 blocks of the torn matrix; thread per leaf (a.depth = depth of the leaves)
__global__ void k_dc_leaves(DcArgs a) {
    if (!dc_select(a)) return;
    const int n = a.n, D = a.depth;
    const double* __restrict__ d = a.d;
    const double* __restrict__ e = a.e;
    double* __restrict__ lam = a.lam;
    double* __restrict__ Q = a.Qo;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (1 << D)) return;
    const int lo = eigcore::tree_bound(n, D, t), hi = eigcore::tree_bound(n, D, t + 1);
    const long long ld = n;
    double av = d[lo] - (lo > 0 ? fabs(e[lo - 1]) : 0.0);
    if (hi - lo == 1) {
        av -= (hi < n ? fabs(e[hi - 1]) : 0.0);
        lam[lo] = av;
        Q[lo + ld * lo] = 1.0;
    } else {
        const double c = d[lo + 1] - (hi < n ? fabs(e[hi - 1]) : 0.0), b = e[lo];
        double rt1, rt2, cs, sn;
        eigcore::sym2x2(av, b, c, rt1, rt2, cs, sn);
        lam[lo] = rt1; lam[lo + 1] = rt2;
        Q[lo + ld * lo] = cs; Q[lo + 1 + ld * lo] = sn;
        Q[lo + ld * (lo + 1)] = -sn; Q[lo + 1 + ld * (lo + 1)] = cs;
    }
}

// update vector z (last row of the left block, first row of the right block) and ascending order of the eigenvalues
__global__ void __launch_bounds__(256) k_dc_sort(DcArgs a) {
    if (!dc_select(a)) return;
    extern __shared__ double sm_dyn[];
    const DcNode nd = dc_node(a, blockIdx.y);
    const int m = nd.hi - nd.lo;
    if ((int)(blockIdx.x * blockDim.x) >= m) return;
    for (int i = threadIdx.x; i < m; i += blockDim.x) sm_dyn[i] = a.lam[nd.lo + i];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const long long ld = a.n;
    const int c = nd.lo + i;
    const double sg = a.e[nd.mid - 1] < 0.0 ? -1.0 : 1.0;
    a.z[c] = (c < nd.mid ? a.Qo[(nd.mid - 1) + ld * c] : sg * a.Qo[nd.mid + ld * c]) * 0.70710678118654752440;
    const double v = sm_dyn[i];
    int r = 0;
    for (int k = 0; k < m; ++k) { const double w = sm_dyn[k]; r += (w < v || (w == v && k < i)) ? 1 : 0; }
    a.perm[nd.lo + r] = i;
}

// deflation: one CTA per node, sequential scan by thread 0 over eigenvalues / z / order / column kinds staged in shared
// memory (use_smem) or, for nodes too large for it, in place in global memory
__global__ void __launch_bounds__(256) k_dc_deflate(DcArgs a, int use_smem) {
    if (!dc_select(a)) return;
    extern __shared__ double sm_dyn[];
    const DcNode nd = dc_node(a, blockIdx.x);
    const int m = nd.hi - nd.lo, lo = nd.lo;
    double* sd = use_smem ? sm_dyn : a.lam + lo;
    double* sz = use_smem ? sm_dyn + m : a.z + lo;
    int* sp = use_smem ? (int*)(sm_dyn + 2 * m) : a.perm + lo;
    int* sty = use_smem ? sp + m : a.ctype + lo;
    __shared__ int sh_k;
    if (use_smem) {
        for (int i = threadIdx.x; i < m; i += blockDim.x) { sd[i] = a.lam[lo + i]; sz[i] = a.z[lo + i]; sp[i] = a.perm[lo + i]; }
    }
    __syncthreads();
    const double rho = 2.0 * fabs(a.e[nd.mid - 1]);
    if (threadIdx.x == 0) {
        int nrot = 0;
        const int k = eigcore::deflate_scan(m, nd.mid - lo, sp, sd, sz, rho, a.cols + lo, sty, a.ra + lo, a.rb + lo, a.rc + lo,
                                            a.rs + lo, &nrot);
        eigcore::group_columns(k, a.cols + lo, sty, a.pos + lo, a.tcnt + 3 * blockIdx.x);
        sh_k = k;
        a.kcnt[blockIdx.x] = k;
        a.nrot[blockIdx.x] = nrot;
    }
    __syncthreads();
    const int k = sh_k;
    for (int p = threadIdx.x; p < m; p += blockDim.x) {
        const int c = a.cols[lo + p];
        if (p < k) { a.dk[lo + p] = sd[c]; a.zk[lo + p] = sz[c]; }
        else a.lam2[lo + p] = sd[c];
    }
}

// Givens rotations of the deflation applied to the eigenvector columns; thread per row
__global__ void __launch_bounds__(256) k_dc_rot(DcArgs a) {
    if (!dc_select(a)) return;
    const int nr = a.nrot[blockIdx.y];
    if (nr == 0) return;
    const DcNode nd = dc_node(a, blockIdx.y);
    const int row = nd.lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= nd.hi) return;
    const long long ld = a.n;
    for (int r = 0; r < nr; ++r) {
        double* px = a.Qo + row + ld * (nd.lo + a.ra[nd.lo + r]);
        double* py = a.Qo + row + ld * (nd.lo + a.rb[nd.lo + r]);
        const double c = a.rc[nd.lo + r], s = a.rs[nd.lo + r], x = *px, y = *py;
        *px = c * x + s * y;
        *py = c * y - s * x;
    }
}

// secular equation: warp per root (the lanes share the poles); delta(i,j) = d_i - lambda_j into U
__global__ void __launch_bounds__(256) k_dc_secular(DcArgs a) {
    if (!dc_select(a)) return;
    extern __shared__ double sm_dyn[];
    const int k = a.kcnt[blockIdx.y];
    if ((int)(blockIdx.x * (blockDim.x >> 5)) >= k) return;
    const DcNode nd = dc_node(a, blockIdx.y);
    double* sd = sm_dyn;
    double* sz = sm_dyn + k;
    for (int i = threadIdx.x; i < k; i += blockDim.x) { sd[i] = a.dk[nd.lo + i]; sz[i] = a.zk[nd.lo + i]; }
    __syncthreads();
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j >= k) return;
    const double rho = 2.0 * fabs(a.e[nd.mid - 1]);
    double lam;
    const int st = eigcore::secular_root(k, j, sd, sz, rho, a.U + nd.lo + (long long)a.n * (nd.lo + j), 1, &lam, WarpPar());
    if ((threadIdx.x & 31) == 0) {
        a.lam2[nd.lo + j] = lam;
        if (st) atomicAdd(a.fail, 1);
    }
}

// Loewner formula for the update vector that makes the computed roots exact.  CTA = 32 entries x 8 slices of the roots
__global__ void __launch_bounds__(256) k_dc_zhat(DcArgs a) {
    if (!dc_select(a)) return;
    __shared__ double sp[8][33];
    const int k = a.kcnt[blockIdx.y];
    if ((int)(blockIdx.x * 32) >= k) return;
    const DcNode nd = dc_node(a, blockIdx.y);
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + tx;
    const long long ld = a.n;
    const double* dk = a.dk + nd.lo;
    double w = 1.0;
    if (i < k) {
        const double* U = a.U + (nd.lo + i) + ld * nd.lo;
        const double di = dk[i];
        for (int j = ty; j < k; j += 8) w *= (j == i) ? U[ld * j] : U[ld * j] / (di - dk[j]);
    }
    sp[ty][tx] = w;
    __syncthreads();
    if (ty == 0 && i < k) {
        for (int q = 1; q < 8; ++q) w *= sp[q][tx];
        a.zhat[nd.lo + i] = copysign(sqrt(fabs(w)), a.zk[nd.lo + i]);
    }
}

// eigenvectors of the rank-one update: U(i,j) = zhat_i / delta(i,j), normalised, rows stored in grouped order; warp per column
__global__ void __launch_bounds__(256) k_dc_vec(DcArgs a) {
    if (!dc_select(a)) return;
    const int k = a.kcnt[blockIdx.y];
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (j >= k) return;
    const DcNode nd = dc_node(a, blockIdx.y);
    const double* U = a.U + nd.lo + (long long)a.n * (nd.lo + j);
    double* Ug = a.Ug + nd.lo + (long long)a.n * (nd.lo + j);
    const double* zh = a.zhat + nd.lo;
    const int* pos = a.pos + nd.lo;
    double s = 0.0;
    for (int i = lane; i < k; i += 32) { const double u = zh[i] / U[i]; s += u * u; }
    s = warp_sum(s);
    const double sc = 1.0 / sqrt(s);
    for (int i = lane; i < k; i += 32) Ug[pos[i]] = zh[i] / U[i] * sc;
}

// GEMM tasks of the level, two per node: the top rows of the merged eigenvectors need the left-only and dense columns,
// the bottom rows the dense and right-only ones; thread per node
__global__ void k_dc_tasks(DcArgs a, int nodes) {
    if (!dc_select(a)) return;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nodes) return;
    const DcNode nd = dc_node(a, t);
    const int ld = a.n, k = a.kcnt[t], c0 = a.tcnt[3 * t], c1 = a.tcnt[3 * t + 1], c2 = a.tcnt[3 * t + 2];
    const long long off = nd.lo + (long long)ld * nd.lo;
    a.tasks[2 * t] = gemm_task_dev(a.Qg + off, ld, a.Ug + off, ld, a.Qn + off, ld, nd.mid - nd.lo, k, c0 + c1);
    a.tasks[2 * t + 1] = gemm_task_dev(a.Qg + nd.mid + (long long)ld * (nd.lo + c0), ld, a.Ug + (nd.lo + c0) + (long long)ld * nd.lo, ld,
                                       a.Qn + nd.mid + (long long)ld * nd.lo, ld, nd.hi - nd.mid, k, c1 + c2);
}

// non-deflated eigenvectors into their grouped slots of Qg, deflated ones unchanged behind the updated ones in Qn;
// grid (row chunks, slot, node)
__global__ void __launch_bounds__(256) k_dc_move(DcArgs a) {
    const DcNode nd = dc_node(a, blockIdx.z);
    const int m = nd.hi - nd.lo, p = blockIdx.y;
    if (p >= m) return;
    const int row = nd.lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= nd.hi) return;
    const int k = a.kcnt[blockIdx.z];
    const long long ld = a.n;
    const double v = a.Qo[row + ld * (nd.lo + a.cols[nd.lo + p])];
    if (p < k) a.Qg[row + ld * (nd.lo + a.pos[nd.lo + p])] = v;
    else a.Qn[row + ld * (nd.lo + p)] = v;
}

// final ascending order: rank of every eigenvalue, sorted eigenvalues
__global__ void __launch_bounds__(256) k_dc_rank(int n, const double* __restrict__ lam, int* __restrict__ rank,
                                                 double* __restrict__ w_sorted) {
    extern __shared__ double sm_dyn[];
    for (int i = threadIdx.x; i < n; i += blockDim.x) sm_dyn[i] = lam[i];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = sm_dyn[i];
    int r = 0;
    for (int k = 0; k < n; ++k) { const double w = sm_dyn[k]; r += (w < v || (w == v && k < i)) ? 1 : 0; }
    rank[i] = r;
    w_sorted[r] = v;
}
__global__ void k_dc_permute(int n, const double* __restrict__ Q, const int* __restrict__ rank, double* __restrict__ Z) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)n * n) return;
    const int r = (int)(t % n), j = (int)(t / n);
    Z[r + (long long)n * rank[j]] = Q[t];
}

// ---------------------------------------------------------------------------------------------------------------
// batch of small matrices of one order (conformer ensembles, n <= kCtaEigMax): the same three stages, every launch
// covering all matrices (blockIdx.z or .x = matrix).  The tridiagonalisation runs with one CTA per matrix, the matrix
// resident in its shared memory (k_sytrd_smem with a single owner: no grid barrier at all).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kCtaEigMax = 160;
constexpr int kCluSize = 8;        // CTAs per cluster of the batched mid-size tridiagonalisation (portable maximum)

// A: matrices ms apart (n x n, both triangles; on exit the reflectors below the sub-diagonal); d, e, tau: vs apart
__global__ void __launch_bounds__(kSytrdThreads, 1) k_sytrd_cta(double* __restrict__ Aall, long long ms, int n, double* __restrict__ dall,
                                                                double* __restrict__ eall, double* __restrict__ tauall, long long vs,
                                                                const int* __restrict__ skip) {
    const long long ib = blockIdx.x;
    if (skip && skip[ib]) return;
    extern __shared__ double sm_dyn[];
    __shared__ double red[32];
    double* cols = sm_dyn;                 // n x n, column-major
    double* sv = sm_dyn + (size_t)n * n;
    double* sw = sv + n;
    double* A = Aall + ib * ms;
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
    for (int t = tid; t < n * n; t += nthr) cols[t] = A[t];
    __syncthreads();
    for (int jj = 0; jj < n; ++jj) {
        const int mm = n - jj - 1;
        double* col = cols + (size_t)jj * n;
        double sq = 0.0;
        for (int r = jj + 2 + tid; r < n; r += nthr) sq += col[r] * col[r];
        sq = block_sum_sh(sq, red);
        if (tid == 0) dall[ib * vs + jj] = col[jj];
        if (mm == 0) break;
        const double alpha = col[jj + 1];
        double beta = alpha, tau = 0.0, scal = 0.0;
        if (sq > 0.0) {
            beta = -copysign(sqrt(alpha * alpha + sq), alpha);
            tau = (beta - alpha) / beta;
            scal = 1.0 / (alpha - beta);
        }
        if (tid == 0) { eall[ib * vs + jj] = beta; tauall[ib * vs + jj] = tau; }
        for (int r = jj + 1 + tid; r < n; r += nthr) {
            const double v = (r == jj + 1) ? 1.0 : col[r] * scal;
            sv[r] = v;
            if (r >= jj + 2) col[r] = v;             // reflector storage (this column is not touched again)
        }
        __syncthreads();
        // y = A v over the trailing block: warp per column
        for (int c = jj + 1 + warp; c < n; c += nwarp) {
            const double* cc = cols + (size_t)c * n;
            double s = 0.0;
            for (int r = jj + 1 + lane; r < n; r += 32) s += cc[r] * sv[r];
            s = warp_sum(s);
            if (lane == 0) sw[c] = s;
        }
        __syncthreads();
        double vy = 0.0;
        for (int r = jj + 1 + tid; r < n; r += nthr) vy += sv[r] * sw[r];
        vy = block_sum_sh(vy, red);
        const double coef = -0.5 * tau * tau * vy;
        __syncthreads();
        for (int r = jj + 1 + tid; r < n; r += nthr) sw[r] = tau * sw[r] + coef * sv[r];
        __syncthreads();
        for (int t = tid; t < mm * mm; t += nthr) {
            const int r = jj + 1 + t % mm, c = jj + 1 + t / mm;
            cols[(size_t)c * n + r] -= sv[r] * sw[c] + sw[r] * sv[c];
        }
        __syncthreads();
    }
    __syncthreads();
    for (int t = tid; t < n * n; t += nthr) A[t] = cols[t];
}

// grouped copies of the non-deflated eigenvectors / deflated ones behind (k_dc_move for the batch): grid (slot, node,
// matrix), the threads run over the rows
__global__ void __launch_bounds__(64) k_dc_move_b(DcArgs a) {
    if (!dc_select(a)) return;
    const DcNode nd = dc_node(a, blockIdx.y);
    const int m = nd.hi - nd.lo, p = blockIdx.x;
    if (p >= m) return;
    const int k = a.kcnt[blockIdx.y];
    const long long ld = a.n;
    const double* src = a.Qo + ld * (nd.lo + a.cols[nd.lo + p]);
    double* dst = p < k ? a.Qg + ld * (nd.lo + a.pos[nd.lo + p]) : a.Qn + ld * (nd.lo + p);
    for (int row = nd.lo + threadIdx.x; row < nd.hi; row += blockDim.x) dst[row] = src[row];
}
// final order of the batch: rank, sorted eigenvalues into w (stride ws), Z[:, rank[j]] = Q[:, j]; CTA per matrix
__global__ void __launch_bounds__(256) k_dc_finish_b(DcArgs a, int* __restrict__ rankall, double* __restrict__ wall, long long ws,
                                                     double* __restrict__ Zall) {
    const long long ib = blockIdx.z;
    if (!dc_select(a)) return;
    extern __shared__ double sm_dyn[];
    const int n = a.n;
    int* rank = rankall + ib * n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) sm_dyn[i] = a.lam[i];
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double v = sm_dyn[i];
        int r = 0;
        for (int k = 0; k < n; ++k) { const double w = sm_dyn[k]; r += (w < v || (w == v && k < i)) ? 1 : 0; }
        rank[i] = r;
        wall[ib * ws + r] = v;
    }
    __syncthreads();
    double* Z = Zall + ib * a.ms;
    for (int t = threadIdx.x; t < n * n; t += blockDim.x) Z[t % n + (long long)n * rank[t / n]] = a.Qo[t];
}
// explicit V of block [jb, jb + kb) for every matrix; grid (elements, 1, matrix)
__global__ void k_bt_formv_b(const double* __restrict__ Aall, long long ms, int n, int jb, int kb, double* __restrict__ Vball,
                             long long vbs, const int* __restrict__ skip) {
    const long long ib = blockIdx.z;
    if (skip && skip[ib]) return;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * kb) return;
    const int r = t % n, k = t / n, piv = jb + k + 1;
    Vball[ib * vbs + t] = r < piv ? 0.0 : (r == piv ? 1.0 : Aall[ib * ms + r + (long long)n * (jb + k)]);
}
// T factor of one block for every matrix from G = V^T V (full kb x kb product, only k < i is used); CTA per matrix
__global__ void __launch_bounds__(NB * kTfParts) k_bt_tfactor_b(const double* __restrict__ Gall, const double* __restrict__ tauall,
                                                                long long vs, int jb, int kb, double* __restrict__ Tall,
                                                                const int* __restrict__ skip) {
    const long long ib = blockIdx.x;
    if (skip && skip[ib]) return;
    extern __shared__ double sm_dyn[];
    double (*T)[NB + 1] = (double (*)[NB + 1])sm_dyn;                   // NB x (NB + 1)
    double (*part)[NB] = (double (*)[NB])(sm_dyn + NB * (NB + 1));     // kTfParts x NB
    const int r = threadIdx.x % NB, pt = threadIdx.x / NB;             // kTfParts threads per row, as in k_bt_tfactor
    const double* G = Gall + ib * NB * NB;
    const double* tau = tauall + ib * vs;
    for (int c = pt; c < NB; c += kTfParts) T[r][c] = 0.0;
    __syncthreads();
    for (int i = 0; i < kb; ++i) {
        double s = 0.0;
        if (r < i)
            for (int k = r + pt; k < i; k += kTfParts) s += T[r][k] * G[k + NB * i];
        part[pt][r] = s;
        __syncthreads();
        if (pt == 0) {
            if (r < i) {
                double t = 0.0;
#pragma unroll
                for (int q = 0; q < kTfParts; ++q) t += part[q][r];
                T[r][i] = -tau[jb + i] * t;
            } else if (r == i) T[r][i] = tau[jb + i];
        }
        __syncthreads();
    }
    double* To = Tall + ib * NB * NB;
    for (int c = pt; c < NB; c += kTfParts) To[r + NB * c] = T[r][c];
}
// GEMM tasks of one back-transformation step for every matrix; which: 0 G = V^T V, 1 W = V^T Z, 2 VT = V T, 3 Z -= VT W
__global__ void k_bt_tasks_b(GemmTask* __restrict__ tasks, int nb, int which, int n, int jb, int kb, double* Vb, long long vbs,
                             double* G, double* T, double* Wz, double* VT, double* Z, long long ms) {
    const long long ib = blockIdx.x * blockDim.x + threadIdx.x;
    if (ib >= nb) return;
    const int r0 = jb + 1, mr = n - r0;
    double* V = Vb + ib * vbs;
    GemmTask t;
    if (which == 0) t = gemm_task_dev(V + r0, n, V + r0, n, G + ib * NB * NB, NB, kb, kb, mr);
    else if (which == 1) t = gemm_task_dev(V + r0, n, Z + ib * ms + r0, n, Wz + ib * vbs, NB, kb, n, mr);
    else if (which == 2) t = gemm_task_dev(V + r0, n, T + ib * NB * NB, NB, VT + ib * vbs + r0, n, mr, kb, kb);
    else { t = gemm_task_dev(VT + ib * vbs + r0, n, Wz + ib * vbs, NB, Z + ib * ms + r0, n, mr, n, kb); t.alpha = -1.0; t.beta = 1.0; }
    tasks[ib] = t;
}
// per-matrix "leave out" flags of the batch from the SCC state
__global__ void k_eig_skip_b(Dev d, int* __restrict__ skip) {
    const int im = blockIdx.x * blockDim.x + threadIdx.x;
    if (im < d.nmol) skip[im] = d.st[im].done ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------------------------------------------
struct EighWork {
    int nmax = 0, nsm = 0, grid = 0;
    double *Qa = nullptr, *Qb = nullptr, *U = nullptr, *Qg = nullptr, *Ug = nullptr;
    double *PB = nullptr, *Vb = nullptr, *Wz = nullptr, *VT = nullptr, *Tf = nullptr, *Gf = nullptr;
    double *dd = nullptr, *ee = nullptr, *tau = nullptr, *y = nullptr, *part = nullptr, *vtv = nullptr, *wtv = nullptr;
    double *lam = nullptr, *lam2 = nullptr, *z = nullptr, *zhat = nullptr, *dk = nullptr, *zk = nullptr, *rc = nullptr, *rs = nullptr;
    int *perm = nullptr, *cols = nullptr, *pos = nullptr, *tcnt = nullptr, *ctype = nullptr, *ra = nullptr, *rb = nullptr, *kcnt = nullptr, *nrot = nullptr,
        *fail = nullptr, *rank = nullptr;
    unsigned* bar = nullptr;
    double *cbuf = nullptr, *ybuf = nullptr;
    size_t res_smem_cap = 0;
    int res_grid_want = 0;   // CTAs of the resident tridiagonalisation (0: automatic)
    size_t smem_limit = 0;
    size_t panel_smem_cap = 0; // dynamic shared memory the panel kernel may use
    double *sym_colbuf = nullptr, *sym_rowbuf = nullptr;   // tile sums of the symmetric trailing-matrix product
    bool sym_off = false;
    int resident_max = 0;      // largest trailing order the shared-memory resident tridiagonalisation takes
    long long* prof = nullptr;
    GemmTask* tasks = nullptr;
    std::vector<void*> allocs;
    size_t bytes = 0;
    float ms[3] = {0, 0, 0};   // tridiagonalisation, divide and conquer, back-transformation of the last timed call
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    int launches = 0;

    // tensor map of a column-major n x n FP64 matrix with boxes of kSyRH rows x kSyCW columns (symmetric panel product)
    typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    EncodeTiledFn encode_tiled = nullptr;
    bool make_tile_map(CUtensorMap* map, double* A, int n) const {
        if (!encode_tiled) return false;
        const cuuint64_t gdim[2] = {(cuuint64_t)n, (cuuint64_t)n}, gstride[1] = {(cuuint64_t)n * sizeof(double)};
        const cuuint32_t box[2] = {(cuuint32_t)kSyRH, (cuuint32_t)kSyCW}, estr[2] = {1, 1};
        return encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, A, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    }

    template <class T>
    T* alloc(size_t cnt) {
        void* p = nullptr;
        const size_t b = std::max<size_t>(cnt, 1) * sizeof(T);
        XTB_CUDA_CHECK(cudaMalloc(&p, b));
        XTB_CUDA_CHECK(cudaMemset(p, 0, b));
        allocs.push_back(p);
        bytes += b;
        return (T*)p;
    }
    static size_t sytrd_smem(int n, int grid) { return ((size_t)n + (size_t)(n / grid + 6) * 8) * sizeof(double); }
    // largest order: shared-memory staging of the secular solver (16 bytes per root) and of the panel kernel's vector
    static int max_order() { return 14000; }

    void init(int nmax_, int device) {
        nmax = nmax_;
        if (nmax > max_order()) throw std::runtime_error("xtb_b200: eigensolver supports matrices up to order 14000");
        cudaDeviceProp prop;
        XTB_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
        nsm = prop.multiProcessorCount;
        if (!prop.cooperativeLaunch) throw std::runtime_error("xtb_b200: device lacks cooperative launch");
        const size_t n = nmax, n2 = n * n;
        Qa = alloc<double>(n2); Qb = alloc<double>(n2); U = alloc<double>(n2); Qg = alloc<double>(n2); Ug = alloc<double>(n2);
        PB = alloc<double>(n * (NB + EB)); Vb = alloc<double>(n * NB);
        Wz = alloc<double>((size_t)kBtSplit * NB * n); VT = alloc<double>((size_t)NB * n);
        const size_t nblk = (n + NB - 1) / NB;
        Tf = alloc<double>(nblk * NB * NB); Gf = alloc<double>(nblk * NB * NB);
        dd = alloc<double>(n); ee = alloc<double>(n); tau = alloc<double>(n); y = alloc<double>(n);
        part = alloc<double>(2 * (size_t)nsm); vtv = alloc<double>(EB); wtv = alloc<double>(EB);
        lam = alloc<double>(n); lam2 = alloc<double>(n); z = alloc<double>(n); zhat = alloc<double>(n);
        dk = alloc<double>(n); zk = alloc<double>(n); rc = alloc<double>(n); rs = alloc<double>(n);
        perm = alloc<int>(n); cols = alloc<int>(n); pos = alloc<int>(n); tcnt = alloc<int>(3 * n); ctype = alloc<int>(n); ra = alloc<int>(n); rb = alloc<int>(n);
        kcnt = alloc<int>(n); nrot = alloc<int>(n); fail = alloc<int>(1); rank = alloc<int>(n);
        bar = alloc<unsigned>(1);
        cbuf = alloc<double>(2 * n); ybuf = alloc<double>(2 * n);
        prof = alloc<long long>(16);
        tasks = alloc<GemmTask>(2 * n);
        grid = nsm;
        {   // the panel kernel may use all the shared memory its static part leaves (tile ring of the symmetric form)
            cudaFuncAttributes fa;
            XTB_CUDA_CHECK(cudaFuncGetAttributes(&fa, k_sytrd_panel));
            panel_smem_cap = (size_t)prop.sharedMemPerBlockOptin - fa.sharedSizeBytes;
            if (sytrd_smem(nmax, grid) > panel_smem_cap) throw std::runtime_error("xtb_b200: tridiagonalisation kernel: matrix order too large");
            XTB_CUDA_CHECK(cudaFuncSetAttribute(k_sytrd_panel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)panel_smem_cap));
            const size_t nS = (n + kSyCW - 1) / kSyCW, nR = (n + kSyRH - 1) / kSyRH;
            sym_colbuf = alloc<double>(nS * nR * kSyCW);
            sym_rowbuf = alloc<double>((nR + (size_t)nsm + 1) * kSyRH);
            sym_off = tuning_env("XTBB200_NO_SYM_SYTRD") != nullptr;
            void* fn = nullptr;
            cudaDriverEntryPointQueryResult qr;
            if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
                encode_tiled = (EncodeTiledFn)fn;
            cudaGetLastError();
        }
        XTB_CUDA_CHECK(cudaFuncSetAttribute(k_dc_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(n * 8)));
        smem_limit = (size_t)prop.sharedMemPerBlockOptin - 1024;
        XTB_CUDA_CHECK(cudaFuncSetAttribute(k_dc_deflate, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)std::min(n * 24 + 16, smem_limit)));
        XTB_CUDA_CHECK(cudaFuncSetAttribute(k_dc_secular, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(n * 16)));
        XTB_CUDA_CHECK(cudaFuncSetAttribute(k_dc_rank, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(n * 8)));
        XTB_CUDA_CHECK(cudaFuncSetAttribute(k_bt_tfactor, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((NB * (NB + 1) + kTfParts * NB) * 8)));
        resident_max = 0;
        const size_t smem_cap = (size_t)prop.sharedMemPerBlockOptin - 1024;   // static shared memory of the kernel
        for (int m = 1; m <= std::min(nmax, kResRows * kSytrdThreads); ++m)
            if (sytrd_smem_resident_bytes(m, grid) <= smem_cap) resident_max = m;
        if (tuning_env("XTBB200_NO_RESIDENT_SYTRD")) resident_max = 0;
        res_smem_cap = smem_cap;
        if (resident_max > 0)   // smaller grids than one CTA per SM hold more columns per CTA: allow the full capacity
        {
            XTB_CUDA_CHECK(cudaFuncSetAttribute(k_sytrd_smem<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap));
            XTB_CUDA_CHECK(cudaFuncSetAttribute(k_sytrd_smem<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap));
            XTB_CUDA_CHECK(cudaFuncSetAttribute(k_sytrd_smem<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap));
            XTB_CUDA_CHECK(cudaFuncSetAttribute(k_sytrd_smem<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap));
            clu16_ok = false;
            if (!tuning_env("XTBB200_NO_CLUSTER16") &&
                cudaFuncSetAttribute(k_sytrd_smem<2, true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
                cudaLaunchConfig_t cfg;
                std::memset(&cfg, 0, sizeof(cfg));
                cfg.gridDim = dim3(2 * kCluSize); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = smem_cap;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeClusterDimension;
                at[0].val.clusterDim.x = 2 * kCluSize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                cfg.attrs = at; cfg.numAttrs = 1;
                int ncl = 0;
                if (cudaOccupancyMaxActiveClusters(&ncl, k_sytrd_smem<2, true>, &cfg) == cudaSuccess && ncl >= 1) clu16_ok = true;
                cudaGetLastError();
            }
        }
        if (const char* e = tuning_env("XTBB200_RES_GRID")) res_grid_want = std::max(1, std::atoi(e));   // tuning knob
        int occ = 0;
        XTB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_sytrd_panel, kSytrdThreads, panel_smem_cap));
        if (occ < 1) throw std::runtime_error("xtb_b200: tridiagonalisation kernel does not fit on an SM");
        for (auto& e : ev) cudaEventCreate(&e);
    }
    void release() {
        for (void* p : allocs) cudaFree(p);
        allocs.clear();
        for (auto& e : ev) if (e) { cudaEventDestroy(e); e = nullptr; }
    }

    // ---- batch of nb matrices of one small order n (conformer ensembles): own scratch, every launch covers the batch
    int bnb = 0, bn = 0;
    bool clu16_ok = false;                       // clusters of 2 kCluSize CTAs can be scheduled (checked in init)
    double *bcbuf = nullptr, *bybuf = nullptr;   // exchange buffers of the cluster tridiagonalisation, 2 n per matrix each
    double *bQa = nullptr, *bQb = nullptr, *bU = nullptr, *bQg = nullptr, *bUg = nullptr, *bVb = nullptr, *bVT = nullptr, *bWz = nullptr,
           *bG = nullptr, *bT = nullptr, *bvec = nullptr;
    int *bint = nullptr, *bskip = nullptr, *bfail = nullptr;
    GemmTask *btasks = nullptr, *bbt = nullptr;
    void init_batch(int nb, int n) {
        if (n > cluster_max()) throw std::runtime_error("xtb_b200: batched eigensolver: matrix order beyond the cluster-resident tridiagonalisation");
        bnb = nb; bn = n;
        const size_t n2 = (size_t)n * n, m = (size_t)nb * n2, v = (size_t)nb * n;
        bQa = alloc<double>(m); bQb = alloc<double>(m); bU = alloc<double>(m); bQg = alloc<double>(m); bUg = alloc<double>(m);
        bVb = alloc<double>((size_t)nb * n * NB); bVT = alloc<double>((size_t)nb * n * NB); bWz = alloc<double>((size_t)nb * n * NB);
        bG = alloc<double>((size_t)nb * NB * NB); bT = alloc<double>((size_t)nb * NB * NB);
        bvec = alloc<double>(11 * v);
        bint = alloc<int>(12 * v);
        bskip = alloc<int>(nb); bfail = alloc<int>(nb);
        btasks = alloc<GemmTask>(2 * v); bbt = alloc<GemmTask>(nb);
        if (n <= kCtaEigMax)
            XTB_CUDA_CHECK(cudaFuncSetAttribute(k_sytrd_cta, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((n2 + 2 * n) * sizeof(double))));
        else { bcbuf = alloc<double>(2 * v); bybuf = alloc<double>(2 * v); }
        XTB_CUDA_CHECK(cudaFuncSetAttribute(k_bt_tfactor_b, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((NB * (NB + 1) + kTfParts * NB) * 8)));
    }
    // largest order of the batched solver: up to kCtaEigMax one CTA holds a matrix (k_sytrd_cta); beyond, a cluster of
    // kCluSize CTAs holds it in their shared memories (k_sytrd_smem<2, true>: the resident kernel with the cluster's hardware
    // barrier in place of the grid barrier), 148 / kCluSize matrices at a time
    // CTAs per cluster for order n: kCluSize (portable) while the matrix fits (n <= 472), 2 kCluSize beyond (n <= ~670;
    // non-portable cluster size: used only when the device can schedule such a cluster with this shared-memory footprint;
    // XTBB200_NO_CLUSTER16 switches it off); 0: does not fit
    int cluster_size(int n) const {
        for (int c = kCluSize; c <= (clu16_ok ? 2 * kCluSize : kCluSize); c *= 2)
            if (sytrd_smem_resident_bytes(n, c) <= res_smem_cap && (n + c - 1) / c <= 96 && n <= 1024) return c;
        return 0;
    }
    int cluster_max() const {
        int m = kCtaEigMax;
        for (int t = kCtaEigMax + 1; t <= 1024; ++t)
            if (cluster_size(t)) m = t; else break;
        return m;
    }
    // A: nb matrices n^2 apart (destroyed), w: eigenvalues n apart, Z: eigenvectors n^2 apart; bskip[] selects (k_eig_skip_b)
    int solve_batch(cudaStream_t st, double* A, double* w, double* Z) {
        const int nb = bnb, n = bn;
        const long long vs = n, ms = (long long)n * n, vbs = (long long)n * NB;
        const size_t v = (size_t)nb * n;
        int nl = 0;
        double *dd_ = bvec, *ee_ = bvec + v, *tau_ = bvec + 2 * v;
        XTB_CUDA_CHECK(cudaMemsetAsync(bfail, 0, nb * sizeof(int), st));
        if (n <= kCtaEigMax) {
            k_sytrd_cta<<<nb, kSytrdThreads, ((size_t)n * n + 2 * n) * sizeof(double), st>>>(A, ms, n, dd_, ee_, tau_, vs, bskip);
        } else {
            SytrdSmemArgs ra;
            std::memset(&ra, 0, sizeof(ra));
            ra.A = A; ra.n = n; ra.j0 = 0; ra.d = dd_; ra.e = ee_; ra.tau = tau_; ra.cbuf = bcbuf; ra.ybuf = bybuf;
            const int clu = cluster_size(n);
            ra.clu = clu; ra.ms = ms; ra.vs = vs; ra.skip = bskip;
            int tres = kSytrdThreads;
            while (tres > 256 && kResRows * (tres / 2) >= n) tres /= 2;
            cudaLaunchConfig_t cfg;
            std::memset(&cfg, 0, sizeof(cfg));
            cfg.gridDim = dim3((unsigned)nb * clu); cfg.blockDim = dim3(tres);
            cfg.dynamicSmemBytes = sytrd_smem_resident_bytes(n, clu); cfg.stream = st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = clu; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            XTB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, k_sytrd_smem<2, true>, ra));
        }
        ++nl;
        XTB_CUDA_CHECK(cudaMemsetAsync(bQa, 0, (size_t)nb * ms * sizeof(double), st));
        XTB_CUDA_CHECK(cudaMemsetAsync(bQb, 0, (size_t)nb * ms * sizeof(double), st));
        DcArgs a;
        std::memset(&a, 0, sizeof(a));
        a.n = n; a.d = dd_; a.e = ee_; a.vs = vs; a.ms = ms; a.skip = bskip;
        a.Qo = bQa; a.Qn = bQb; a.U = bU; a.Qg = bQg; a.Ug = bUg;
        a.lam = bvec + 3 * v; a.lam2 = bvec + 4 * v; a.z = bvec + 5 * v; a.zhat = bvec + 6 * v; a.dk = bvec + 7 * v; a.zk = bvec + 8 * v;
        a.rc = bvec + 9 * v; a.rs = bvec + 10 * v;
        a.perm = bint; a.cols = bint + v; a.pos = bint + 2 * v; a.tcnt = bint + 3 * v; a.ctype = bint + 6 * v; a.ra = bint + 7 * v;
        a.rb = bint + 8 * v; a.kcnt = bint + 9 * v; a.nrot = bint + 10 * v;
        int* rank_ = bint + 11 * v;
        a.fail = bfail; a.tasks = btasks;
        if (n < 2) throw std::runtime_error("xtb_b200: batched eigensolver needs n > 1");
        const int D = eigcore::tree_depth(n);
        a.depth = D;
        k_dc_leaves<<<dim3(((1 << D) + 127) / 128, 1, nb), 128, 0, st>>>(a);
        ++nl;
        for (int depth = D - 1; depth >= 0; --depth) {
            const int nodes = 1 << depth, mmax = (n + nodes - 1) / nodes;
            a.depth = depth;
            k_dc_sort<<<dim3((mmax + 255) / 256, nodes, nb), 256, (size_t)mmax * 8, st>>>(a);
            k_dc_deflate<<<dim3(nodes, 1, nb), 256, (size_t)mmax * 24 + 16, st>>>(a, 1);
            k_dc_rot<<<dim3((mmax + 255) / 256, nodes, nb), 256, 0, st>>>(a);
            k_dc_move_b<<<dim3(mmax, nodes, nb), 64, 0, st>>>(a);
            k_dc_secular<<<dim3((mmax + 7) / 8, nodes, nb), 256, (size_t)mmax * 16, st>>>(a);
            k_dc_zhat<<<dim3((mmax + 31) / 32, nodes, nb), 256, 0, st>>>(a);
            k_dc_vec<<<dim3((mmax + 7) / 8, nodes, nb), 256, 0, st>>>(a);
            k_dc_tasks<<<dim3((nodes + 127) / 128, 1, nb), 128, 0, st>>>(a, nodes);
            const int th = (mmax / 2 + 1 + GT - 1) / GT, tw = (mmax + GT - 1) / GT;
            k_eg_gemm<false, false><<<dim3(th * tw, 2 * nodes, nb), 128, 0, st>>>(GemmTask(), btasks, 2 * n, bskip);
            nl += 9;
            std::swap(a.Qo, a.Qn);
            std::swap(a.lam, a.lam2);
        }
        k_dc_finish_b<<<dim3(1, 1, nb), 256, (size_t)n * 8, st>>>(a, rank_, w, vs, Z);
        ++nl;
        const int nrefl = n - 1, nbb = (nrefl + NB - 1) / NB;
        const int tz = (n + GT - 1) / GT;
        for (int b = nbb - 1; b >= 0; --b) {
            const int jb = b * NB, kb = std::min(NB, nrefl - jb), tk = (kb + GT - 1) / GT;
            k_bt_formv_b<<<dim3((n * kb + 255) / 256, 1, nb), 256, 0, st>>>(A, ms, n, jb, kb, bVb, vbs, bskip);
            k_bt_tasks_b<<<(nb + 127) / 128, 128, 0, st>>>(bbt, nb, 0, n, jb, kb, bVb, vbs, bG, bT, bWz, bVT, Z, ms);
            k_eg_gemm<true, false><<<dim3(tk * tk, 1, nb), 128, 0, st>>>(GemmTask(), bbt, 1, bskip);
            k_bt_tfactor_b<<<nb, NB * kTfParts, (NB * (NB + 1) + kTfParts * NB) * sizeof(double), st>>>(bG, tau_, vs, jb, kb, bT, bskip);
            k_bt_tasks_b<<<(nb + 127) / 128, 128, 0, st>>>(bbt, nb, 1, n, jb, kb, bVb, vbs, bG, bT, bWz, bVT, Z, ms);
            k_eg_gemm<true, false><<<dim3(tk * tz, 1, nb), 128, 0, st>>>(GemmTask(), bbt, 1, bskip);
            k_bt_tasks_b<<<(nb + 127) / 128, 128, 0, st>>>(bbt, nb, 2, n, jb, kb, bVb, vbs, bG, bT, bWz, bVT, Z, ms);
            k_eg_gemm<false, false><<<dim3(tz * tk, 1, nb), 128, 0, st>>>(GemmTask(), bbt, 1, bskip);
            k_bt_tasks_b<<<(nb + 127) / 128, 128, 0, st>>>(bbt, nb, 3, n, jb, kb, bVb, vbs, bG, bT, bWz, bVT, Z, ms);
            k_eg_gemm<false, false><<<dim3(tz * tz, 1, nb), 128, 0, st>>>(GemmTask(), bbt, 1, bskip);
            nl += 10;
        }
        return nl;
    }

    // eligibility of the bulk-copy kernel: alignment of every copied line and whole k-tiles
    static bool gemm_tma_ok(const GemmTask& t, int nz) {
        // Opt-in since the end of round 2 (XTBB200_TUNING=1 XTBB200_TMA_GEMM=1): at three CTAs per SM (__launch_bounds__(160, 3))
        // the kernel gave wrong trailing updates for orders with ragged edge tiles (n = 3006, 6000; n = 2432 was right), varying
        // from run to run, while two CTAs per SM passed every test -- an unexplained race is not acceptable on the product path,
        // and the register-staged kernel is as fast (17.5 vs 17.7 TFLOP/s on the K = 128 update).
        static const bool on = tuning_env("XTBB200_TMA_GEMM") != nullptr;
        if (!on) return false;
        auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
        if (!al16(t.A) || !al16(t.B) || (t.lda & 1) || (t.ldb & 1) || (t.M & 1) || (t.N & 1)) return false;
        if (t.K % GK || t.K < 2 * GK || t.ksegA % GK || t.ksegB % GK || (t.offA2 & 1) || (t.offB2 & 1)) return false;
        if (t.ksplit > 1 && nz != t.ksplit) return false;
        return nz == 1 || t.ksplit > 1;
    }
    template <bool TA, bool TB>
    static void gemm_tma(cudaStream_t st, const GemmTask& t, const dim3& g) {
        static bool once = [] {
            cudaFuncSetAttribute(k_eg_gemm_tma<TA, TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)eg_tma_smem<TA, TB>());
            return true;
        }();
        (void)once;
        k_eg_gemm_tma<TA, TB><<<g, 160, eg_tma_smem<TA, TB>(), st>>>(t);
    }
    static void gemm(cudaStream_t st, bool ta, bool tb, const GemmTask& t, int nz = 1) {
        const int tm = (t.M + GT - 1) / GT, tn = (t.N + GT - 1) / GT;
        if (tm * tn == 0) return;
        const dim3 g(tm * tn, 1, nz);
        // the bulk-copy kernel only for the shape it is verified and pays off for: A not transposed, B transposed (trailing update:
        // 32 line copies of 512 B per k-tile).  With 128-byte lines (TA / !TB: 128 copies per k-tile) the producer warp cannot
        // issue fast enough (every copy is a separate uniform-datapath instruction), measured 2x slower than register staging.
        if (!ta && tb && gemm_tma_ok(t, nz)) {
            if (!ta && !tb) gemm_tma<false, false>(st, t, g);
            else if (ta && !tb) gemm_tma<true, false>(st, t, g);
            else if (!ta && tb) gemm_tma<false, true>(st, t, g);
            else gemm_tma<true, true>(st, t, g);
            return;
        }
        if (!ta && !tb) k_eg_gemm<false, false><<<g, 128, 0, st>>>(t, nullptr);
        else if (ta && !tb) k_eg_gemm<true, false><<<g, 128, 0, st>>>(t, nullptr);
        else if (!ta && tb) k_eg_gemm<false, true><<<g, 128, 0, st>>>(t, nullptr);
        else k_eg_gemm<true, true><<<g, 128, 0, st>>>(t, nullptr);
    }

    // A (n x n, leading dimension n, both triangles; destroyed) -> eigenvalues w[n] ascending, eigenvectors Z (n x n).
    // *fail (device int) is incremented when a secular equation did not converge.  Returns the number of launches.
    // hook(stage, end, launches) brackets the stages for the caller's timers: 0 panel kernel, 1 trailing update,
    // 2 divide and conquer, 3 back-transformation.
    // zc0 / zc1 (multiples of GT, zc1 < 0: all): the back-transformation is applied to the columns [zc0, zc1) of Z only -- one
    // large system with the column-parallel products split over the GPUs; the caller gathers the column blocks afterwards.
    // Tile boundaries and the K split are those of the full matrix, so the columns are bit-identical to the unsplit run.
    template <class Hook>
    int solve(cudaStream_t st, double* A, int n, double* w, double* Z, bool timed, Hook hook, int zc0 = 0, int zc1 = -1) {
        int nl = 0;
        const long long ld = n;
        const size_t n2 = (size_t)n * n;
        if (timed) cudaEventRecord(ev[0], st);
        // ---- 1. tridiagonalisation; panel buffer PB = [V of even panels | V of odd panels | W]
        const int npan = (n + EB - 1) / EB, nblk = (n + NB - 1) / NB;
        double* Wp = PB + (size_t)n * NB;
        XTB_CUDA_CHECK(cudaMemsetAsync(Gf, 0, (size_t)nblk * NB * NB * sizeof(double), st));
        int jres = -1;   // column from which the shared-memory resident kernel takes over (block reflector boundary)
        // symmetric form of the trailing-matrix product (lower triangle only, tiles streamed by the copy engine): needs
        // 16-byte aligned tile lines (even order, aligned base) and room for at least two ring stages
        int sym_nst = 0;
        CUtensorMap tmap;
        std::memset(&tmap, 0, sizeof(tmap));
        if (!sym_off && (n & 1) == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0 && n - EB > resident_max) {
            const size_t off = sytrd_sym_ring_offset(n, grid) * sizeof(double);
            if (panel_smem_cap > off) sym_nst = (int)std::min<size_t>(kSyStagesMax, (panel_smem_cap - off) / (kSyTile * sizeof(double)));
            if (sym_nst < kSyGroups || !make_tile_map(&tmap, A, n)) sym_nst = 0;   // one ring stage per group of warps
            // single-pass column update, every CTA owns tiles, and a slice of threads holds its share of the terms of a row
            const int jjmax = n - resident_max + EB;
            if (n - 1 > grid * kSyRows || (n + kSyRH - 1) / kSyRH > kSyMaxChunks || sym_grid(std::min(jjmax, n - 1), n).T < grid ||
                sym_gather_max_terms(n, grid, jjmax) > kSySlices * kSyGatherTerms)
                sym_nst = 0;
        }
        const size_t panel_smem = sym_nst ? sytrd_sym_ring_offset(n, grid) * sizeof(double) + (size_t)sym_nst * kSyTile * sizeof(double)
                                          : sytrd_smem(n, grid);
        for (int p = 0; p < npan; ++p) {
            const int j0 = p * EB, nc = std::min(EB, n - j0), q = p & 1;
            if (q == 0 && n - j0 <= resident_max) { jres = j0; break; }
            double* V = PB + (size_t)n * EB * q;
            XTB_CUDA_CHECK(cudaMemsetAsync(V, 0, (size_t)n * EB * sizeof(double), st));
            XTB_CUDA_CHECK(cudaMemsetAsync(Wp, 0, (size_t)n * EB * sizeof(double), st));
            XTB_CUDA_CHECK(cudaMemsetAsync(bar, 0, sizeof(unsigned), st));
            SytrdArgs sa;
            sa.A = A; sa.n = n; sa.j0 = j0; sa.ncols = nc; sa.V = V; sa.W = Wp; sa.Vprev = q ? PB : nullptr;
            sa.d = dd; sa.e = ee; sa.tau = tau; sa.y = y; sa.part_n = part; sa.part_vy = part + nsm; sa.vtv = vtv; sa.wtv = wtv;
            sa.G = Gf + (size_t)(p >> 1) * NB * NB; sa.gq = EB * q; sa.bar = bar;
            sa.prof = tuning_env("XTBB200_SYTRD_PROF") ? prof : nullptr;
            sa.sym = sym_nst ? 1 : 0; sa.nst = sym_nst; sa.colbuf = sym_colbuf; sa.rowbuf = sym_rowbuf;
            void* args[] = {&sa, &tmap};
            hook(0, 0, 0);
            XTB_CUDA_CHECK(cudaLaunchCooperativeKernel((void*)k_sytrd_panel, dim3(grid), dim3(kSytrdThreads), args, panel_smem, st));
            hook(0, 1, 1);
            ++nl;
            const int j1 = j0 + nc, mt = n - j1;
            if (mt > 0) {   // A22 -= [V | W] [W | V]^T on both triangles (only full panels have a trailing matrix)
                GemmTask t = gemm_task(V + j1, n, Wp + j1, n, A + j1 + ld * j1, n, mt, mt, 2 * EB, -1.0, 1.0);
                t.ksegA = EB; t.offA2 = NB - EB * q;          // K index of W's first column relative to V
                t.ksegB = EB; t.offB2 = -(NB - EB * q);       // K index of V's first column relative to W
                t.lower = sym_nst ? 1 : 0;                     // symmetric panel kernel: nobody reads the upper triangle any more
                hook(1, 0, 0);
                gemm(st, false, true, t);
                hook(1, 1, 1);
                ++nl;
            }
        }
        if (jres >= 0) {
            XTB_CUDA_CHECK(cudaMemsetAsync(bar, 0, sizeof(unsigned), st));
            SytrdSmemArgs ra;
            ra.A = A; ra.n = n; ra.j0 = jres; ra.d = dd; ra.e = ee; ra.tau = tau; ra.cbuf = cbuf; ra.ybuf = ybuf;
            ra.Gf = Gf; ra.bar = bar; ra.clu = 0; ra.ms = 0; ra.vs = 0; ra.skip = nullptr;
            ra.lower_only = (sym_nst && jres > 0) ? 1 : 0;
            void* args[] = {&ra};
            hook(0, 0, 0);
            // Grid of the resident kernel: the grid barrier per column costs more the more CTAs take part, and a mid-size
            // matrix does not need the shared memory of all SMs: the smallest grid that holds it, but at least kResGridMin
            int gres = grid;
            {
                int gmin = 1;
                while (gmin < grid && (sytrd_smem_resident_bytes(n - jres, gmin) > res_smem_cap || (n - jres + gmin - 1) / gmin > 96)) ++gmin;   // 96: job slots of the CTA's partial sums
                gres = std::min(grid, std::max(gmin, res_grid_want > 0 ? res_grid_want : kResGridMin));
            }
            // threads: every thread owns kResRows rows of a column; fewer warps make the CTA barriers of a column step cheaper
            // measured on B200 (tools/res_grid_sweep.sh, sytrd ms): n = 552: 3.09 (1024 threads) -> 2.61 (2 rows x 512);
            // n = 900: 5.55 -> 4.70; n = 1400: 9.56 (2 x 1024) -> 9.01 (4 x 512); 8 rows per thread lose everywhere
            int rr = (n - jres > 1024) ? 4 : kResRows, tmin = 256;
            if (const char* e = tuning_env("XTBB200_RES_RR")) rr = std::atoi(e);            // tuning knobs
            if (const char* e = tuning_env("XTBB200_RES_TMIN")) tmin = std::atoi(e);
            while (rr * kSytrdThreads < n - jres) rr *= 2;
            int tres = kSytrdThreads;
            while (tres > tmin && rr * (tres / 2) >= n - jres) tres /= 2;
            void* fn = rr == 2 ? (void*)k_sytrd_smem<2> : rr == 4 ? (void*)k_sytrd_smem<4> : (void*)k_sytrd_smem<8>;
            XTB_CUDA_CHECK(cudaLaunchCooperativeKernel(fn, dim3(gres), dim3(tres), args, sytrd_smem_resident_bytes(n - jres, gres), st));
            hook(0, 1, 1);
            ++nl;
            const int jg = sym_nst ? 0 : jres;               // symmetric panel kernel: it leaves all Gram matrices to this step
            const int nrb = (n - 1 - jg + NB - 1) / NB;      // block reflectors of the resident range
            if (nrb > 0) {
                hook(1, 0, 0);
                k_gram_formv<<<(unsigned)(((size_t)n * (n - 1 - jg) + 255) / 256), 256, 0, st>>>(A, n, jg, Qg);
                k_gram_tasks<<<(nrb + 63) / 64, 64, 0, st>>>(tasks, n, jg, Qg, Gf);
                k_eg_gemm<true, false><<<dim3(4, nrb), 128, 0, st>>>(GemmTask(), tasks);
                hook(1, 1, 3);
                nl += 3;
            }
        }
        if (timed) cudaEventRecord(ev[1], st);
        hook(2, 0, 0);
        const int nl_dc0 = nl;
        // ---- 2. divide and conquer
        XTB_CUDA_CHECK(cudaMemsetAsync(Qa, 0, n2 * sizeof(double), st));
        XTB_CUDA_CHECK(cudaMemsetAsync(Qb, 0, n2 * sizeof(double), st));
        double* Qo = Qa;
        double* Qn = Qb;
        double* la = lam;
        double* lb = lam2;
        if (n == 1) {
            XTB_CUDA_CHECK(cudaMemcpyAsync(la, dd, sizeof(double), cudaMemcpyDeviceToDevice, st));
            const double one = 1.0;
            XTB_CUDA_CHECK(cudaMemcpyAsync(Qo, &one, sizeof(double), cudaMemcpyHostToDevice, st));
        } else {
            const int D = eigcore::tree_depth(n);
            {
                DcArgs a;
                std::memset(&a, 0, sizeof(a));
                a.n = n; a.depth = D; a.d = dd; a.e = ee; a.lam = la; a.Qo = Qo;
                k_dc_leaves<<<((1 << D) + 127) / 128, 128, 0, st>>>(a);
            }
            ++nl;
            for (int depth = D - 1; depth >= 0; --depth) {
                const int nodes = 1 << depth, mmax = (n + nodes - 1) / nodes;
                DcArgs a;
                a.d = dd; a.vs = 0; a.ms = 0; a.skip = nullptr;
                a.n = n; a.depth = depth; a.e = ee; a.Qo = Qo; a.Qn = Qn; a.U = U; a.Qg = Qg; a.Ug = Ug; a.lam = la; a.lam2 = lb;
                a.z = z; a.zhat = zhat; a.dk = dk; a.zk = zk; a.rc = rc; a.rs = rs; a.perm = perm; a.cols = cols; a.pos = pos;
                a.tcnt = tcnt; a.ctype = ctype; a.ra = ra; a.rb = rb; a.kcnt = kcnt; a.nrot = nrot; a.fail = fail; a.tasks = tasks;
                k_dc_sort<<<dim3((mmax + 255) / 256, nodes), 256, (size_t)mmax * 8, st>>>(a);
                {
                    const size_t need = (size_t)mmax * 24 + 16;
                    if (need <= smem_limit) k_dc_deflate<<<nodes, 256, need, st>>>(a, 1);
                    else k_dc_deflate<<<nodes, 256, 0, st>>>(a, 0);
                }
                k_dc_rot<<<dim3((mmax + 255) / 256, nodes), 256, 0, st>>>(a);
                k_dc_move<<<dim3((mmax + 255) / 256, mmax, nodes), 256, 0, st>>>(a);
                k_dc_secular<<<dim3((mmax + 7) / 8, nodes), 256, (size_t)mmax * 16, st>>>(a);
                k_dc_zhat<<<dim3((mmax + 31) / 32, nodes), 256, 0, st>>>(a);
                k_dc_vec<<<dim3((mmax + 7) / 8, nodes), 256, 0, st>>>(a);
                k_dc_tasks<<<(nodes + 127) / 128, 128, 0, st>>>(a, nodes);
                const int th = (mmax / 2 + 1 + GT - 1) / GT, tw = (mmax + GT - 1) / GT;
                k_eg_gemm<false, false><<<dim3(th * tw, 2 * nodes), 128, 0, st>>>(GemmTask(), tasks);
                nl += 9;
                std::swap(Qo, Qn);
                std::swap(la, lb);
            }
        }
        k_dc_rank<<<(n + 255) / 256, 256, (size_t)n * 8, st>>>(n, la, rank, w);
        k_dc_permute<<<(unsigned)((n2 + 255) / 256), 256, 0, st>>>(n, Qo, rank, Z);
        nl += 2;
        hook(2, 1, nl - nl_dc0);
        if (timed) cudaEventRecord(ev[2], st);
        hook(3, 0, 0);
        const int nl_bt0 = nl;
        // ---- 3. back-transformation, last block reflector first
        const int nrefl = n - 1;   // reflectors 0 .. n-2 (the last one is the identity)
        if (nrefl > 0) {
            const int nbb = (nrefl + NB - 1) / NB;
            k_bt_tfactor<<<nbb, NB * kTfParts, (NB * (NB + 1) + kTfParts * NB) * sizeof(double), st>>>(Gf, tau, n, Tf);
            ++nl;
            for (int b = nbb - 1; b >= 0; --b) {
                const int jb = b * NB, kb = std::min(NB, nrefl - jb), r0 = jb + 1, mr = n - r0;
                k_bt_formv<<<(unsigned)(((size_t)n * kb + 255) / 256), 256, 0, st>>>(A, n, jb, kb, Vb);
                // (I - V T V^T) Z = Z - (V T) (V^T Z): W = V^T Z (K split only while there are too few tiles), VT = V T, Z -= VT W
                const int tiles = ((kb + GT - 1) / GT) * ((n + GT - 1) / GT);
                const int split = std::max(1, std::min(kBtSplit, (2 * nsm + tiles - 1) / tiles));
                const int zb = zc0, ze = zc1 < 0 ? n : std::min(n, zc1), nz = ze - zb;   // columns of Z of this rank
                if (nz > 0) {
                    double* Zc = Z + (size_t)n * zb;
                    GemmTask t = gemm_task(Vb + r0, n, Zc + r0, n, Wz, NB, kb, nz, mr, 1.0, 0.0);
                    t.ksplit = split; t.csplit = (long long)NB * n;
                    gemm(st, true, false, t, split);
                    if (split > 1) k_bt_sum<<<(unsigned)(((size_t)NB * nz + 255) / 256), 256, 0, st>>>(Wz, (long long)NB * n, split, (long long)NB * nz);
                    gemm(st, false, false, gemm_task(Vb + r0, n, Tf + (size_t)b * NB * NB, NB, VT + r0, n, mr, kb, kb, 1.0, 0.0));
                    gemm(st, false, false, gemm_task(VT + r0, n, Wz, NB, Zc + r0, n, mr, nz, kb, -1.0, 1.0));
                }
                nl += split > 1 ? 5 : 4;
            }
        }
        hook(3, 1, nl - nl_bt0);
        if (timed) {
            cudaEventRecord(ev[3], st);
            cudaEventSynchronize(ev[3]);
            for (int k = 0; k < 3; ++k) cudaEventElapsedTime(&ms[k], ev[k], ev[k + 1]);
        }
        if (timed && tuning_env("XTBB200_SYTRD_PROF")) {
            long long h[16];
            cudaMemcpy(h, prof, sizeof(h), cudaMemcpyDeviceToHost);
            static const char* nm[16] = {"A:gather+scalars", "A:rows", "A:reduce", "barrier1", "B:reflector+v", "B:column dots", "B:collect",
                                         "barrier2", "A:loads issued", "A:scalars loaded", "-", "-", "-", "-", "-", "-"};
            for (int k = 0; k < 10; ++k) std::fprintf(stderr, "sytrd phase %-14s %10.3f us per column (n=%d)\n", nm[k], h[k] / 1.965e3 / n, n);
            cudaMemset(prof, 0, sizeof(h));
        }
        launches = nl;
        return nl;
    }
};

}  // namespace xtbb200
