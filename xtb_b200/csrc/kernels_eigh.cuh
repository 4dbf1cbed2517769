// Symmetric eigensolver for matrices that do not fit into shared memory (crambin, peptides, water clusters):
// blocked Householder tridiagonalisation, divide and conquer on the tridiagonal matrix, blocked back-transformation.
//
//   A = Q T Q^T      k_sytrd_panel (persistent cooperative kernel, one launch per panel of EB columns: column update,
//                    reflector, w = tau (A v - V W^T v - W V^T v) with the trailing-matrix product as warp-per-column
//                    dot products -- the HBM-bound part, 8 (n-j)^2 bytes per column) + rank-2*EB trailing update on the
//                    FP64 tensor path (k_eg_gemm, DMMA m8n8k4)
//   T = Qt L Qt^T    Cuppen divide and conquer down to 1x1 / 2x2 leaves, one batch of kernels per tree level:
//                    rank sort, deflation scan, Givens rotations, secular equation (thread per root), Loewner update
//                    vector, eigenvector update as one batched DMMA GEMM per level (scalar routines: eig_dc_core.h)
//   Z = Q Qt         block reflectors I - V T V^T, the T factors assembled from the V^T v products that the
//                    tridiagonalisation computes anyway; two DMMA GEMMs per panel
//
// Replaces dsyevd('V') inside TEigenSolver%fact_solve (src/mctc/lapack/eigensolve.F90:218-246) for large n; the
// reduction to standard form (dsygst) and the back-substitution (dtrsm) are the X^T H X and X Z products of engine.cu.
#pragma once
#include "common.cuh"
#include "eig_dc_core.h"
#include "kernels_linalg.cuh"

namespace xtbb200 {

constexpr int EB = 64;             // panel width: tridiagonalisation panel = block reflector of the back-transformation
constexpr int kSytrdThreads = 1024;
constexpr int kBtSplit = 4;        // split of the long K dimension of W = V^T C

// ---------------------------------------------------------------------------------------------------------------
// general GEMM  C = alpha op(A) op(B) + beta C  on 64x64 tiles (128 threads, 4 warps of 32x32, DMMA m8n8k4)
// ---------------------------------------------------------------------------------------------------------------
struct GemmTask {
    const double* A;
    const double* B;
    double* C;
    long long lda, ldb, ldc;
    int M, N, K;
    double alpha, beta;
    const int* gA;        // optional column gather of a non-transposed A: op(A)(r,k) = A[r + lda gA[k]]
    int ksplit;           // > 1: blockIdx.z selects a K chunk, partial products go to C + blockIdx.z * csplit
    long long csplit;
};

template <bool TA, bool TB>
__global__ void __launch_bounds__(128) k_eg_gemm(GemmTask t0, const GemmTask* __restrict__ tasks) {
    const GemmTask t = tasks ? tasks[blockIdx.y] : t0;
    const int tm = (t.M + GT - 1) / GT, tn = (t.N + GT - 1) / GT;
    if ((int)blockIdx.x >= tm * tn) return;
    const int r0 = (blockIdx.x % tm) * GT, c0 = (blockIdx.x / tm) * GT;
    int kb = 0, ke = t.K;
    double* C = t.C;
    if (t.ksplit > 1) {
        const int ch = (((t.K + t.ksplit - 1) / t.ksplit + GK - 1) / GK) * GK;
        kb = min(t.K, (int)blockIdx.z * ch);
        ke = min(t.K, kb + ch);
        C += (long long)blockIdx.z * t.csplit;
    }
    const double* __restrict__ A = t.A;
    const double* __restrict__ B = t.B;
    const int* __restrict__ gA = t.gA;
    const long long lda = t.lda, ldb = t.ldb, ldc = t.ldc;
    const int M = t.M, N = t.N;
    __shared__ double As[GK][GLD], Bs[GK][GLD];
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
                const int r = tt % GT, kk = tt / GT, gr = r0 + r, gk = k0 + kk;
                ra[u] = (gr < M && gk < ke) ? A[gr + lda * (gA ? gA[gk] : gk)] : 0.0;
            } else {
                const int kk = tt % GK, r = tt / GK, gr = r0 + r, gk = k0 + kk;
                ra[u] = (gr < M && gk < ke) ? A[gk + lda * gr] : 0.0;
            }
            if (!TB) {
                const int kk = tt % GK, c = tt / GK, gc = c0 + c, gk = k0 + kk;
                rb[u] = (gc < N && gk < ke) ? B[gk + ldb * gc] : 0.0;
            } else {
                const int c = tt % GT, kk = tt / GT, gc = c0 + c, gk = k0 + kk;
                rb[u] = (gc < N && gk < ke) ? B[gc + ldb * gk] : 0.0;
            }
        }
    };
    auto stage = [&]() {
#pragma unroll
        for (int u = 0; u < GT * GK / 128; ++u) {
            const int tt = tid + 128 * u;
            if (!TA) As[tt / GT][tt % GT] = ra[u]; else As[tt % GK][tt / GK] = ra[u];
            if (!TB) Bs[tt % GK][tt / GK] = rb[u]; else Bs[tt / GT][tt % GT] = rb[u];
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
    const double alpha = t.alpha, beta = t.beta;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gr = r0 + wr + 8 * i + (lane >> 2);
            const int gc = c0 + wc + 8 * j + 2 * (lane & 3);
            if (gr < M) {
#pragma unroll
                for (int h = 0; h < 2; ++h)
                    if (gc + h < N) {
                        double* p = C + gr + ldc * (gc + h);
                        const double v = alpha * acc[i][j][h];
                        *p = (beta == 0.0) ? v : v + beta * *p;
                    }
            }
        }
}

// ---------------------------------------------------------------------------------------------------------------
// tridiagonalisation
// ---------------------------------------------------------------------------------------------------------------
struct SytrdArgs {
    double* A;          // n x n, column-major, both triangles valid on entry of every panel
    int n, j0, ncols;   // panel columns [j0, j0 + ncols)
    double* V;          // n x EB reflectors of the panel (explicit, zero above the unit entry)
    double* W;          // n x EB
    double* d;          // diagonal of T
    double* e;          // sub-diagonal of T
    double* tau;
    double* y;          // n: trailing-matrix product A v
    double* part_n;     // [grid] partial sums of squares
    double* part_vy;    // [grid] partial v^T y
    double* vtv;        // [EB]  V^T v
    double* wtv;        // [EB]  W^T v
    double* G;          // EB x EB of this panel: G[k + EB i] = v_k^T v_i (k < i), input of the T factor
    unsigned* bar;      // grid barrier counter (zeroed before the launch)
};

__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned nblocks, unsigned& target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += nblocks;
        __threadfence();
        atomicAdd(bar, 1u);
        unsigned v;
        do {
            asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
        } while (v < target);
        __threadfence();
    }
    __syncthreads();
}

// dynamic shared memory: sv[n] | sh_part[(n / grid + 4) * 8]
__global__ void __launch_bounds__(kSytrdThreads, 1) k_sytrd_panel(SytrdArgs a) {
    extern __shared__ double sm_dyn[];
    __shared__ double sh_vtv[EB], sh_wtv[EB], sh_vrow[EB], sh_wrow[EB], red[32], sh_sc[4];
    const int n = a.n, G = gridDim.x, cta = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
    const long long ld = n;
    double* __restrict__ A = a.A;
    double* __restrict__ V = a.V;
    double* __restrict__ W = a.W;
    double* sv = sm_dyn;
    double* sh_part = sm_dyn + n;
    unsigned target = 0;
    double taup = 0.0;   // tau of the previous column
    for (int i = 0;; ++i) {
        const int j = a.j0 + i;
        const bool fin = i > 0, upd = i < a.ncols;
        const int ip = i - 1, jp = j - 1;
        double alpha2 = 0.0;
        // ---------------- phase A: finish w of column j-1, update column j, partial norm
        if (fin) {
            for (int k = tid; k < ip; k += nthr) {
                sh_vtv[k] = __ldcg(a.vtv + k); sh_wtv[k] = __ldcg(a.wtv + k);
                sh_vrow[k] = __ldcg(V + j + ld * k); sh_wrow[k] = __ldcg(W + j + ld * k);
            }
            __syncthreads();
            if (warp == 0) {
                double vy = 0.0, vw = 0.0, s = 0.0;
                for (int c = lane; c < G; c += 32) vy += __ldcg(a.part_vy + c);
                for (int k = lane; k < ip; k += 32) { vw += sh_vtv[k] * sh_wtv[k]; s += sh_vrow[k] * sh_wtv[k] + sh_wrow[k] * sh_vtv[k]; }
                vy = warp_sum(vy); vw = warp_sum(vw); s = warp_sum(s);
                if (lane == 0) {
                    const double al2 = -0.5 * taup * taup * (vy - 2.0 * vw);
                    sh_sc[0] = al2;
                    sh_vrow[ip] = 1.0;                                        // V[j, ip]: the unit entry of v
                    sh_wrow[ip] = taup * (__ldcg(a.y + j) - s) + al2;          // W[j, ip]
                }
            }
            __syncthreads();
            alpha2 = sh_sc[0];
        }
        double sq = 0.0;
        for (long long r = (long long)j + (long long)cta * nthr + tid; r < n; r += (long long)G * nthr) {
            double s1 = 0.0, s2 = 0.0;
            if (fin) {
                for (int k = 0; k < ip; ++k) {
                    const double vk = __ldcg(V + r + ld * k), wk = __ldcg(W + r + ld * k);
                    s1 += vk * sh_wtv[k] + wk * sh_vtv[k];
                    s2 += vk * sh_wrow[k] + wk * sh_vrow[k];
                }
                const double vr = __ldcg(V + r + ld * ip);
                const double wr = taup * (__ldcg(a.y + r) - s1) + alpha2 * vr;
                W[r + ld * ip] = wr;
                if (r > j) A[r + ld * jp] = vr;                               // reflector storage below the sub-diagonal
                s2 += vr * sh_wrow[ip] + wr * sh_vrow[ip];
            }
            if (upd) {
                const double av = __ldcg(A + r + ld * j) - s2;
                A[r + ld * j] = av;
                if (r == j) a.d[j] = av;
                else if (r >= j + 2) sq += av * av;
            }
        }
        if (!upd || j == n - 1) break;
        sq = block_sum(sq, red);
        if (tid == 0) a.part_n[cta] = sq;
        grid_barrier(a.bar, G, target);
        // ---------------- phase B: reflector, trailing-matrix product, V^T v, W^T v
        const int m = n - j - 1;
        if (warp == 0) {
            double s = 0.0;
            for (int c = lane; c < G; c += 32) s += __ldcg(a.part_n + c);
            s = warp_sum(s);
            if (lane == 0) {
                const double alpha = __ldcg(A + (j + 1) + ld * j);
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
        __syncthreads();
        const double tau = sh_sc[1], scal = sh_sc[2];
        for (int t = tid; t < m; t += nthr) sv[t] = (t == 0) ? 1.0 : __ldcg(A + (j + 1 + t) + ld * j) * scal;
        __syncthreads();
        for (long long t = (long long)cta * nthr + tid; t < m; t += (long long)G * nthr) V[(j + 1 + t) + ld * i] = sv[t];
        // jobs of this CTA: its share of the trailing columns, plus V / W columns dealt round-robin
        const int cb = j + 1 + (int)(((long long)m * cta) / G), ce = j + 1 + (int)(((long long)m * (cta + 1)) / G);
        const int ncA = ce - cb;
        int nex = 0;
        for (int k = cta; k < i; k += G) ++nex;
        const int nloc = ncA + 2 * nex;
        int Q = 1;
        if (nloc > 0) {
            Q = (4 * nwarp + nloc - 1) / nloc;
            Q = max(1, min(8, min(Q, m / 128)));
        }
        const int njobs = nloc * Q;
        for (int job = warp; job < njobs; job += nwarp) {
            const int lc = job / Q, q = job - lc * Q;
            const double* col;
            if (lc < ncA) col = A + (j + 1) + ld * (cb + lc);
            else {
                const int x = lc - ncA, k = cta + (x >> 1) * G;
                col = ((x & 1) ? W : V) + (j + 1) + ld * k;
            }
            const int tb = (int)(((long long)m * q) / Q), te = (int)(((long long)m * (q + 1)) / Q);
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            int t = tb + lane;
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
        double vyl = 0.0;
        for (int lc = tid; lc < nloc; lc += nthr) {
            double s = 0.0;
            for (int q = 0; q < Q; ++q) s += sh_part[lc * Q + q];
            if (lc < ncA) {
                const int c = cb + lc;
                a.y[c] = s;
                vyl += sv[c - (j + 1)] * s;
            } else {
                const int x = lc - ncA, k = cta + (x >> 1) * G;
                if (x & 1) a.wtv[k] = s;
                else { a.vtv[k] = s; a.G[k + EB * i] = s; }
            }
        }
        vyl = block_sum(vyl, red);
        if (tid == 0) a.part_vy[cta] = vyl;
        taup = tau;
        grid_barrier(a.bar, G, target);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// back-transformation  Z <- (I - V T V^T) Z  per panel
// ---------------------------------------------------------------------------------------------------------------
// explicit V of panel [jb, jb+kb): Vb[r + n k], rows absolute
__global__ void k_bt_formv(const double* __restrict__ A, int n, int jb, int kb, double* __restrict__ Vb) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)n * kb) return;
    const int r = (int)(t % n), k = (int)(t / n), piv = jb + k + 1;
    Vb[t] = r < piv ? 0.0 : (r == piv ? 1.0 : A[r + (long long)n * (jb + k)]);
}
// T factors of all panels (one CTA per panel): T(0:i,i) = -tau_i T(0:i,0:i) G(0:i,i), T(i,i) = tau_i
__global__ void __launch_bounds__(EB) k_bt_tfactor(const double* __restrict__ Gall, const double* __restrict__ tau, int n,
                                                   double* __restrict__ Tall) {
    __shared__ double T[EB][EB + 1];
    const int p = blockIdx.x, jb = p * EB, kb = min(EB, n - 1 - jb), r = threadIdx.x;
    const double* G = Gall + (long long)p * EB * EB;
    for (int c = 0; c < EB; ++c) T[r][c] = 0.0;
    __syncthreads();
    for (int i = 0; i < kb; ++i) {
        const double ti = tau[jb + i];
        if (r < i) {
            double s = 0.0;
            for (int k = r; k < i; ++k) s += T[r][k] * G[k + EB * i];
            T[r][i] = -ti * s;
        } else if (r == i) T[r][i] = ti;
        __syncthreads();
    }
    double* To = Tall + (long long)p * EB * EB;
    for (int c = 0; c < EB; ++c) To[r + EB * c] = T[r][c];
}
// W2 = T * sum_z Wz  (EB x n, leading dimension EB); thread per output element
__global__ void k_bt_tw(const double* __restrict__ T, const double* __restrict__ Wz, long long zstride, int nz, int kb, int n,
                        double* __restrict__ W2) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)EB * n) return;
    const int i = (int)(t % EB);
    const long long c = t / EB;
    double s = 0.0;
    if (i < kb)
        for (int k = i; k < kb; ++k) {
            double w = 0.0;
            for (int z = 0; z < nz; ++z) w += Wz[(long long)z * zstride + k + EB * c];
            s += T[i + EB * k] * w;
        }
    W2[t] = s;
}

// ---------------------------------------------------------------------------------------------------------------
// divide and conquer on the tridiagonal matrix; blockIdx.y (or .z) = node of the current tree level
// ---------------------------------------------------------------------------------------------------------------
struct DcArgs {
    int n, depth;
    const double* e;      // sub-diagonal of T (coupling of the torn blocks)
    double *Qo, *Qn, *U;  // eigenvector blocks of the children / of the merged nodes, delta / update matrix
    double *lam, *lam2, *z, *zhat, *dk, *zk, *rc, *rs;
    int *perm, *cols, *ra, *rb, *kcnt, *nrot, *fail;
    GemmTask* tasks;
};
struct DcNode { int lo, mid, hi; };
__device__ __forceinline__ DcNode dc_node(const DcArgs& a, int t) {
    DcNode nd;
    nd.lo = eigcore::tree_bound(a.n, a.depth, t);
    nd.hi = eigcore::tree_bound(a.n, a.depth, t + 1);
    nd.mid = eigcore::tree_bound(a.n, a.depth + 1, 2 * t + 1);
    return nd;
}

// leaves of the tearing tree: 1x1 and 2x2 blocks of the torn matrix; thread per leaf
__global__ void k_dc_leaves(int n, int D, const double* __restrict__ d, const double* __restrict__ e, double* __restrict__ lam,
                            double* __restrict__ Q) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (1 << D)) return;
    const int lo = eigcore::tree_bound(n, D, t), hi = eigcore::tree_bound(n, D, t + 1);
    const long long ld = n;
    double a = d[lo] - (lo > 0 ? fabs(e[lo - 1]) : 0.0);
    if (hi - lo == 1) {
        a -= (hi < n ? fabs(e[hi - 1]) : 0.0);
        lam[lo] = a;
        Q[lo + ld * lo] = 1.0;
    } else {
        const double c = d[lo + 1] - (hi < n ? fabs(e[hi - 1]) : 0.0), b = e[lo];
        double rt1, rt2, cs, sn;
        eigcore::sym2x2(a, b, c, rt1, rt2, cs, sn);
        lam[lo] = rt1; lam[lo + 1] = rt2;
        Q[lo + ld * lo] = cs; Q[lo + 1 + ld * lo] = sn;
        Q[lo + ld * (lo + 1)] = -sn; Q[lo + 1 + ld * (lo + 1)] = cs;
    }
}

// update vector z (last row of the left block, first row of the right block) and ascending order of the eigenvalues
__global__ void __launch_bounds__(256) k_dc_sort(DcArgs a) {
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

// deflation: one CTA per node, eigenvalues / z / order staged in shared memory, sequential scan by thread 0
__global__ void __launch_bounds__(256) k_dc_deflate(DcArgs a) {
    extern __shared__ double sm_dyn[];
    const DcNode nd = dc_node(a, blockIdx.x);
    const int m = nd.hi - nd.lo, lo = nd.lo;
    double* sd = sm_dyn;
    double* sz = sm_dyn + m;
    int* sp = (int*)(sm_dyn + 2 * m);
    __shared__ int sh_k;
    for (int i = threadIdx.x; i < m; i += blockDim.x) { sd[i] = a.lam[lo + i]; sz[i] = a.z[lo + i]; sp[i] = a.perm[lo + i]; }
    __syncthreads();
    const double rho = 2.0 * fabs(a.e[nd.mid - 1]);
    if (threadIdx.x == 0) {
        int nrot = 0;
        sh_k = eigcore::deflate_scan(m, sp, sd, sz, rho, a.cols + lo, a.ra + lo, a.rb + lo, a.rc + lo, a.rs + lo, &nrot);
        a.kcnt[blockIdx.x] = sh_k;
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

// secular equation: thread per root; delta(i,j) = d_i - lambda_j into U
__global__ void __launch_bounds__(64) k_dc_secular(DcArgs a) {
    extern __shared__ double sm_dyn[];
    const int k = a.kcnt[blockIdx.y];
    if ((int)(blockIdx.x * blockDim.x) >= k) return;
    const DcNode nd = dc_node(a, blockIdx.y);
    double* sd = sm_dyn;
    double* sz = sm_dyn + k;
    for (int i = threadIdx.x; i < k; i += blockDim.x) { sd[i] = a.dk[nd.lo + i]; sz[i] = a.zk[nd.lo + i]; }
    __syncthreads();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= k) return;
    const double rho = 2.0 * fabs(a.e[nd.mid - 1]);
    double lam;
    const int st = eigcore::secular_root(k, j, sd, sz, rho, a.U + nd.lo + (long long)a.n * (nd.lo + j), 1, &lam);
    a.lam2[nd.lo + j] = lam;
    if (st) atomicAdd(a.fail, 1);
}

// Loewner formula for the update vector that makes the computed roots exact; thread per entry
__global__ void __launch_bounds__(128) k_dc_zhat(DcArgs a) {
    const int k = a.kcnt[blockIdx.y];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k) return;
    const DcNode nd = dc_node(a, blockIdx.y);
    const long long ld = a.n;
    const double* U = a.U + (nd.lo + i) + ld * nd.lo;
    const double* dk = a.dk + nd.lo;
    const double di = dk[i];
    double w = U[ld * i];
    for (int j = 0; j < k; ++j)
        if (j != i) w *= U[ld * j] / (di - dk[j]);
    a.zhat[nd.lo + i] = copysign(sqrt(fabs(w)), a.zk[nd.lo + i]);
}

// eigenvectors of the rank-one update: U(i,j) = zhat_i / delta(i,j), normalised; warp per column
__global__ void __launch_bounds__(256) k_dc_vec(DcArgs a) {
    const int k = a.kcnt[blockIdx.y];
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (j >= k) return;
    const DcNode nd = dc_node(a, blockIdx.y);
    double* U = a.U + nd.lo + (long long)a.n * (nd.lo + j);
    const double* zh = a.zhat + nd.lo;
    double s = 0.0;
    for (int i = lane; i < k; i += 32) { const double u = zh[i] / U[i]; s += u * u; }
    s = warp_sum(s);
    const double sc = 1.0 / sqrt(s);
    for (int i = lane; i < k; i += 32) U[i] = zh[i] / U[i] * sc;
}

// GEMM tasks of the level: Qn[lo:hi, lo:lo+k) = Qo[lo:hi, cols[0:k)] U(0:k, 0:k); thread per node
__global__ void k_dc_tasks(DcArgs a, int nodes) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nodes) return;
    const DcNode nd = dc_node(a, t);
    const long long ld = a.n, off = nd.lo + ld * nd.lo;
    GemmTask g;
    g.A = a.Qo + off; g.B = a.U + off; g.C = a.Qn + off;
    g.lda = ld; g.ldb = ld; g.ldc = ld;
    g.M = nd.hi - nd.lo; g.N = a.kcnt[t]; g.K = a.kcnt[t];
    g.alpha = 1.0; g.beta = 0.0;
    g.gA = a.cols + nd.lo;
    g.ksplit = 1; g.csplit = 0;
    a.tasks[t] = g;
}

// deflated eigenvectors are carried over unchanged behind the updated ones; grid (row chunks, slot, node)
__global__ void __launch_bounds__(256) k_dc_copy(DcArgs a) {
    const int k = a.kcnt[blockIdx.z];
    const DcNode nd = dc_node(a, blockIdx.z);
    const int m = nd.hi - nd.lo;
    const int p = k + blockIdx.y;
    if (p >= m) return;
    const int row = nd.lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= nd.hi) return;
    const long long ld = a.n;
    a.Qn[row + ld * (nd.lo + p)] = a.Qo[row + ld * (nd.lo + a.cols[nd.lo + p])];
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
// host driver
// ---------------------------------------------------------------------------------------------------------------
struct EighWork {
    int nmax = 0, nsm = 0, grid = 0;
    double *Qa = nullptr, *Qb = nullptr, *U = nullptr;
    double *V = nullptr, *W = nullptr, *Vb = nullptr, *Wz = nullptr, *W2 = nullptr, *Tf = nullptr, *Gf = nullptr;
    double *dd = nullptr, *ee = nullptr, *tau = nullptr, *y = nullptr, *part = nullptr, *vtv = nullptr, *wtv = nullptr;
    double *lam = nullptr, *lam2 = nullptr, *z = nullptr, *zhat = nullptr, *dk = nullptr, *zk = nullptr, *rc = nullptr, *rs = nullptr;
    int *perm = nullptr, *cols = nullptr, *ra = nullptr, *rb = nullptr, *kcnt = nullptr, *nrot = nullptr, *fail = nullptr, *rank = nullptr;
    unsigned* bar = nullptr;
    GemmTask* tasks = nullptr;
    std::vector<void*> allocs;
    size_t bytes = 0;
    float ms[3] = {0, 0, 0};   // tridiagonalisation, divide and conquer, back-transformation of the last timed call
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    int launches = 0;

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
    static size_t sytrd_smem(int n, int grid) { return ((size_t)n + (size_t)(n / grid + 4) * 8) * sizeof(double); }
    // largest order the shared-memory staging of the panel kernel and of the deflation scan allows
    static int max_order() { return 11000; }

    void init(int nmax_, int device) {
        nmax = nmax_;
        if (nmax > max_order()) throw std::runtime_error("xtb_b200: eigensolver supports matrices up to order 11000");
        cudaDeviceProp prop;
        XTB_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
        nsm = prop.multiProcessorCount;
        if (!prop.cooperativeLaunch) throw std::runtime_error("xtb_b200: device lacks cooperative launch");
        const size_t n = nmax, n2 = n * n;
        Qa = alloc<double>(n2); Qb = alloc<double>(n2); U = alloc<double>(n2);
        V = alloc<double>(n * EB); W = alloc<double>(n * EB); Vb = alloc<double>(n * EB);
        Wz = alloc<double>((size_t)kBtSplit * EB * n); W2 = alloc<double>((size_t)EB * n);
        const size_t npan = (n + EB - 1) / EB;
        Tf = alloc<double>(npan * EB * EB); Gf = alloc<double>(npan * EB * EB);
        dd = alloc<double>(n); ee = alloc<double>(n); tau = alloc<double>(n); y = alloc<double>(n);
        part = alloc<double>(2 * (size_t)nsm); vtv = alloc<double>(EB); wtv = alloc<double>(EB);
        lam = alloc<double>(n); lam2 = alloc<double>(n); z = alloc<double>(n); zhat = alloc<double>(n);
        dk = alloc<double>(n); zk = alloc<double>(n); rc = alloc<double>(n); rs = alloc<double>(n);
        perm = alloc<int>(n); cols = alloc<int>(n); ra = alloc<int>(n); rb = alloc<int>(n);
        kcnt = alloc<int>(n); nrot = alloc<int>(n); fail = alloc<int>(1); rank = alloc<int>(n);
        bar = alloc<unsigned>(1);
        tasks = alloc<GemmTask>(n);
        grid = nsm;
        XTB_CUDA_CHECK(cudaFuncSetAttribute(k_sytrd_panel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sytrd_smem(nmax, grid)));
        XTB_CUDA_CHECK(cudaFuncSetAttribute(k_dc_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(n * 8)));
        XTB_CUDA_CHECK(cudaFuncSetAttribute(k_dc_deflate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(n * 20 + 16)));
        XTB_CUDA_CHECK(cudaFuncSetAttribute(k_dc_secular, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(n * 16)));
        XTB_CUDA_CHECK(cudaFuncSetAttribute(k_dc_rank, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(n * 8)));
        int occ = 0;
        XTB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_sytrd_panel, kSytrdThreads, sytrd_smem(nmax, grid)));
        if (occ < 1) throw std::runtime_error("xtb_b200: tridiagonalisation kernel does not fit on an SM");
        for (auto& e : ev) cudaEventCreate(&e);
    }
    void release() {
        for (void* p : allocs) cudaFree(p);
        allocs.clear();
        for (auto& e : ev) if (e) { cudaEventDestroy(e); e = nullptr; }
    }

    static void gemm(cudaStream_t st, bool ta, bool tb, const GemmTask& t, int nz = 1) {
        const int tm = (t.M + GT - 1) / GT, tn = (t.N + GT - 1) / GT;
        if (tm * tn == 0) return;
        const dim3 g(tm * tn, 1, nz);
        if (!ta && !tb) k_eg_gemm<false, false><<<g, 128, 0, st>>>(t, nullptr);
        else if (ta && !tb) k_eg_gemm<true, false><<<g, 128, 0, st>>>(t, nullptr);
        else if (!ta && tb) k_eg_gemm<false, true><<<g, 128, 0, st>>>(t, nullptr);
        else k_eg_gemm<true, true><<<g, 128, 0, st>>>(t, nullptr);
    }

    // A (n x n, leading dimension n, both triangles; destroyed) -> eigenvalues w[n] ascending, eigenvectors Z (n x n).
    // *fail (device int) is incremented when a secular equation did not converge.  Returns the number of launches.
    int solve(cudaStream_t st, double* A, int n, double* w, double* Z, bool timed) {
        int nl = 0;
        const long long ld = n;
        const size_t n2 = (size_t)n * n;
        if (timed) cudaEventRecord(ev[0], st);
        // ---- 1. tridiagonalisation
        const int npan = (n + EB - 1) / EB;
        XTB_CUDA_CHECK(cudaMemsetAsync(Gf, 0, (size_t)npan * EB * EB * sizeof(double), st));
        for (int p = 0; p < npan; ++p) {
            const int j0 = p * EB, nc = std::min(EB, n - j0);
            XTB_CUDA_CHECK(cudaMemsetAsync(V, 0, (size_t)n * EB * sizeof(double), st));
            XTB_CUDA_CHECK(cudaMemsetAsync(W, 0, (size_t)n * EB * sizeof(double), st));
            XTB_CUDA_CHECK(cudaMemsetAsync(bar, 0, sizeof(unsigned), st));
            SytrdArgs sa;
            sa.A = A; sa.n = n; sa.j0 = j0; sa.ncols = nc; sa.V = V; sa.W = W; sa.d = dd; sa.e = ee; sa.tau = tau; sa.y = y;
            sa.part_n = part; sa.part_vy = part + nsm; sa.vtv = vtv; sa.wtv = wtv; sa.G = Gf + (size_t)p * EB * EB; sa.bar = bar;
            void* args[] = {&sa};
            XTB_CUDA_CHECK(cudaLaunchCooperativeKernel((void*)k_sytrd_panel, dim3(grid), dim3(kSytrdThreads), args,
                                                       sytrd_smem(n, grid), st));
            ++nl;
            const int j1 = j0 + nc, mt = n - j1;
            if (mt > 0) {   // A22 -= V W^T + W V^T on both triangles
                GemmTask t;
                t.C = A + j1 + ld * j1; t.ldc = ld; t.M = mt; t.N = mt; t.K = nc; t.alpha = -1.0; t.beta = 1.0;
                t.gA = nullptr; t.ksplit = 1; t.csplit = 0; t.lda = ld; t.ldb = ld;
                t.A = V + j1; t.B = W + j1;
                gemm(st, false, true, t);
                t.A = W + j1; t.B = V + j1;
                gemm(st, false, true, t);
                nl += 2;
            }
        }
        if (timed) cudaEventRecord(ev[1], st);
        // ---- 2. divide and conquer
        XTB_CUDA_CHECK(cudaMemsetAsync(Qa, 0, n2 * sizeof(double), st));
        XTB_CUDA_CHECK(cudaMemsetAsync(Qb, 0, n2 * sizeof(double), st));
        double* Qo = Qa;
        double* Qn = Qb;
        if (n == 1) {
            XTB_CUDA_CHECK(cudaMemcpyAsync(lam, dd, sizeof(double), cudaMemcpyDeviceToDevice, st));
            const double one = 1.0;
            XTB_CUDA_CHECK(cudaMemcpyAsync(Qo, &one, sizeof(double), cudaMemcpyHostToDevice, st));
        } else {
            const int D = eigcore::tree_depth(n);
            k_dc_leaves<<<((1 << D) + 127) / 128, 128, 0, st>>>(n, D, dd, ee, lam, Qo);
            ++nl;
            for (int depth = D - 1; depth >= 0; --depth) {
                const int nodes = 1 << depth, mmax = (n + nodes - 1) / nodes;
                DcArgs a;
                a.n = n; a.depth = depth; a.e = ee; a.Qo = Qo; a.Qn = Qn; a.U = U; a.lam = lam; a.lam2 = lam2; a.z = z; a.zhat = zhat;
                a.dk = dk; a.zk = zk; a.rc = rc; a.rs = rs; a.perm = perm; a.cols = cols; a.ra = ra; a.rb = rb; a.kcnt = kcnt;
                a.nrot = nrot; a.fail = fail; a.tasks = tasks;
                k_dc_sort<<<dim3((mmax + 255) / 256, nodes), 256, (size_t)mmax * 8, st>>>(a);
                k_dc_deflate<<<nodes, 256, (size_t)mmax * 20 + 16, st>>>(a);
                k_dc_rot<<<dim3((mmax + 255) / 256, nodes), 256, 0, st>>>(a);
                k_dc_secular<<<dim3((mmax + 63) / 64, nodes), 64, (size_t)mmax * 16, st>>>(a);
                k_dc_zhat<<<dim3((mmax + 127) / 128, nodes), 128, 0, st>>>(a);
                k_dc_vec<<<dim3((mmax + 7) / 8, nodes), 256, 0, st>>>(a);
                k_dc_tasks<<<(nodes + 127) / 128, 128, 0, st>>>(a, nodes);
                const int tl = (mmax + GT - 1) / GT;
                k_eg_gemm<false, false><<<dim3(tl * tl, nodes), 128, 0, st>>>(GemmTask(), tasks);
                k_dc_copy<<<dim3((mmax + 255) / 256, mmax, nodes), 256, 0, st>>>(a);
                nl += 9;
                std::swap(Qo, Qn);
                std::swap(lam, lam2);
            }
        }
        k_dc_rank<<<(n + 255) / 256, 256, (size_t)n * 8, st>>>(n, lam, rank, w);
        k_dc_permute<<<(unsigned)((n2 + 255) / 256), 256, 0, st>>>(n, Qo, rank, Z);
        nl += 2;
        if (timed) cudaEventRecord(ev[2], st);
        // ---- 3. back-transformation, last panel first
        const int nrefl = n - 1;   // reflectors 0 .. n-2 (the last one is the identity)
        if (nrefl > 0) {
            const int nbp = (nrefl + EB - 1) / EB;
            k_bt_tfactor<<<nbp, EB, 0, st>>>(Gf, tau, n, Tf);
            ++nl;
            for (int p = nbp - 1; p >= 0; --p) {
                const int jb = p * EB, kb = std::min(EB, nrefl - jb), r0 = jb + 1, mr = n - r0;
                k_bt_formv<<<(unsigned)(((size_t)n * kb + 255) / 256), 256, 0, st>>>(A, n, jb, kb, Vb);
                GemmTask t;
                t.A = Vb + r0; t.lda = ld; t.B = Z + r0; t.ldb = ld; t.C = Wz; t.ldc = EB; t.M = kb; t.N = n; t.K = mr;
                t.alpha = 1.0; t.beta = 0.0; t.gA = nullptr; t.ksplit = kBtSplit; t.csplit = (long long)EB * n;
                gemm(st, true, false, t, kBtSplit);
                k_bt_tw<<<(unsigned)(((size_t)EB * n + 255) / 256), 256, 0, st>>>(Tf + (size_t)p * EB * EB, Wz, (long long)EB * n,
                                                                                    kBtSplit, kb, n, W2);
                GemmTask u;
                u.A = Vb + r0; u.lda = ld; u.B = W2; u.ldb = EB; u.C = Z + r0; u.ldc = ld; u.M = mr; u.N = n; u.K = kb;
                u.alpha = -1.0; u.beta = 1.0; u.gA = nullptr; u.ksplit = 1; u.csplit = 0;
                gemm(st, false, false, u);
                nl += 4;
            }
        }
        if (timed) {
            cudaEventRecord(ev[3], st);
            cudaEventSynchronize(ev[3]);
            for (int k = 0; k < 3; ++k) cudaEventElapsedTime(&ms[k], ev[k], ev[k + 1]);
        }
        launches = nl;
        return nl;
    }
};

}  // namespace xtbb200
