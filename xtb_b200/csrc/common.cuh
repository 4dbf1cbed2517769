// Common device definitions for the B200 GFN2-xTB engine.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdlib>

#include "system.h"

namespace xtbb200 {

#define XTB_CUDA_CHECK(x)                                                                              \
    do {                                                                                               \
        cudaError_t err__ = (x);                                                                       \
        if (err__ != cudaSuccess)                                                                      \
            throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(err__) + " at " + \
                                     __FILE__ + ":" + std::to_string(__LINE__));                       \
    } while (0)

// Diagnostic / tuning switches ("Diagnostic switches") are honoured only when the master switch
// XTBB200_TUNING=1 is set, so that a stray variable in a user's shell cannot change the solver or its thresholds.
inline const char* tuning_env(const char* name) {
    static const bool on = [] { const char* e = std::getenv("XTBB200_TUNING"); return e && e[0] == '1'; }();
    return on ? std::getenv(name) : nullptr;
}

// per-molecule SCC state, lives on the device, mirrored to the host for convergence polling
struct MolState {
    int iter;        // iterations done
    int converged;   // econverged && qconverged seen once
    int lastdiag;
    int done;        // SCC loop left
    int fail;        // error code (0 ok)
    int kocc;        // number of leading orbitals with non-zero occupation
    int ndp, nqp;    // screening list lengths
    int sweeps;      // Jacobi sweeps of the last diagonalisation
    int sweeps_total;
    int zvalid;      // Zm holds orthonormal eigenvectors of an earlier diagonalisation (Jacobi warm start)
    int pad0;
    double eel, eold, ees, eaes, epol, ed, egap, ga, gb, efa, efb, rmsq;
    double ep, embd, eh0;
    double epcem;    // interaction energy of the shell charges with the external point charges
    double energy, gnorm;
    double dipole[3];
};

// device pointer bundle
constexpr int kNCnt = 12;
// 0 shell pairs of build_SDQH0 inside r^2 <= 2000, 1 their surviving primitive pairs, 2 off-site shell pairs of
// build_dSDQH0 inside r^2 <= 1600, 3 their surviving primitive pairs, 4 pairs of the gfn / covalent CN, 5 pairs of the
// erf / D4 CN, 6 repulsion pairs, 7 non-zero reference pairs of the D4 self-consistent matrix, 8 D4 gradient pairs,
// 9 ATM triples, 10 bond-order pairs, 11 Born-radius pairs evaluated with the preceding pair's 1/r (kernels_solv.cuh)
struct Dev {
    const MolDesc* mol;
    MolState* st;
    const AtomPar* atoms;
    const ShellPar* shells;
    const RefPar* refs;
    const int* ao2sh;
    const int* ao2at;
    const int* leader;   // per molecule: first molecule of the batch with the same atoms in the same order (itself if none before)
    const double* c6ref;
    int ntype;
    int nmol;
    int maxhist;
    // split of one large system over GPUs: this process is `rank` of `nranks` (1: not split)
    int srank, snranks;
    // options
    double intcut, neglect, scfconv, qconv, etemp, broydamp;
    int maxiter;
    int jb_inner;     // inner sweeps per Gram block of the block Jacobi solver
    double jb_tol;    // relative orthogonality threshold of the block Jacobi solver (0: 8 sqrt(n) eps)
    double jb_qtol;   // quadratic stop of the block Jacobi solver: skip the verification sweep after a sweep whose largest
                      // relative off-diagonal was below this (0: always verify)
    // diagnostic: counters of the discrete decisions (cutoffs, screening) at the places where the kernels take them,
    // kNCnt per molecule; nullptr unless xtbb200_setDecisionCounting switched them on
    unsigned long long* cnt;
    // diagnostic / verbose printout: per SCC iteration of molecule 0 {E_el, E_el - E_old, rms dq, gap / eV, omega} (the columns
    // of the reference's iteration table, src/scc_core.f90:621-622); nullptr unless Batch::set_iteration_log switched it on
    double* iterlog;
    // per atom
    double *xyz, *cn, *cn_cov, *cn_d4, *cn_erf, *radcn;
    double *q, *q_out, *dipm, *dipm_out, *qp, *qp_out;
    double *vs, *vd, *vq, *ashift;
    double *grad, *dhdcn, *w_aes, *dEdcn, *e_atom;   // e_atom: [4] per atom partial energies (rep, aes, atm, spare)
    double *zw;       // D4 gradient weights per atom: [5][7]
    // per shell
    double *selfE, *qsh, *qsh_out, *sshift, *jq;
    // per AO
    double *emo, *focc, *focca, *foccb, *aop;   // aop: [12] per AO partials
    double *nrm;      // Jacobi column norms
    int *rank;        // Jacobi eigenvalue ranks
    // per ref
    double *gw, *zeta, *dzeta, *dvec, *edref;
    // matrices (nao^2)
    double *S, *D, *Q, *H0, *H0n, *H, *X, *T, *Zm, *C, *P, *W;
    uint8_t* mask;
    long long tot_mat;
    // nat^2
    double *gab3, *gab5, *c6, *dc6, *c60, *dc60, *wbo;
    // nsh^2 / ndim^2 / eeq
    double *jmat, *dispmat, *eeqA;
    const long long* eeq0;
    // broyden
    double *bdf, *bu, *ba, *bbeta, *bomega, *bqlast, *bdqlast, *bqin, *bdq;
    // two-level Jacobi (large matrices): G in HBM, per block pair Gram matrix / rotation, flags and counters
    double *hjG, *hjM, *hjJ, *hj_sigma;
    int *hj_flag, *hj_count, *hj_conv;
    // external point charges (xtb_setExternalCharges): per molecule the range [pc0[im], pc0[im+1]) of the concatenated
    // arrays; pc0 == nullptr: none.  vpc: their potential on the shells, pc_grad: gradient on the charges
    const long long* pc0;
    const double *pc_xyz, *pc_q, *pc_gam;
    double *pc_grad, *vpc;
    // blocked Cholesky of the overlap matrix (kernels_chol.cuh): inverse diagonal blocks, chol_stride doubles per molecule
    // (their number of 64-blocks times 64 x 64 can exceed nao^2 for small members of a batch, so they have their own buffer)
    double* cholD;
    long long chol_stride;
    // implicit solvation, GBSA / ALPB (kernels_solv.cuh; xtb_setSolvent): solv_on == 0: none.  One model per batch.
    int solv_on, solv_kernel /* 1 Still, 2 P16 */, solv_salt, solv_hb;
    double solv_keps, solv_alpbet, solv_gshift, solv_c1, solv_eps, solv_kappa, solv_ionrad;
    const double* sv_par;   // [8] per atom: vdwr, rho, svdw, vdwsa, wrp, gamsasa, hbmag, vdwr^3
    double *sv_brad, *sv_dbr, *sv_sasa, *sv_ion /* [2]: ionscr, discr */, *sv_grddb, *sv_bq, *sv_shift;
    double *sv_bmat;        // Born interaction matrix, nat^2 per molecule (offsets nn0)
    double *sv_dsdrt;       // d sasa_i / d R_j, [i][j][3], 3 nat^2 per molecule
    double *sv_mol;         // [24] per molecule: surface energy, shape descriptor, inertia tensor, energy parts
    // shell-pair gradient partials
    double *gpair;   // [npair][4] : gx,gy,gz,dCN
    const long long* pair0;
    // shell pairs of a molecule in the order the integral kernels work through them: descending cost (components x
    // primitive pairs), so that the warps of a CTA hold pairs of one kind and finish together; pord0[im]: the molecule's
    // slice of pair_order (molecules of one composition share it)
    const int* pair_order;
    const long long* pord0;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// Atom loops of the O(N^2) pair kernels: kAS consecutive lanes share one atom i and stride over its partners j, the
// partial sums are combined by xor shuffles in a fixed order (bit-reproducible).  A single 200-atom molecule then spreads
// over 8x more threads than with one thread per atom; ensembles lose nothing (their atoms already fill the GPU).
// No thread may leave such a kernel before the shuffles: inactive groups run empty loops.
constexpr int kAS = 8;
__device__ __forceinline__ double slice_sum(double v) {
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

// deterministic block reduction (sum); result valid in all threads. red: >= 32 doubles of shared memory
__device__ __forceinline__ double block_sum(double v, double* red) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    double r = 0.0;
    for (int i = 0; i < nw; ++i) r += red[i];
    return r;
}

// Ownership rules of the multi-GPU split: atoms are owned in warp-sized chunks (whole warps of the thread-per-atom
// kernels drop out), shell-pair CTAs are interleaved over the ranks (neighbouring CTAs have similar cost).
__host__ __device__ __forceinline__ int shard_atom_owner(int i, int nranks) { return nranks <= 1 ? 0 : (i >> 5) % nranks; }
__device__ __forceinline__ bool shard_owns_atom(const Dev& d, int i) { return shard_atom_owner(i, d.snranks) == d.srank; }
__device__ __forceinline__ long long shard_pair_cta(const Dev& d) { return (long long)blockIdx.x * d.snranks + d.srank; }

__device__ __forceinline__ int lin0(int a, int b) {  // packed index, 0-based (src/lin_mod.f90:28-36)
    int mx = a > b ? a : b, mn = a > b ? b : a;
    return mn + mx * (mx + 1) / 2;
}

// one decision of molecule im counted (diagnostic, see Dev::cnt)
__device__ __forceinline__ void cnt_add(const Dev& d, int im, int k, unsigned long long v) {
    if (d.cnt && v) atomicAdd(d.cnt + (size_t)kNCnt * im + k, v);
}

}  // namespace xtbb200
