// B200-native batched GFN2-xTB single point + gradient engine (host driver).
//
// One Batch object = a set of molecules resident on one GPU.  All per-geometry state lives in HBM;
// the host only launches kernels and polls a 4-byte "molecules still iterating" counter.
// Reference path replaced: TxTBCalculator%singlepoint -> scf -> scc
// (src/xtb/calculator.f90:195-350, src/scf_module.F90:74-920, src/scc_core.f90:255-640).
#include "engine.h"

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "kernels_geom.cuh"
#include "kernels_grad.cuh"
#include "kernels_int.cuh"
#include "kernels_hjacobi.cuh"
#include "kernels_chol.cuh"
#include "kernels_eigh.cuh"
#include "kernels_linalg.cuh"
#include "kernels_scc.cuh"
#include "kernels_solv.cuh"
#include "kernels_md.cuh"
#include "md_host.h"
#include "solvation.h"

namespace xtbb200 {

__global__ void k_reset_state(Dev d, int keep_wfn) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.nmol) return;
    MolState& s = d.st[i];
    s.iter = 0; s.converged = 0; s.lastdiag = 0; s.done = 0; s.fail = 0; s.kocc = 0; s.ndp = 0; s.nqp = 0;
    s.sweeps = 0; s.sweeps_total = 0;
    if (!keep_wfn) s.zvalid = 0;
    s.eel = 0; s.eold = 0; s.ees = 0; s.eaes = 0; s.epol = 0; s.ed = 0; s.egap = 0; s.ga = 0; s.gb = 0;
    s.efa = 0; s.efb = 0; s.rmsq = 0; s.ep = 0; s.embd = 0; s.eh0 = 0; s.energy = 0; s.gnorm = 0;
}
__global__ void k_count_active(Dev d, int* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int a = (i < d.nmol && !d.st[i].done) ? 1 : 0;
    unsigned b = __ballot_sync(0xffffffffu, a);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(out, __popc(b));
}

// after the Householder / divide-and-conquer eigensolver: bookkeeping of the molecules that were diagonalised
__global__ void k_eigh_finish(Dev d, const int* fail, int fail_stride) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.nmol) return;
    MolState& s = d.st[i];
    if (s.done) return;
    s.sweeps = 0;
    s.zvalid = 1;
    if (fail[(long long)i * fail_stride]) s.fail = 12;
}

// kernel classes for the per-kernel device timers (bench.py roofline, profiles/)
enum KernelId { K_SETUP = 0, K_EEQ, K_SDQH0, K_DQMASK, K_CHOL, K_POT, K_D4SCC, K_FOCK, K_GEMM_XHX, K_JACOBI, K_GEMM_C,
                K_FERMI, K_DENSITY, K_POP, K_ENERGY, K_SCCFIN, K_GRAD_H0, K_GRAD_AES, K_GRAD_D4, K_GRAD_ATM, K_GRAD_MISC,
                K_WBO, K_FINISH, K_COMM, K_EIG_PANEL, K_EIG_TRAIL, K_EIG_DC, K_EIG_BT, K_SOLV_SETUP, K_SOLV_SCC, K_SOLV_GRAD, K_MD, K_COUNT };
static const char* const kKernelNames[K_COUNT] = {
    "setup(cn,gab,jmat,dispmat,rep)", "eeq_guess", "k_sdqh0", "k_dqmask", "k_chol_inv", "potentials(shell,aes)",
    "d4_scc(zeta,matvec,shift)", "k_fock", "k_gemm(X^T H X)", "k_jacobi",   /* = the eigensolver class, whichever kernel family runs */ "k_gemm(C=XZ)", "k_fermi",
    "density(scale+k_gemm)", "populations(k_pop_rows,k_pop_atoms)", "energies(es,aes)", "k_scc_finish(broyden)",
    "k_grad_h0(+W build,collect)", "k_grad_aes", "grad_d4(weights,c6,pairs)", "k_grad_atm", "k_grad_cn", "wbo(PS gemm,k_wbo)",
    "k_finish", "nccl_allreduce(S,D,Q,H0 | dE/dCN | gradient)", "eig:k_sytrd_panel", "eig:trailing update(k_eg_gemm)",
    "eig:divide and conquer", "eig:back-transformation", "solv:setup(born,sasa,shape,bornmat)", "solv:scc(k_solv_matvec)",
    "solv:gradient(pairs,energy,chain)", "k_md_update"};

// NCCL is needed only when one large system is split over GPUs; it is resolved at run time so that the
// single-GPU library has no link dependency on it.  Inside a PyTorch process the already loaded torch-bundled
// libnccl.so.2 is picked up (same SONAME); XTBB200_NCCL_LIB overrides the name.
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    static NcclApi& get() {
        static NcclApi api;
        if (api.handle) return api;
        const char* name = std::getenv("XTBB200_NCCL_LIB");
        void* h = dlopen(name ? name : "libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) throw std::runtime_error(std::string("xtb_b200: cannot load NCCL for the multi-GPU split: ") + dlerror());
        auto sym = [&](const char* n) {
            void* p = dlsym(h, n);
            if (!p) throw std::runtime_error(std::string("xtb_b200: NCCL symbol missing: ") + n);
            return p;
        };
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
        api.Broadcast = (decltype(api.Broadcast))sym("ncclBroadcast");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
        api.handle = h;
        return api;
    }
    void check(ncclResult_t r, const char* what) const {
        if (r != ncclSuccess) throw std::runtime_error(std::string("NCCL error in ") + what + ": " + GetErrorString(r));
    }
};

void shard_unique_id(char* out128) {
    NcclApi& n = NcclApi::get();
    ncclUniqueId id;
    n.check(n.GetUniqueId(&id), "ncclGetUniqueId");
    static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
    std::memcpy(out128, &id, sizeof(id));
}
int shard_atom_owner_host(int iatom, int nranks) { return shard_atom_owner(iatom, nranks); }
int shard_pair_cta_owner_host(long long cta, int nranks) { return nranks <= 1 ? 0 : (int)(cta % nranks); }

struct Batch::Impl {
    ncclComm_t comm = nullptr;
    double* xbuf = nullptr;   // exchange buffer of the gradient phase, 4 * natoms doubles
    float t_comm = 0;
    void allreduce(double* p, size_t count, cudaStream_t st) {
        NcclApi& n = NcclApi::get();
        n.check(n.AllReduce(p, p, count, ncclDouble, ncclSum, comm, st), "ncclAllReduce");
    }
    // column blocks of an n x n matrix, block r (tile columns split[r] .. split[r + 1], gt columns each) valid on rank r ->
    // every rank holds the whole matrix: one in-place broadcast per owner, grouped
    void gather_cols(double* p, int n, const std::vector<int>& split, int gt, cudaStream_t st) {
        NcclApi& a = NcclApi::get();
        a.check(a.GroupStart(), "ncclGroupStart");
        for (int r = 0; r + 1 < (int)split.size(); ++r) {
            const long long c0 = std::min<long long>(n, (long long)split[r] * gt), c1 = std::min<long long>(n, (long long)split[r + 1] * gt);
            if (c1 > c0) a.check(a.Broadcast(p + c0 * n, p + c0 * n, (size_t)((c1 - c0) * n), ncclDouble, r, comm, st), "ncclBroadcast");
        }
        a.check(a.GroupEnd(), "ncclGroupEnd");
    }
    // per-kernel-class timers: events bracket every launch whose class bit is set in prof_mask
    unsigned long long prof_mask = 0;
    std::vector<cudaEvent_t> ev_pool;
    std::vector<int> ev_ids;       // class of each recorded (begin,end) pair
    size_t ev_used = 0;
    float k_ms[K_COUNT] = {0};
    int k_calls[K_COUNT] = {0};
    int k_launches[K_COUNT] = {0};
    void prof_begin(int id, cudaStream_t st) {
        if (!((prof_mask >> id) & 1ull)) return;
        if (ev_used + 2 > ev_pool.size()) {
            size_t old = ev_pool.size();
            ev_pool.resize(old + 256);
            for (size_t i = old; i < ev_pool.size(); ++i) cudaEventCreate(&ev_pool[i]);
        }
        cudaEventRecord(ev_pool[ev_used], st);
    }
    void prof_end(int id, cudaStream_t st) {
        if (!((prof_mask >> id) & 1ull)) return;
        cudaEventRecord(ev_pool[ev_used + 1], st);
        ev_ids.push_back(id);
        ev_used += 2;
    }
    void prof_collect() {   // call after the stream is synchronised
        for (size_t i = 0; i < ev_ids.size(); ++i) {
            float ms = 0;
            cudaEventElapsedTime(&ms, ev_pool[2 * i], ev_pool[2 * i + 1]);
            k_ms[ev_ids[i]] += ms;
            k_calls[ev_ids[i]]++;
        }
        ev_ids.clear();
        ev_used = 0;
    }
    BatchSystem sys;
    Dev d;
    EighWork eig;             // Householder + divide-and-conquer eigensolver for matrices beyond shared memory
    bool have_eig = false;
    bool blocked_chol = false;     // S = L L^T, X = L^-T by the blocked DMMA kernels instead of one CTA per molecule
    bool cta_eig = false;          // batched Householder + divide and conquer for ensembles of one matrix order (solver mode 4)
    bool has_followers = false;    // some molecules share their composition with an earlier one (conformer ensemble)
    std::vector<long long> h_pc0;  // external point charges: per-molecule ranges (empty: none)
    int max_npc = 0;
    long long pc_cap = 0;          // capacity (charges) of the device buffers below
    long long* pc0_dev = nullptr;
    double* pc_buf = nullptr;      // [gam | q | xyz(3) | grad(3)] x pc_cap
    MdDev md;                 // molecular dynamics state (md_ready == false: md_init not called)
    bool md_ready = false, md_first = true;
    long long md_ncons = 0;
    std::vector<double> md_out;
    std::vector<void*> md_allocs;   // device buffers of the current constraint tables (subset of allocs)
    SolvModel solv;           // implicit solvation model (solv.on == false: none)
    double* sv_par_dev = nullptr;
    bool leb_uploaded = false;
    std::vector<int> h_done;  // per-molecule "left the SCC loop" flags, mirrored for the per-molecule solver loop
    std::vector<void*> allocs;
    std::vector<long long> eeq0, pair0;
    int* d_active = nullptr;
    cudaStream_t stream = nullptr;
    bool have_wfn = false;
    bool positions_set = false;
    std::vector<MolState> h_state;
    int device = 0;
    // options
    double acc = 1.0, etemp = 300.0;
    int maxiter = 250;
    int guess = 0;
    int solver = 0;     // 0: automatic, 1: scalar one-sided Jacobi (k_jacobi) even where the block kernel fits
    bool fits_blk = false, fits_scalar = false;
    // diagnostic knobs, read once at construction and only under XTBB200_TUNING=1 (common.cuh: tuning_env)
    double knob_jb_tol = 0.0, knob_jb_qtol = kJbQuadraticStop;
    bool iterlog_on = false;
    double* iterlog_buf = nullptr;
    int iterlog_rows = 0;
    unsigned long long* cnt_buf = nullptr;   // decision counters (diagnostic), allocated by set_decision_counting
    int knob_jb_inner = kJbInnerSweeps;
    int knob_jb_variant = 1;    // 1: jb_sweeps2 (cross-block Gram, unnormalised rotations), 0: jb_sweeps
    bool knob_no_cluster_eig = false, knob_cta_chol = false, knob_cluster_single = false, knob_no_gemm_split = false;
    size_t bytes = 0;
    // timing (ms) of the last single point
    float t_setup = 0, t_scc = 0, t_grad = 0, t_fact = 0;   // t_fact: overlap factorisation, part of t_setup
    int launches = 0;
    int scc_passes = 0;
    cudaEvent_t ev_phase[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // [4]: begin of the overlap factorisation
    cudaEvent_t ev_call[2] = {nullptr, nullptr};
    cudaEvent_t ev_comm[4] = {nullptr, nullptr, nullptr, nullptr};   // integral exchange begin/end, gradient exchange begin/end
    cudaEvent_t ev_gather[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // the three gathers of an SCC iteration
    float t_gather = 0;                                              // summed over the iterations of the last call
    long long pair_total = 0;
    float t_call = 0;

    template <class T>
    T* dalloc(size_t n) {
        void* p = nullptr;
        size_t b = std::max<size_t>(n, 1) * sizeof(T);
        XTB_CUDA_CHECK(cudaMalloc(&p, b));
        XTB_CUDA_CHECK(cudaMemsetAsync(p, 0, b, stream));
        allocs.push_back(p);
        bytes += b;
        return (T*)p;
    }
    template <class T>
    T* upload(const std::vector<T>& v) {
        T* p = dalloc<T>(v.size());
        if (!v.empty()) XTB_CUDA_CHECK(cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, stream));
        return p;
    }
    ~Impl() {
        for (void* p : allocs) cudaFree(p);
        if (pc_buf) cudaFree(pc_buf);
        if (have_eig) eig.release();
        for (auto& e : ev_pool) cudaEventDestroy(e);
        for (auto& e : ev_phase) if (e) cudaEventDestroy(e);
        for (auto& e : ev_call) if (e) cudaEventDestroy(e);
        for (auto& e : ev_comm) if (e) cudaEventDestroy(e);
        for (auto& e : ev_gather) if (e) cudaEventDestroy(e);
        if (comm) NcclApi::get().CommDestroy(comm);
        if (stream) cudaStreamDestroy(stream);
    }
};

int device_count() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

static dim3 grid2(long long items, int block, int nmol) { return dim3((unsigned)std::max<long long>(1, (items + block - 1) / block), nmol); }

Batch::Batch(int nmol, const int* natoms, const int* numbers, const double* charges, const int* uhf, int device) {
    p_ = new Impl();
    Impl& I = *p_;
    try {
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0)
            throw std::runtime_error("xtb_b200: no CUDA device available (this engine has no CPU fallback)");
        I.device = device;
        XTB_CUDA_CHECK(cudaSetDevice(device));
        XTB_CUDA_CHECK(cudaStreamCreate(&I.stream));
        if (const char* e = tuning_env("XTBB200_JB_TOL")) I.knob_jb_tol = std::atof(e);
        if (const char* e = tuning_env("XTBB200_JB_QTOL")) I.knob_jb_qtol = std::atof(e);
        if (const char* e = tuning_env("XTBB200_JB_INNER")) I.knob_jb_inner = std::max(1, std::atoi(e));
        if (const char* e = tuning_env("XTBB200_JB_VARIANT")) I.knob_jb_variant = std::atoi(e) ? 1 : 0;
        if (const char* e = tuning_env("XTBB200_SOLVER")) I.solver = std::atoi(e);   // initial value only: set_solver() wins
        I.knob_no_cluster_eig = tuning_env("XTBB200_NO_CLUSTER_EIG") != nullptr;
        I.knob_cluster_single = tuning_env("XTBB200_CLUSTER_SINGLE") != nullptr;
        I.knob_no_gemm_split = tuning_env("XTBB200_NO_GEMM_SPLIT") != nullptr;
        I.knob_cta_chol = tuning_env("XTBB200_CTA_CHOL") != nullptr;
        const int maxhist = 120;
        int off = 0;
        for (int i = 0; i < nmol; ++i) {
            I.sys.add_molecule(natoms[i], numbers + off, charges ? charges[i] : 0.0, uhf ? uhf[i] : 0, maxhist);
            off += natoms[i];
        }
        I.sys.finalize();
        BatchSystem& s = I.sys;
        Dev& d = I.d;
        std::memset(&d, 0, sizeof(d));
        d.nmol = nmol; d.ntype = (int)s.etypes.size(); d.maxhist = maxhist;
        d.srank = 0; d.snranks = 1;
        d.mol = I.upload(s.mol);
        d.atoms = I.upload(s.atoms);
        d.shells = I.upload(s.shells);
        d.refs = I.upload(s.refs);
        d.ao2sh = I.upload(s.ao2sh);
        d.ao2at = I.upload(s.ao2at);
        d.c6ref = I.upload(s.c6ref);
        d.st = I.dalloc<MolState>(nmol);
        const size_t na = s.atoms.size(), ns = s.shells.size(), nao = s.ao2sh.size(), nr = s.refs.size();
        d.xyz = I.dalloc<double>(3 * na); d.cn = I.dalloc<double>(na); d.cn_cov = I.dalloc<double>(na);
        d.cn_d4 = I.dalloc<double>(na); d.cn_erf = I.dalloc<double>(na); d.radcn = I.dalloc<double>(na);
        d.q = I.dalloc<double>(na); d.q_out = I.dalloc<double>(na);
        d.dipm = I.dalloc<double>(3 * na); d.dipm_out = I.dalloc<double>(3 * na);
        d.qp = I.dalloc<double>(6 * na); d.qp_out = I.dalloc<double>(6 * na);
        d.vs = I.dalloc<double>(na); d.vd = I.dalloc<double>(3 * na); d.vq = I.dalloc<double>(6 * na);
        d.ashift = I.dalloc<double>(na);
        d.grad = I.dalloc<double>(3 * na); d.dhdcn = I.dalloc<double>(na); d.w_aes = I.dalloc<double>(na);
        d.dEdcn = I.dalloc<double>(na); d.e_atom = I.dalloc<double>(4 * na); d.zw = I.dalloc<double>(35 * na);
        d.selfE = I.dalloc<double>(ns); d.qsh = I.dalloc<double>(ns); d.qsh_out = I.dalloc<double>(ns);
        d.sshift = I.dalloc<double>(ns); d.jq = I.dalloc<double>(ns);
        d.emo = I.dalloc<double>(nao); d.focc = I.dalloc<double>(nao); d.focca = I.dalloc<double>(nao);
        d.foccb = I.dalloc<double>(nao); d.aop = I.dalloc<double>(12 * nao); d.nrm = I.dalloc<double>(nao);
        d.rank = I.dalloc<int>(nao);
        d.gw = I.dalloc<double>(nr); d.zeta = I.dalloc<double>(nr); d.dzeta = I.dalloc<double>(nr);
        d.dvec = I.dalloc<double>(nr); d.edref = nullptr;
        d.tot_mat = s.tot_mat;
        const size_t tm = (size_t)s.tot_mat;
        d.S = I.dalloc<double>(tm); d.D = I.dalloc<double>(3 * tm); d.Q = I.dalloc<double>(6 * tm);
        d.H0 = I.dalloc<double>(tm); d.H0n = I.dalloc<double>(tm); d.H = I.dalloc<double>(tm);
        d.X = I.dalloc<double>(tm); d.T = I.dalloc<double>(tm); d.Zm = I.dalloc<double>(tm);
        d.C = I.dalloc<double>(tm); d.P = I.dalloc<double>(tm); d.W = I.dalloc<double>(tm);
        d.mask = I.dalloc<uint8_t>(tm);
        const size_t tn = (size_t)s.tot_nn;
        d.gab3 = I.dalloc<double>(tn); d.gab5 = I.dalloc<double>(tn); d.c6 = I.dalloc<double>(tn);
        d.dc6 = I.dalloc<double>(tn); d.c60 = I.dalloc<double>(tn); d.dc60 = I.dalloc<double>(tn);
        d.wbo = I.dalloc<double>(tn);
        d.jmat = I.dalloc<double>((size_t)s.tot_ss);
        d.dispmat = I.dalloc<double>((size_t)s.tot_dd);
        long long te = 0, tp = 0;
        for (int i = 0; i < nmol; ++i) {
            I.eeq0.push_back(te); te += (long long)(s.mol[i].nat + 1) * (s.mol[i].nat + 2);
            I.pair0.push_back(tp); tp += (long long)s.mol[i].nsh * (s.mol[i].nsh + 1) / 2;
        }
        d.eeqA = I.dalloc<double>((size_t)te);
        d.eeq0 = I.upload(I.eeq0);
        d.gpair = I.dalloc<double>(4 * (size_t)tp);
        I.pair_total = tp;
        d.pair0 = I.upload(I.pair0);
        d.bdf = I.dalloc<double>((size_t)s.tot_br); d.bu = I.dalloc<double>((size_t)s.tot_br);
        d.ba = I.dalloc<double>((size_t)nmol * maxhist * maxhist); d.bbeta = I.dalloc<double>((size_t)nmol * maxhist * maxhist);
        d.bomega = I.dalloc<double>((size_t)nmol * maxhist);
        d.bqlast = I.dalloc<double>((size_t)s.tot_q); d.bdqlast = I.dalloc<double>((size_t)s.tot_q);
        d.bqin = I.dalloc<double>((size_t)s.tot_q); d.bdq = I.dalloc<double>((size_t)s.tot_q);
        // solver selection by the largest matrix of the batch: shared-memory kernels when it fits, two-level Jacobi else
        {
            const int jbw = std::max(1, std::min(20, jb_dims(s.max_nao).nb / 2));
            I.fits_blk = jb_smem_bytes(s.max_nao, jbw) <= (size_t)(227 * 1024 - 1024);
            I.fits_scalar = (size_t)s.max_nao * (s.max_nao | 1) * sizeof(double) <= (size_t)(227 * 1024 - 4096);
        }
        if (!I.fits_blk || !I.fits_scalar) {
            d.hjG = I.dalloc<double>(tm);
            d.hjM = I.dalloc<double>((size_t)s.tot_hj); d.hjJ = I.dalloc<double>((size_t)s.tot_hj);
            d.hj_flag = I.dalloc<int>((size_t)(s.tot_hj / (128 * 128)));
            d.hj_count = I.dalloc<int>(nmol); d.hj_conv = I.dalloc<int>(nmol); d.hj_sigma = I.dalloc<double>(nmol);
            I.eig.init(s.max_nao, device);
            I.have_eig = true;
            I.bytes += I.eig.bytes;
        }
        // blocked Cholesky (kernels_chol.cuh) from kBlockedCholMin AOs on, also when the eigensolver runs in shared memory
        I.blocked_chol = (!I.fits_blk || !I.fits_scalar) || s.max_nao >= kBlockedCholMin;
        if (I.blocked_chol) {
            d.chol_stride = (long long)((s.max_nao + CB - 1) / CB) * CB * CB;
            d.cholD = I.dalloc<double>((size_t)nmol * d.chol_stride);
        }
        I.h_done.assign(nmol, 0);
        {   // leaders of the groups of identical compositions (warm start of the first diagonalisation, k_lead_copy)
            std::vector<int> lead(nmol);
            std::map<std::vector<int>, int> first;
            int o2 = 0;
            for (int i = 0; i < nmol; ++i) {
                std::vector<int> key(numbers + o2, numbers + o2 + natoms[i]);
                key.push_back((int)std::lround(100 * (charges ? charges[i] : 0.0)));
                key.push_back(uhf ? uhf[i] : 0);
                auto it = first.find(key);
                if (it == first.end()) { first[key] = i; lead[i] = i; }
                else { lead[i] = it->second; I.has_followers = true; }
                o2 += natoms[i];
            }
            d.leader = I.upload(lead);
            // cost-sorted shell-pair order, one list per composition (k_sdqh0 / k_grad_h0)
            std::vector<long long> pord0(nmol);
            std::vector<int> order;
            static const int ncart[3] = {1, 3, 6};
            for (int i = 0; i < nmol; ++i) {
                if (lead[i] != i) { pord0[i] = pord0[lead[i]]; continue; }
                const MolDesc& mm = s.mol[i];
                const long long np = (long long)mm.nsh * (mm.nsh + 1) / 2;
                pord0[i] = (long long)order.size();
                std::vector<std::pair<int, int>> cost((size_t)np);
                int pidx = 0;
                for (int is = 0; is < mm.nsh; ++is)
                    for (int js = 0; js <= is; ++js, ++pidx) {
                        const ShellPar& a = s.shells[mm.sh0 + is];
                        const ShellPar& b = s.shells[mm.sh0 + js];
                        const int ncomp = ncart[a.l] * ncart[b.l];
                        // passes over the component groups of a warp x primitive pairs per lane, as the kernels loop
                        const int lpc = ncomp >= kIntLG ? 1 : kIntLG / ncomp, ngrp = kIntLG / lpc;
                        const int c = ((ncomp + ngrp - 1) / ngrp) * ((a.nprim * b.nprim + lpc - 1) / lpc) * 64 + (a.atom == b.atom ? 0 : 1);
                        cost[pidx] = std::make_pair(-c, pidx);
                    }
                std::sort(cost.begin(), cost.end());
                for (auto& c : cost) order.push_back(c.second);
            }
            d.pair_order = I.upload(order);
            d.pord0 = I.upload(pord0);
            XTB_CUDA_CHECK(cudaStreamSynchronize(I.stream));
        }
        I.d_active = I.dalloc<int>(1);
        I.h_state.resize(nmol);
        XTB_CUDA_CHECK(cudaStreamSynchronize(I.stream));
        // opt in to large dynamic shared memory for the in-SM solvers
        const int smax = 227 * 1024 - 2048;
        cudaFuncSetAttribute(k_chol_inv<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smax);
        cudaFuncSetAttribute(k_jacobi_blk<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smax);
        cudaFuncSetAttribute(k_jacobi_blk<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smax);
        cudaFuncSetAttribute(k_hj_inner, cudaFuncAttributeMaxDynamicSharedMemorySize, smax);
        cudaFuncSetAttribute(k_grad_atm, cudaFuncAttributeMaxDynamicSharedMemorySize, smax);
        XTB_CUDA_CHECK(cudaFuncSetAttribute(k_sdqh0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSdqSmem));
        XTB_CUDA_CHECK(cudaFuncSetAttribute(k_grad_h0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGradSmem));
        cudaFuncSetAttribute(k_bchol_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBcholDiagSmem);
        cudaFuncSetAttribute(k_hj_apply, cudaFuncAttributeMaxDynamicSharedMemorySize, HP * GLD * (int)sizeof(double));
        cudaFuncSetAttribute(k_jacobi<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smax);
        cudaFuncSetAttribute(k_jacobi<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smax);
        cudaFuncSetAttribute(k_jacobi<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smax);
        cudaFuncSetAttribute(k_jacobi<32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smax);
    } catch (...) {
        delete p_;
        p_ = nullptr;
        throw;
    }
}

Batch::~Batch() { delete p_; }

int Batch::nmol() const { return p_->d.nmol; }
int Batch::natoms(int i) const { return p_->sys.mol[i].nat; }
int Batch::nao(int i) const { return p_->sys.mol[i].nao; }
int Batch::nshell(int i) const { return p_->sys.mol[i].nsh; }
int Batch::nel(int i) const { return p_->sys.mol[i].nel; }
int Batch::nopen(int i) const { return p_->sys.mol[i].nopen; }
int Batch::atom_offset(int i) const { return p_->sys.mol[i].at0; }
int Batch::shell_offset(int i) const { return p_->sys.mol[i].sh0; }
int Batch::ao_offset(int i) const { return p_->sys.mol[i].ao0; }
long long Batch::mat_offset(int i) const { return p_->sys.mol[i].mat0; }
long long Batch::nn_offset(int i) const { return p_->sys.mol[i].nn0; }
int Batch::total_atoms() const { return (int)p_->sys.atoms.size(); }
int Batch::total_shells() const { return (int)p_->sys.shells.size(); }
int Batch::total_ao() const { return (int)p_->sys.ao2sh.size(); }
size_t Batch::device_bytes() const { return p_->bytes; }
void* Batch::stream() const { return (void*)p_->stream; }

void Batch::set_options(double accuracy, double etemp, int maxiter) {
    p_->acc = accuracy; p_->etemp = etemp; p_->maxiter = maxiter;
}
void Batch::set_guess(int mode) { p_->guess = mode; }
void Batch::set_solver(int mode) { p_->solver = mode; }
void Batch::set_jacobi_stop(double qtol) { p_->knob_jb_qtol = qtol > 0.0 ? qtol : kJbQuadraticStop; }

void Batch::set_shard(int rank, int nranks, const char* id128) {
    Impl& I = *p_;
    if (nranks < 1 || rank < 0 || rank >= nranks) throw std::runtime_error("set_shard: need 0 <= rank < nranks");
    XTB_CUDA_CHECK(cudaSetDevice(I.device));
    NcclApi* n = nranks > 1 || I.comm ? &NcclApi::get() : nullptr;
    if (I.comm) { n->CommDestroy(I.comm); I.comm = nullptr; }
    I.d.srank = 0; I.d.snranks = 1;
    if (nranks == 1) return;
    if (!id128) throw std::runtime_error("set_shard: the NCCL unique id of rank 0 is required");
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof(id));
    n->check(n->CommInitRank(&I.comm, nranks, id, rank), "ncclCommInitRank");
    if (!I.xbuf) I.xbuf = I.dalloc<double>(4 * I.sys.atoms.size());
    I.d.srank = rank; I.d.snranks = nranks;
    I.have_wfn = false;
    // first collective of a communicator sets up its channels (hundreds of ms): pay that here, not in the first call
    XTB_CUDA_CHECK(cudaMemsetAsync(I.xbuf, 0, 4 * I.sys.atoms.size() * sizeof(double), I.stream));
    I.allreduce(I.xbuf, 4 * I.sys.atoms.size(), I.stream);
    XTB_CUDA_CHECK(cudaStreamSynchronize(I.stream));
}
int Batch::shard_rank() const { return p_->d.srank; }
int Batch::shard_nranks() const { return p_->d.snranks; }
float Batch::comm_ms() const { return p_->t_comm; }

void Batch::set_positions(const double* xyz_host) {
    Impl& I = *p_;
    XTB_CUDA_CHECK(cudaSetDevice(I.device));
    XTB_CUDA_CHECK(cudaMemcpyAsync(I.d.xyz, xyz_host, 3 * I.sys.atoms.size() * sizeof(double), cudaMemcpyHostToDevice, I.stream));
    I.positions_set = true;
}
void Batch::set_positions_device(const double* xyz_dev) {
    Impl& I = *p_;
    XTB_CUDA_CHECK(cudaSetDevice(I.device));
    XTB_CUDA_CHECK(cudaMemcpyAsync(I.d.xyz, xyz_dev, 3 * I.sys.atoms.size() * sizeof(double), cudaMemcpyDeviceToDevice, I.stream));
    I.positions_set = true;
}
void Batch::invalidate_wavefunction() { p_->have_wfn = false; }
bool Batch::has_wavefunction() const { return p_->have_wfn; }

void Batch::set_wavefunction(const double* qsh, const double* dipm, const double* qp) {
    Impl& I = *p_;
    XTB_CUDA_CHECK(cudaSetDevice(I.device));
    XTB_CUDA_CHECK(cudaMemcpyAsync(I.d.qsh, qsh, I.sys.shells.size() * sizeof(double), cudaMemcpyHostToDevice, I.stream));
    XTB_CUDA_CHECK(cudaMemcpyAsync(I.d.dipm, dipm, 3 * I.sys.atoms.size() * sizeof(double), cudaMemcpyHostToDevice, I.stream));
    XTB_CUDA_CHECK(cudaMemcpyAsync(I.d.qp, qp, 6 * I.sys.atoms.size() * sizeof(double), cudaMemcpyHostToDevice, I.stream));
    I.have_wfn = true;
}

void Batch::set_external_charges(const int* npc, const int* numbers, const double* charges, const double* xyz) {
    Impl& I = *p_;
    XTB_CUDA_CHECK(cudaSetDevice(I.device));
    XTB_CUDA_CHECK(cudaStreamSynchronize(I.stream));
    const int nmol = I.d.nmol;
    long long tot = 0;
    if (npc) for (int i = 0; i < nmol; ++i) { if (npc[i] < 0) throw std::runtime_error("negative number of point charges"); tot += npc[i]; }
    I.d.pc0 = nullptr; I.d.pc_xyz = I.d.pc_q = I.d.pc_gam = nullptr; I.d.pc_grad = nullptr;
    I.h_pc0.clear(); I.max_npc = 0;
    if (tot == 0) return;
    if (!numbers || !charges || !xyz) throw std::runtime_error("external charges: numbers, charges and positions are required");
    I.h_pc0.assign(nmol + 1, 0);
    for (int i = 0; i < nmol; ++i) { I.h_pc0[i + 1] = I.h_pc0[i] + npc[i]; I.max_npc = std::max(I.max_npc, npc[i]); }
    std::vector<double> gam(tot);
    for (long long k = 0; k < tot; ++k) {
        if (numbers[k] < 1 || numbers[k] > 86) throw std::runtime_error("external charges: element number outside 1..86");
        gam[k] = gfn2_chemical_hardness[numbers[k] - 1];   // src/api/calculator.f90:533
    }
    if (!I.pc0_dev) I.pc0_dev = I.dalloc<long long>(nmol + 1);
    if (!I.d.vpc) I.d.vpc = I.dalloc<double>(I.sys.shells.size());
    if (tot > I.pc_cap) {   // grow-only buffer, so that per-step updates (QM/MM) do not allocate
        if (I.pc_buf) cudaFree(I.pc_buf);
        I.pc_buf = nullptr;
        XTB_CUDA_CHECK(cudaMalloc((void**)&I.pc_buf, (size_t)tot * 8 * sizeof(double)));
        I.pc_cap = tot;
    }
    double* gam_d = I.pc_buf;
    double* q_d = I.pc_buf + I.pc_cap;
    double* xyz_d = I.pc_buf + 2 * I.pc_cap;
    double* grd_d = I.pc_buf + 5 * I.pc_cap;
    XTB_CUDA_CHECK(cudaMemcpyAsync(I.pc0_dev, I.h_pc0.data(), (nmol + 1) * sizeof(long long), cudaMemcpyHostToDevice, I.stream));
    XTB_CUDA_CHECK(cudaMemcpyAsync(gam_d, gam.data(), tot * sizeof(double), cudaMemcpyHostToDevice, I.stream));
    XTB_CUDA_CHECK(cudaMemcpyAsync(q_d, charges, tot * sizeof(double), cudaMemcpyHostToDevice, I.stream));
    XTB_CUDA_CHECK(cudaMemcpyAsync(xyz_d, xyz, 3 * tot * sizeof(double), cudaMemcpyHostToDevice, I.stream));
    XTB_CUDA_CHECK(cudaMemsetAsync(grd_d, 0, 3 * tot * sizeof(double), I.stream));
    XTB_CUDA_CHECK(cudaStreamSynchronize(I.stream));
    I.d.pc0 = I.pc0_dev; I.d.pc_gam = gam_d; I.d.pc_q = q_d; I.d.pc_xyz = xyz_d; I.d.pc_grad = grd_d;
}
void Batch::set_solvent(const char* solvent, bool alpb, int kernel, int state, double temperature, double ion_strength,
                        double ion_rad) {
    Impl& I = *p_;
    Dev& d = I.d;
    if (kernel != 1 && kernel != 2) throw std::runtime_error("solvation: kernel must be 1 (Still) or 2 (P16)");
    SolvModel sm = make_solv_model(solvent, alpb, kernel, state, temperature, ion_strength, ion_rad);   // throws: unknown solvent
    XTB_CUDA_CHECK(cudaSetDevice(I.device));
    XTB_CUDA_CHECK(cudaStreamSynchronize(I.stream));
    const size_t na = I.sys.atoms.size(), tn = (size_t)I.sys.tot_nn;
    std::vector<double> par(8 * na);
    bool hb = false;
    for (size_t i = 0; i < na; ++i) { solv_atom_par(sm, I.sys.atoms[i].Z, &par[8 * i]); hb = hb || par[8 * i + 6] < 0.0; }
    if (!I.sv_par_dev) {
        I.sv_par_dev = I.dalloc<double>(8 * na);
        d.sv_brad = I.dalloc<double>(na); d.sv_dbr = I.dalloc<double>(na); d.sv_sasa = I.dalloc<double>(na);
        d.sv_ion = I.dalloc<double>(2 * na); d.sv_grddb = I.dalloc<double>(na); d.sv_bq = I.dalloc<double>(na);
        d.sv_shift = I.dalloc<double>(na);
        d.sv_bmat = I.dalloc<double>(tn); d.sv_dsdrt = I.dalloc<double>(3 * tn);
        d.sv_mol = I.dalloc<double>((size_t)SVM_STRIDE * d.nmol);
    }
    if (!I.leb_uploaded) {
        XTB_CUDA_CHECK(cudaMemcpyToSymbolAsync(c_leb230, solv_leb230, sizeof(double) * kSolvNang * 4, 0, cudaMemcpyHostToDevice, I.stream));
        I.leb_uploaded = true;
    }
    XTB_CUDA_CHECK(cudaMemcpyAsync(I.sv_par_dev, par.data(), par.size() * sizeof(double), cudaMemcpyHostToDevice, I.stream));
    XTB_CUDA_CHECK(cudaStreamSynchronize(I.stream));
    d.sv_par = I.sv_par_dev;
    d.solv_on = 1; d.solv_kernel = kernel; d.solv_salt = sm.salt ? 1 : 0; d.solv_hb = hb ? 1 : 0;
    d.solv_keps = sm.keps; d.solv_alpbet = sm.alpbet; d.solv_gshift = sm.gshift; d.solv_c1 = sm.c1; d.solv_eps = sm.eps;
    d.solv_kappa = sm.kappa; d.solv_ionrad = sm.ion_rad;
    I.solv = sm;
}
void Batch::release_solvent() { p_->d.solv_on = 0; p_->solv.on = false; }
bool Batch::has_solvent() const { return p_->d.solv_on != 0; }
void Batch::release_external_charges() { set_external_charges(nullptr, nullptr, nullptr, nullptr); }
long long Batch::external_charge_count() const { return p_->h_pc0.empty() ? 0 : p_->h_pc0.back(); }

// ---- molecular dynamics (kernels_md.cuh) ----
void Batch::md_init(const double* xyz_host, double tstep_fs, double temperature, bool thermostat, int shake_mode, int hmass,
                    const double* velocities, unsigned long long seed, bool restart) {
    Impl& I = *p_;
    Dev& d = I.d;
    if (!(tstep_fs > 0.0) || !(temperature > 0.0)) throw std::runtime_error("md: time step and temperature must be positive");
    if (shake_mode < 0 || shake_mode > 2) throw std::runtime_error("md: shake mode must be 0, 1 or 2");
    const double fstoau = kFsToAU;
    const int nmol = d.nmol;
    const size_t na = I.sys.atoms.size();
    set_positions(xyz_host);
    singlepoint(restart);   // the single point of the main program: wavefunction and bond orders (src/prog/main.F90:1147)
    std::vector<double> wbo((size_t)I.sys.tot_nn), mass(na), velo(3 * na), nfree(nmol);
    fetch("wbo", wbo.data());
    // masses, dynamic.f90:237-252
    for (size_t a = 0; a < na; ++a) mass[a] = md_mass(I.sys.atoms[a].Z, hmass);
    // SHAKE constraints, init_shake (src/shake_module.f90:50-148)
    std::vector<long long> cons0(nmol + 1, 0), adj0(na + 1, 0);
    std::vector<int> cons, adj;
    std::vector<double> dist2;
    for (int im = 0; im < nmol; ++im) {
        const MolDesc& m = I.sys.mol[im];
        const double* x = xyz_host + 3 * (size_t)m.at0;
        const double* w = wbo.data() + m.nn0;
        std::vector<int> zs(m.nat);
        for (int i = 0; i < m.nat; ++i) zs[i] = I.sys.atoms[m.at0 + i].Z;
        const std::vector<std::pair<int, int>> list = shake_constraint_list(m.nat, zs.data(), x, w, shake_mode);
        auto r2 = [&](int i, int j) { double s = 0; for (int k = 0; k < 3; ++k) { const double t = x[3 * i + k] - x[3 * j + k]; s += t * t; } return s; };
        std::vector<std::vector<int>> per_atom(m.nat);
        for (size_t c = 0; c < list.size(); ++c) {
            cons.push_back(list[c].first); cons.push_back(list[c].second);
            dist2.push_back(r2(list[c].first, list[c].second));
            per_atom[list[c].first].push_back(2 * (int)c);
            per_atom[list[c].second].push_back(2 * (int)c + 1);
        }
        cons0[im + 1] = cons0[im] + (long long)list.size();
        for (int i = 0; i < m.nat; ++i) {
            adj.insert(adj.end(), per_atom[i].begin(), per_atom[i].end());
            adj0[m.at0 + i + 1] = (long long)adj.size();
        }
        nfree[im] = 3.0 * m.nat - (shake_mode ? (double)list.size() : 0.0);   // dynamic.f90:277-280
    }
    I.md_ncons = cons0[nmol];
    // velocities, mdinitu (dynamic.f90:850-890): every component +-sqrt(2 e / m), hydrogens twice that; e = Ekin / (3 n)
    if (velocities) velo.assign(velocities, velocities + 3 * na);
    else {
        unsigned long long sdm = seed ? seed : 0x9E3779B97F4A7C15ull;
        auto next = [&]() {   // splitmix64
            sdm += 0x9E3779B97F4A7C15ull;
            unsigned long long z = sdm;
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
            return z ^ (z >> 31);
        };
        for (int im = 0; im < nmol; ++im) {
            const MolDesc& m = I.sys.mol[im];
            const double f = thermostat ? 1.0 : 2.0;
            const double eperat = f * temperature * 0.5 * kBoltz * nfree[im] / (3.0 * m.nat);
            for (int i = 0; i < m.nat; ++i) {
                const size_t a = m.at0 + i;
                const double v = std::sqrt(2.0 * eperat / mass[a]) * (I.sys.atoms[a].Z == 1 ? 2.0 : 1.0);
                for (int k = 0; k < 3; ++k) velo[3 * a + k] = (next() >> 63) ? -v : v;
            }
        }
    }
    MdDev& md = I.md;
    if (!I.md_ready) {
        md.velo = I.dalloc<double>(3 * na); md.vel = I.dalloc<double>(3 * na); md.mass = I.dalloc<double>(na);
        md.xyzo = I.dalloc<double>(3 * na); md.acc = I.dalloc<double>(3 * na); md.xyzt = I.dalloc<double>(3 * na);
        md.out = I.dalloc<double>(4 * (size_t)nmol);
    }
    // constraint tables change with every md_init: the ones of an earlier call are released first
    XTB_CUDA_CHECK(cudaStreamSynchronize(I.stream));
    for (void* p : I.md_allocs) {
        auto it = std::find(I.allocs.begin(), I.allocs.end(), p);
        if (it != I.allocs.end()) { I.allocs.erase(it); cudaFree(p); }
    }
    I.md_allocs.clear();
    const size_t first_md_alloc = I.allocs.size();
    md.cons0 = I.upload(cons0); md.cons = I.upload(cons); md.dist2 = I.upload(dist2);
    md.dro = I.dalloc<double>(3 * dist2.size()); md.gc = I.dalloc<double>(dist2.size());
    md.adj0 = I.upload(adj0); md.adj = I.upload(adj); md.nfree = I.upload(nfree);
    I.md_allocs.assign(I.allocs.begin() + first_md_alloc, I.allocs.end());
    XTB_CUDA_CHECK(cudaMemcpyAsync(md.velo, velo.data(), velo.size() * sizeof(double), cudaMemcpyHostToDevice, I.stream));
    XTB_CUDA_CHECK(cudaMemcpyAsync(md.mass, mass.data(), mass.size() * sizeof(double), cudaMemcpyHostToDevice, I.stream));
    XTB_CUDA_CHECK(cudaStreamSynchronize(I.stream));
    md.tstep0 = tstep_fs; md.tstep = tstep_fs * fstoau; md.taut = 500.0; md.tinit = temperature;
    md.thermostat = thermostat ? 1 : 0; md.shake = shake_mode ? 1 : 0;
    I.md_out.assign(4 * (size_t)nmol, 0.0);
    I.md_ready = true; I.md_first = true;
}
long long Batch::md_constraints() const { return p_->md_ncons; }

int Batch::md_run(int nsteps, double* epot, double* ekin, double* temp) {
    Impl& I = *p_;
    if (!I.md_ready) throw std::runtime_error("md_run: md_init has not been called");
    const int nmol = I.d.nmol;
    if (I.md_first) { singlepoint(true); I.md_first = false; }   // the single point before the MD loop, dynamic.f90:383-385
    int done = 0;
    for (int k = 0; k < nsteps; ++k) {
        singlepoint(true);
        I.prof_begin(K_MD, I.stream);
        k_md_update<<<nmol, 256, 0, I.stream>>>(I.d, I.md);
        I.prof_end(K_MD, I.stream);
        I.k_launches[K_MD]++;
        I.launches++;
        XTB_CUDA_CHECK(cudaMemcpyAsync(I.md_out.data(), I.md.out, 4 * (size_t)nmol * sizeof(double), cudaMemcpyDeviceToHost, I.stream));
        XTB_CUDA_CHECK(cudaStreamSynchronize(I.stream));
        XTB_CUDA_CHECK(cudaGetLastError());
        I.prof_collect();
        bool bad = false;
        for (int i = 0; i < nmol; ++i) {
            const double* o = &I.md_out[4 * (size_t)i];
            if (epot) epot[(size_t)k * nmol + i] = o[0];
            if (ekin) ekin[(size_t)k * nmol + i] = o[1];
            if (temp) temp[(size_t)k * nmol + i] = o[2];
            if (!(o[0] == o[0]) || o[0] > 1.e5 || o[0] < -1.e5 || o[2] > 10000.0) bad = true;   // dynamic.f90:424-431
        }
        done = k + 1;
        if (bad) break;
    }
    return done;
}

template <int TG>
static void launch_jacobi(Dev& d, int nmol, int max_nao, bool sm, cudaStream_t st) {
    const int npairs = (max_nao + 1) / 2;
    int threads = std::min(1024, ((TG * npairs + 31) / 32) * 32);
    threads = std::max(threads, 32);
    if (sm) {
        size_t bytes = (size_t)max_nao * (max_nao | 1) * sizeof(double);
        k_jacobi<TG, true><<<nmol, threads, bytes, st>>>(d);
    } else {
        k_jacobi<TG, false><<<nmol, threads, 0, st>>>(d);
    }
}

// KL(class, launch): one kernel launch, counted, and bracketed by events when its class is being profiled
#define KL(id, ...)                 \
    do {                            \
        I.prof_begin(id, st);       \
        __VA_ARGS__;                \
        I.prof_end(id, st);         \
        I.k_launches[id]++;         \
        nl++;                       \
    } while (0)

int Batch::singlepoint(bool restart) {
    Impl& I = *p_;
    Dev& d = I.d;
    BatchSystem& s = I.sys;
    if (!I.positions_set) throw std::runtime_error("positions not set");
    XTB_CUDA_CHECK(cudaSetDevice(I.device));
    cudaStream_t st = I.stream;
    const int nmol = d.nmol;
    // thresholds, src/scf_module.F90:264-277
    d.intcut = std::max(20.0, 25.0 - 10.0 * std::log10(I.acc));
    d.neglect = 10.0e-9 * I.acc;
    d.scfconv = 1.e-6 * I.acc;
    d.qconv = 1.e-4 * I.acc;
    d.etemp = I.etemp;
    d.broydamp = 0.4;   // src/setparam.f90:226
    d.maxiter = I.maxiter;
    d.jb_inner = I.knob_jb_inner;
    d.jb_tol = I.knob_jb_tol;
    d.jb_qtol = I.knob_jb_qtol;
    if (d.cnt) XTB_CUDA_CHECK(cudaMemsetAsync(d.cnt, 0, sizeof(unsigned long long) * kNCnt * nmol, st));
    if (I.iterlog_on) {
        if (I.iterlog_rows < I.maxiter) { I.iterlog_buf = I.dalloc<double>(5 * (size_t)I.maxiter); I.iterlog_rows = I.maxiter; }
        d.iterlog = I.iterlog_buf;
    } else d.iterlog = nullptr;
    if (!I.ev_phase[0]) for (auto& e : I.ev_phase) cudaEventCreate(&e);
    if (!I.ev_comm[0]) for (auto& e : I.ev_comm) cudaEventCreate(&e);
    cudaEvent_t* ev = I.ev_phase;
    int nl = 0;
    const int B = 128;
    cudaEventRecord(ev[0], st);
    KL(K_SETUP, k_reset_state<<<(nmol + 127) / 128, 128, 0, st>>>(d, (restart && I.have_wfn) ? 1 : 0));
    // ---- geometry dependent setup
    KL(K_SETUP, k_cn<<<grid2((long long)kAS * s.max_nat, B, nmol), B, 0, st>>>(d));
    const int mx = std::max(s.max_nat, std::max(s.max_nsh, s.max_ndim));
    KL(K_SETUP, k_cn_dependent<<<grid2(mx, B, nmol), B, 0, st>>>(d));
    KL(K_SETUP, k_gab<<<grid2((long long)s.max_nat * s.max_nat, 256, nmol), 256, 0, st>>>(d));
    KL(K_SETUP, k_jmat<<<grid2((long long)s.max_nsh * s.max_nsh, 256, nmol), 256, 0, st>>>(d));
    KL(K_SETUP, k_dispmat<<<grid2((long long)s.max_ndim * s.max_ndim, 256, nmol), 256, 0, st>>>(d));
    KL(K_SETUP, k_repulsion<<<grid2((long long)kAS * s.max_nat, B, nmol), B, 0, st>>>(d));
    if (d.pc0) KL(K_SETUP, k_pcem_pot<<<grid2(s.max_nsh, B, nmol), B, 0, st>>>(d));
    if (d.solv_on) {   // solvation%update, src/scf_module.F90:299-320 (every rank of a split system computes all of it)
        XTB_CUDA_CHECK(cudaMemsetAsync(d.sv_dsdrt, 0, 3 * (size_t)s.tot_nn * sizeof(double), st));
        KL(K_SOLV_SETUP, k_solv_born<<<dim3(s.max_nat, nmol), 128, 0, st>>>(d));
        KL(K_SOLV_SETUP, k_solv_sasa<<<dim3(s.max_nat, nmol), 256, 0, st>>>(d));
        KL(K_SOLV_SETUP, k_solv_shape<<<nmol, 256, 0, st>>>(d));
        KL(K_SOLV_SETUP, k_solv_bornmat<<<grid2((long long)s.max_nat * s.max_nat, 256, nmol), 256, 0, st>>>(d));
    }
    // ---- guess (src/api/interface.f90:82-101: reuse the cached wavefunction when present)
    if (!(restart && I.have_wfn)) {
        if (I.guess == 0) {
            KL(K_EEQ, k_eeq_build<<<grid2((long long)(s.max_nat + 1) * (s.max_nat + 2), 256, nmol), 256, 0, st>>>(d));
            if (s.max_nat <= 320) {
                KL(K_EEQ, k_eeq_solve<<<nmol, 256, 0, st>>>(d));
            } else if (nmol == 1 && I.have_eig && (s.max_nat + 2) / I.eig.nsm + 2 <= kEeqRowMax && !tuning_env("XTBB200_EEQ_LAUNCHES")) {
                // one large molecule: the same elimination as one cooperative launch (grid barriers instead of 2 N launches)
                XTB_CUDA_CHECK(cudaMemsetAsync(I.eig.bar, 0, sizeof(unsigned), st));
                Dev dd = d;
                unsigned* barp = I.eig.bar;
                void* args[] = {&dd, &barp};
                I.prof_begin(K_EEQ, st);
                XTB_CUDA_CHECK(cudaLaunchCooperativeKernel((void*)k_eeq_lu_coop, dim3(I.eig.nsm), dim3(1024), args, 0, st));
                I.prof_end(K_EEQ, st);
                I.k_launches[K_EEQ] += 1; nl += 1;
                KL(K_EEQ, k_eeq_finish<<<nmol, 1024, 0, st>>>(d));
            } else {   // large molecules: the same elimination spread over the GPU, two launches per column
                for (int k = 0; k < s.max_nat + 1; ++k) {
                    KL(K_EEQ, k_eeq_pivot<<<nmol, 1024, 0, st>>>(d, k));
                    const long long el = (long long)(s.max_nat - k) * (s.max_nat + 1 - k);
                    if (el > 0) KL(K_EEQ, k_eeq_elim<<<grid2(el, 256, nmol), 256, 0, st>>>(d, k));
                }
                KL(K_EEQ, k_eeq_finish<<<nmol, 1024, 0, st>>>(d));
            }
        } else {
            KL(K_EEQ, k_guess_uniform<<<grid2(s.max_nat, B, nmol), B, 0, st>>>(d));
        }
    }
    KL(K_EEQ, k_qsh2qat<<<grid2(s.max_nat, B, nmol), B, 0, st>>>(d));
    // ---- integrals
    const long long max_pairs = (long long)s.max_nsh * (s.max_nsh + 1) / 2;
    // one large system split over GPUs: every rank evaluates an interleaved subset of the shell-pair CTAs into zeroed
    // matrices; the NCCL allreduce(sum) of disjoint blocks reproduces the single-GPU matrices bit for bit
    const int R = d.snranks;
    const long long pair_ctas = (max_pairs + kIntWarps - 1) / kIntWarps;
    const dim3 pair_grid((unsigned)std::max<long long>(1, (pair_ctas + R - 1) / R), nmol);
    const size_t tmb = (size_t)s.tot_mat * sizeof(double);
    if (R > 1) {
        XTB_CUDA_CHECK(cudaMemsetAsync(d.S, 0, tmb, st));
        XTB_CUDA_CHECK(cudaMemsetAsync(d.D, 0, 3 * tmb, st));
        XTB_CUDA_CHECK(cudaMemsetAsync(d.Q, 0, 6 * tmb, st));
        XTB_CUDA_CHECK(cudaMemsetAsync(d.H0, 0, tmb, st));
        XTB_CUDA_CHECK(cudaMemsetAsync(d.H0n, 0, tmb, st));
    }
    const long long pair_ctas2 = (max_pairs + kIntPairs - 1) / kIntPairs;   // k_sdqh0: half a warp per shell pair
    const dim3 pair_grid2((unsigned)std::max<long long>(1, (pair_ctas2 + R - 1) / R), nmol);
    KL(K_SDQH0, k_sdqh0<<<pair_grid2, kIntWarps * 32, kSdqSmem, st>>>(d));
    if (R > 1) {
        NcclApi& n = NcclApi::get();
        I.prof_begin(K_COMM, st);
        cudaEventRecord(I.ev_comm[0], st);
        n.check(n.GroupStart(), "ncclGroupStart");
        I.allreduce(d.S, (size_t)s.tot_mat, st);
        I.allreduce(d.D, 3 * (size_t)s.tot_mat, st);
        I.allreduce(d.Q, 6 * (size_t)s.tot_mat, st);
        I.allreduce(d.H0, (size_t)s.tot_mat, st);
        I.allreduce(d.H0n, (size_t)s.tot_mat, st);
        n.check(n.GroupEnd(), "ncclGroupEnd");
        cudaEventRecord(I.ev_comm[1], st);
        I.prof_end(K_COMM, st);
    }
    const long long nn2 = (long long)s.max_nao * s.max_nao;
    KL(K_DQMASK, k_dqmask<<<grid2(nn2, 256, nmol), 256, 0, st>>>(d));
    // ---- solver selection and overlap factorisation
    const bool use_blk = I.fits_blk && I.solver != 1;             // block Jacobi, matrix in shared memory
    const bool use_scalar = !use_blk && I.fits_scalar;            // scalar Jacobi, matrix in shared memory
    const bool hier = !use_blk && !use_scalar;                    // matrix in HBM
    const bool dc = hier && I.solver != 3;                        // Householder + divide and conquer (3: two-level Jacobi)
    bool bdc = false;                                             // the same for a batch of small matrices of one order
    // (a) solver mode 4 for ensemble-sized matrices (cross-check of the block Jacobi); (b) batches of mid-size matrices of one
    // order (conformers of a 60-150 atom molecule): instead of one matrix after the other through the GPU-wide solver, all
    // of them at once with one thread-block cluster per matrix in the tridiagonalisation
    const bool clu_batch = dc && nmol >= (I.knob_cluster_single ? 1 : 2) && (I.solver == 4 || (I.solver == 0 && !I.knob_no_cluster_eig));
    if ((use_blk && I.solver == 4) || clu_batch) {
        bool uniform = true;
        for (auto& m : s.mol) uniform = uniform && m.nao == s.max_nao;
        if (!I.have_eig) { I.eig.init(s.max_nao, I.device); I.have_eig = true; }
        const int nmaxb = use_blk ? kCtaEigMax : I.eig.cluster_max();
        if (uniform && s.max_nao > 1 && s.max_nao <= nmaxb) {
            if (!I.cta_eig) {
                I.eig.init_batch(nmol, s.max_nao);
                I.cta_eig = true;
            }
            bdc = true;
        }
    }
    const bool xtri = !(hier && I.solver == 2);                   // X = L^-T is upper triangular (dense when taken from eig(S))
    const bool sm = true;
    const int hj_nb = hj_nblocks(s.max_nao), hj_nbp = hj_nb / 2, hj_rt = (s.max_nao + GT - 1) / GT;
    const size_t hj_inner_bytes = jb_smem_bytes(HP, kHjInnerWarps), hj_apply_bytes = (size_t)HP * GLD * sizeof(double);
    // eigen-decomposition of Asrc (+ warm start from Zm when allowed) by the two-level block Jacobi: Zm, emo
    auto hjacobi = [&](const double* Asrc, int allow_warm, int kid) {
        KL(kid, k_hj_begin<<<(nmol + 127) / 128, 128, 0, st>>>(d));
        KL(kid, k_hj_sigma<<<nmol, 256, 0, st>>>(d, Asrc));
        KL(kid, k_hj_load<<<grid2((long long)s.max_nao * s.max_nao, 256, nmol), 256, 0, st>>>(d, Asrc, allow_warm));
        int sweep = 0, still = 1;
        for (; sweep < 40 && still; ++sweep) {
            for (int step = 0; step < hj_nb - 1; ++step) {
                KL(kid, k_hj_gram<<<dim3(4, hj_nbp, nmol), 128, 0, st>>>(d, step));
                KL(kid, k_hj_inner<<<dim3(hj_nbp, nmol), 32 * kHjInnerWarps, hj_inner_bytes, st>>>(d, step));
                KL(kid, k_hj_apply<<<dim3(hj_rt, hj_nbp, nmol), 128, hj_apply_bytes, st>>>(d, step));
            }
            XTB_CUDA_CHECK(cudaMemsetAsync(I.d_active, 0, sizeof(int), st));
            KL(kid, k_hj_sweep_end<<<(nmol + 127) / 128, 128, 0, st>>>(d, I.d_active));
            XTB_CUDA_CHECK(cudaMemcpyAsync(&still, I.d_active, sizeof(int), cudaMemcpyDeviceToHost, st));
            XTB_CUDA_CHECK(cudaStreamSynchronize(st));
        }
        KL(kid, k_hj_norms<<<grid2((long long)s.max_nao * 32, 256, nmol), 256, 0, st>>>(d));
        KL(kid, k_hj_rank<<<grid2(s.max_nao, 128, nmol), 128, 0, st>>>(d));
        KL(kid, k_hj_write<<<grid2((long long)s.max_nao * s.max_nao, 256, nmol), 256, 0, st>>>(d, still ? 1 : 0));
    };
    cudaEventRecord(ev[4], st);
    if (!hier && !(I.blocked_chol && !I.knob_cta_chol)) {
        const size_t chol_bytes = (size_t)s.max_nao * (s.max_nao | 1) * sizeof(double);
        KL(K_CHOL, k_chol_inv<true><<<nmol, 512, chol_bytes, st>>>(d));
    } else if (hier && I.solver == 2) {                           // cross-check path: X = V s^-1/2 from eig(S)
        hjacobi(d.S, 0, K_CHOL);                                  // S = V s V^T
        KL(K_CHOL, k_hj_xscale<<<grid2((long long)s.max_nao * s.max_nao, 256, nmol), 256, 0, st>>>(d));
    } else {                                                      // blocked Cholesky S = L L^T, X = L^-T (kernels_chol.cuh)
        XTB_CUDA_CHECK(cudaMemcpyAsync(d.W, d.S, tmb, cudaMemcpyDeviceToDevice, st));
        XTB_CUDA_CHECK(cudaMemsetAsync(d.X, 0, tmb, st));
        const int cnb = (s.max_nao + CB - 1) / CB;
        for (int k = 0; k < cnb; ++k) {
            KL(K_CHOL, k_bchol_diag<<<nmol, 256, kBcholDiagSmem, st>>>(d, k));
            const int mt = cnb - k - 1;
            if (mt > 0) {
                KL(K_CHOL, k_bchol_panel<<<dim3(mt, nmol), 128, 0, st>>>(d, k));
                KL(K_CHOL, k_bchol_update<<<dim3(mt * (mt + 1) / 2, nmol), 128, 0, st>>>(d, k));
            }
        }
        if (cnb > 1) KL(K_CHOL, k_bchol_scale<<<dim3(cnb * (cnb - 1) / 2, nmol), 128, 0, st>>>(d));
        KL(K_CHOL, k_bchol_inv_diag<<<dim3(cnb, nmol), 256, 0, st>>>(d));
        for (int i = 1; i < cnb; ++i) KL(K_CHOL, k_bchol_inv_row<<<dim3(i, nmol), 128, 0, st>>>(d, i));
    }
    cudaEventRecord(ev[1], st);
    // ---- SCC iterations
    // GEMM tile: 48 or 64, whichever pads the batch's matrices less (144 AOs: 3 x 48 exactly instead of 192 x 192)
    double pad48 = 0, pad64 = 0;
    for (const MolDesc& mm : s.mol) {
        const double t48 = 48.0 * ((mm.nao + 47) / 48), t64 = 64.0 * ((mm.nao + 63) / 64);
        pad48 += t48 * t48 * t48; pad64 += t64 * t64 * t64;
    }
    const bool gt48 = pad48 < 0.9 * pad64;
    const int gtile = gt48 ? 48 : 64;
    const int tiles = (s.max_nao + gtile - 1) / gtile;
    const dim3 ggrid(tiles * tiles, nmol);
#define KGEMM(id, TA, TB, ...)                                                          \
    do {                                                                                \
        if (gt48) KL(id, (k_gemm<TA, TB, 48><<<ggrid, 128, 0, st>>>(__VA_ARGS__)));   \
        else KL(id, (k_gemm<TA, TB, 64><<<ggrid, 128, 0, st>>>(__VA_ARGS__)));        \
    } while (0)
    // One large system on several GPUs: the column-parallel products of an SCC iteration (X^T H X, back-transformation, C = X Z,
    // density) are split by tile columns, every rank computes its block with the tile decomposition of the unsplit run (bit-
    // identical columns) and the blocks are gathered by grouped NCCL broadcasts.  X^T H X has triangular work (K stops at the
    // tile, W on the lower tiles): its split balances the cost, the others are even.
    // (from 2500 AOs on: below, the three gathers per iteration cost more than the split products save -- 552 AOs: 470 instead
    // of 453 ms of SCC on two GPUs, 1998 AOs: break-even; XTBB200_GEMM_SPLIT_MIN overrides the threshold for tests)
    int gsplit_min = 2500;
    if (const char* e = tuning_env("XTBB200_GEMM_SPLIT_MIN")) gsplit_min = std::atoi(e);
    const bool gsplit = d.snranks > 1 && nmol == 1 && dc && !I.knob_no_gemm_split && s.max_nao >= gsplit_min;
    std::vector<int> sp_xhx(d.snranks + 1, 0), sp_eq(d.snranks + 1, 0);
    if (gsplit) {
        std::vector<double> cost(tiles);
        double tot = 0.0;
        for (int tc = 0; tc < tiles; ++tc) {
            cost[tc] = (double)tiles * (tc + 1) + 0.5 * ((double)tiles * (tiles + 1) - (double)tc * (tc + 1));
            tot += cost[tc];
        }
        double acc = 0.0;
        int r = 1;
        for (int tc = 0; tc < tiles && r < d.snranks; ++tc) {
            acc += cost[tc];
            while (r < d.snranks && acc >= tot * r / d.snranks) sp_xhx[r++] = tc + 1;
        }
        for (; r < d.snranks; ++r) sp_xhx[r] = tiles;
        sp_xhx[d.snranks] = tiles;
        for (int q = 0; q <= d.snranks; ++q) sp_eq[q] = (int)((long long)tiles * q / d.snranks);
    }
    const int xb = gsplit ? sp_xhx[d.srank] : 0, xe = gsplit ? sp_xhx[d.srank + 1] : 1 << 30;
    const int eb = gsplit ? sp_eq[d.srank] : 0, ee = gsplit ? sp_eq[d.srank + 1] : 1 << 30;
    if (gsplit && !I.ev_gather[0]) for (auto& e : I.ev_gather) cudaEventCreate(&e);
    I.t_gather = 0;
    auto gather = [&](double* p, const std::vector<int>& split, int which) {
        I.prof_begin(K_COMM, st);
        cudaEventRecord(I.ev_gather[2 * which], st);
        I.gather_cols(p, s.max_nao, split, gtile, st);
        cudaEventRecord(I.ev_gather[2 * which + 1], st);
        I.prof_end(K_COMM, st);
    };
    const int npairs = (s.max_nao + 1) / 2;
    // ensemble-sized matrices: block Jacobi on the FP64 tensor path, matrix resident in shared memory
    const int jb_warps = std::max(1, std::min(kJbMaxWarps, jb_dims(s.max_nao).nb / 2));   // k_jacobi_blk: __launch_bounds__(32 * kJbMaxWarps)
    const size_t jb_bytes = jb_smem_bytes(s.max_nao, jb_warps);
    auto jacobi_blk = I.knob_jb_variant ? k_jacobi_blk<1> : k_jacobi_blk<0>;
    int active = nmol;
    int iters = 0;
    std::fill(I.h_done.begin(), I.h_done.end(), 0);
    for (int it = 0; it < I.maxiter && active > 0; ++it) {
        KL(K_POT, k_pot_shell<<<grid2(8LL * s.max_nsh, 256, nmol), 256, 0, st>>>(d));
        KL(K_POT, k_pot_aes<true><<<grid2(8LL * s.max_nat, 256, nmol), 256, 0, st>>>(d));
        KL(K_D4SCC, k_d4_zeta<<<grid2(s.max_ndim, B, nmol), B, 0, st>>>(d, 0));
        KL(K_D4SCC, k_d4_matvec<<<grid2(8LL * s.max_ndim, 256, nmol), 256, 0, st>>>(d));
        if (d.solv_on) KL(K_SOLV_SCC, k_solv_matvec<<<dim3((s.max_nat + 31) / 32, nmol), 256, 0, st>>>(d, 0));
        KL(K_D4SCC, k_d4_shift<<<grid2(s.max_nat, B, nmol), B, 0, st>>>(d));
        KL(K_FOCK, k_fock<<<grid2(nn2, 256, nmol), 256, 0, st>>>(d));
        KGEMM(K_GEMM_XHX, false, false, d, d.H, d.X, d.T, xtri ? 2 : 0, 1, xb, xe);   // T = H X
        KGEMM(K_GEMM_XHX, true, false, d, d.X, d.T, d.W, xtri ? 3 : 0, 1, xb, xe);    // W = X^T T
        if (gsplit) {
            gather(d.W, sp_xhx, 0);
            if (xtri) KL(K_GEMM_XHX, k_mirror_lower<<<(unsigned)((nn2 + 255) / 256), 256, 0, st>>>(d.W, s.max_nao, gtile));
        }
        if (!dc && !bdc && (it > 0 || (restart && I.have_wfn)))
            KGEMM(K_GEMM_XHX, false, false, d, d.W, d.Zm, d.T, 0, 2);   // T = A' Z_prev
        if (bdc) {
            I.prof_begin(K_JACOBI, st);
            k_eig_skip_b<<<(nmol + 127) / 128, 128, 0, st>>>(d, I.eig.bskip);
            const int nla = I.eig.solve_batch(st, d.W, d.emo, d.Zm) + 2;
            k_eigh_finish<<<(nmol + 127) / 128, 128, 0, st>>>(d, I.eig.bfail, 1);
            I.prof_end(K_JACOBI, st);
            I.k_launches[K_JACOBI] += nla; nl += nla;
        } else if (use_blk && it == 0 && I.has_followers && !(restart && I.have_wfn) && I.solver == 0) {
            // first diagonalisation of an ensemble: leaders cold, followers warm from their leader's eigenvectors
            KL(K_JACOBI, jacobi_blk<<<nmol, 32 * jb_warps, jb_bytes, st>>>(d, 1));
            KL(K_JACOBI, k_lead_copy<<<grid2(nn2, 256, nmol), 256, 0, st>>>(d));
            KGEMM(K_GEMM_XHX, false, false, d, d.W, d.Zm, d.T, 0, 3);
            KL(K_JACOBI, jacobi_blk<<<nmol, 32 * jb_warps, jb_bytes, st>>>(d, 2));
        } else if (use_blk) KL(K_JACOBI, jacobi_blk<<<nmol, 32 * jb_warps, jb_bytes, st>>>(d, 0));
        else if (dc) {
            XTB_CUDA_CHECK(cudaMemsetAsync(I.eig.fail, 0, sizeof(int), st));
            // stage hooks of the solver -> kernel classes (stage 0: panel kernel, 1: trailing update, 2: D&C, 3: back-transformation)
            auto hook = [&](int stage, int end, int launches) {
                const int id = K_EIG_PANEL + stage;
                if (!end) I.prof_begin(id, st);
                else { I.prof_end(id, st); I.k_launches[id] += launches; nl += launches; }
            };
            for (int im = 0; im < nmol; ++im)
                if (!I.h_done[im])
                    I.eig.solve(st, d.W + s.mol[im].mat0, s.mol[im].nao, d.emo + s.mol[im].ao0, d.Zm + s.mol[im].mat0, false, hook,
                                gsplit ? eb * gtile : 0, gsplit ? ee * gtile : -1);
            KL(K_EIG_BT, k_eigh_finish<<<(nmol + 127) / 128, 128, 0, st>>>(d, I.eig.fail, 0));
        } else if (hier) hjacobi(d.W, 1, K_JACOBI);
        else if (npairs * 32 <= 1024) KL(K_JACOBI, launch_jacobi<32>(d, nmol, s.max_nao, sm, st));
        else if (npairs * 16 <= 1024) KL(K_JACOBI, launch_jacobi<16>(d, nmol, s.max_nao, sm, st));
        else if (npairs * 8 <= 1024) KL(K_JACOBI, launch_jacobi<8>(d, nmol, s.max_nao, sm, st));
        else KL(K_JACOBI, launch_jacobi<4>(d, nmol, s.max_nao, sm, st));
        KGEMM(K_GEMM_C, false, false, d, d.X, d.Zm, d.C, xtri ? 4 : 0, 1, eb, ee);   // C = X Z
        if (gsplit) gather(d.C, sp_eq, 1);
        KL(K_FERMI, k_fermi<<<nmol, 32, 0, st>>>(d));
        KL(K_DENSITY, k_scale_cols<<<grid2(nn2, 256, nmol), 256, 0, st>>>(d, 0));                // T = C f
        KGEMM(K_DENSITY, false, true, d, d.T, d.C, d.P, 1, 1, eb, ee);       // P = T C^T
        if (gsplit) gather(d.P, sp_eq, 2);
        KL(K_POP, k_pop_rows<<<grid2((long long)s.max_nao * 32, 256, nmol), 256, 0, st>>>(d));
        KL(K_POP, k_pop_atoms<<<grid2(s.max_nat, 64, nmol), 64, 0, st>>>(d));
        KL(K_ENERGY, k_es_rows<<<grid2(8LL * s.max_nsh, 256, nmol), 256, 0, st>>>(d));
        KL(K_ENERGY, k_aes_rows<<<grid2((long long)kAS * s.max_nat, 64, nmol), 64, 0, st>>>(d));
        KL(K_D4SCC, k_d4_zeta<<<grid2(s.max_ndim, B, nmol), B, 0, st>>>(d, 1));
        KL(K_D4SCC, k_d4_matvec<<<grid2(8LL * s.max_ndim, 256, nmol), 256, 0, st>>>(d));
        if (d.solv_on) KL(K_SOLV_SCC, k_solv_matvec<<<dim3((s.max_nat + 31) / 32, nmol), 256, 0, st>>>(d, 1));
        KL(K_SCCFIN, k_scc_finish<<<nmol, 256, 0, st>>>(d));
        XTB_CUDA_CHECK(cudaMemsetAsync(I.d_active, 0, sizeof(int), st));
        KL(K_SCCFIN, k_count_active<<<(nmol + 127) / 128, 128, 0, st>>>(d, I.d_active));
        XTB_CUDA_CHECK(cudaMemcpyAsync(&active, I.d_active, sizeof(int), cudaMemcpyDeviceToHost, st));
        if (dc && nmol > 1)
            XTB_CUDA_CHECK(cudaMemcpy2DAsync(I.h_done.data(), sizeof(int), &d.st[0].done, sizeof(MolState), sizeof(int), nmol,
                                             cudaMemcpyDeviceToHost, st));
        XTB_CUDA_CHECK(cudaStreamSynchronize(st));
        if (gsplit)
            for (int w = 0; w < 3; ++w) {
                float ms = 0;
                cudaEventElapsedTime(&ms, I.ev_gather[2 * w], I.ev_gather[2 * w + 1]);
                I.t_gather += ms;
            }
        iters++;
    }
    cudaEventRecord(ev[2], st);
    // ---- gradient
    KL(K_GRAD_H0, k_scale_cols<<<grid2(nn2, 256, nmol), 256, 0, st>>>(d, 1));                     // T = C (f e)
    KGEMM(K_GRAD_H0, false, true, d, d.T, d.C, d.H, 1, 0);            // H <- energy weighted density
    KL(K_GRAD_AES, k_pot_aes<false><<<grid2(8LL * s.max_nat, 256, nmol), 256, 0, st>>>(d));
    if (R > 1) {   // partial gradients: repulsion stays on rank 0, unowned shell pairs contribute zeros
        if (d.srank != 0) XTB_CUDA_CHECK(cudaMemsetAsync(d.grad, 0, 3 * s.atoms.size() * sizeof(double), st));
        XTB_CUDA_CHECK(cudaMemsetAsync(d.gpair, 0, 4 * (size_t)I.pair_total * sizeof(double), st));
    }
    {
        const long long pc2 = (max_pairs + kIntPairs - 1) / kIntPairs;   // half a warp per shell pair
        const dim3 pg2((unsigned)std::max<long long>(1, (pc2 + d.snranks - 1) / d.snranks), nmol);
        KL(K_GRAD_H0, k_grad_h0<<<pg2, kIntWarps * 32, kGradSmem, st>>>(d));
    }
    KL(K_GRAD_H0, k_grad_h0_collect<<<grid2((long long)kAS * s.max_nat, 64, nmol), 64, 0, st>>>(d));
    KL(K_GRAD_AES, k_grad_aes<<<grid2((long long)kAS * s.max_nat, 64, nmol), 64, 0, st>>>(d));
    KL(K_GRAD_D4, k_d4_weights<<<grid2(s.max_nat, 64, nmol), 64, 0, st>>>(d));
    KL(K_GRAD_D4, k_d4_c6<<<grid2((long long)s.max_nat * s.max_nat, 128, nmol), 128, 0, st>>>(d));
    KL(K_GRAD_D4, k_grad_pairs<<<grid2((long long)kAS * s.max_nat, 64, nmol), 64, 0, st>>>(d));
    KL(K_GRAD_ATM, k_grad_atm<<<dim3(s.max_nat, nmol), 128, (size_t)s.max_nat * sizeof(double), st>>>(d));
    if (d.solv_on) {   // solvation%addGradient + getEnergyParts with the mixed charges, src/scf_module.F90:759-777
        KL(K_SOLV_GRAD, k_solv_matvec<<<dim3((s.max_nat + 31) / 32, nmol), 256, 0, st>>>(d, -1));
        KL(K_SOLV_GRAD, k_solv_energy<<<nmol, 256, 0, st>>>(d));
        KL(K_SOLV_GRAD, k_solv_grad_pairs<<<dim3(s.max_nat, nmol), 64, 0, st>>>(d));
        KL(K_SOLV_GRAD, k_solv_grad_chain<<<dim3(s.max_nat, nmol), 64, 0, st>>>(d));
    }
    if (d.pc0) {
        KL(K_GRAD_MISC, k_grad_pcem_atoms<<<grid2(s.max_nat, 64, nmol), 64, 0, st>>>(d));
        KL(K_GRAD_MISC, k_grad_pcem_charges<<<grid2(std::max(I.max_npc, 1), 64, nmol), 64, 0, st>>>(d));
    }
    const int na_tot = (int)s.atoms.size();
    if (R > 1) {   // dE/dCN vectors and ATM energies: owned / partial -> complete on every rank
        I.prof_begin(K_COMM, st);
        cudaEventRecord(I.ev_comm[2], st);
        k_shard_pack<<<(na_tot + 127) / 128, 128, 0, st>>>(d, I.xbuf, na_tot);
        I.allreduce(I.xbuf, 4 * (size_t)na_tot, st);
        k_shard_unpack<<<(na_tot + 127) / 128, 128, 0, st>>>(d, I.xbuf, na_tot);
        I.prof_end(K_COMM, st);
        nl += 2;
    }
    KL(K_GRAD_MISC, k_grad_cn<<<grid2((long long)kAS * s.max_nat, 64, nmol), 64, 0, st>>>(d));
    if (R > 1) {
        I.prof_begin(K_COMM, st);
        I.allreduce(d.grad, 3 * (size_t)na_tot, st);
        cudaEventRecord(I.ev_comm[3], st);
        I.prof_end(K_COMM, st);
    }
    // ---- properties: bond orders and dipole
    bool any_open = false;
    for (auto& m : s.mol) any_open |= (m.nopen != 0);
    if (!any_open) {
        KGEMM(K_WBO, false, false, d, d.P, d.S, d.W, 0, 0);
    } else {
        // alpha / beta densities (scf_module.F90:869-879); closed-shell members of a mixed batch take this path too
        KL(K_WBO, k_scale_cols<<<grid2(nn2, 256, nmol), 256, 0, st>>>(d, 2));
        KGEMM(K_WBO, false, true, d, d.T, d.C, d.X, 0, 0);
        KGEMM(K_WBO, false, false, d, d.X, d.S, d.W, 0, 0);
        KL(K_WBO, k_scale_cols<<<grid2(nn2, 256, nmol), 256, 0, st>>>(d, 3));
        KGEMM(K_WBO, false, true, d, d.T, d.C, d.X, 0, 0);
        KGEMM(K_WBO, false, false, d, d.X, d.S, d.H, 0, 0);   // not Zm: it is the warm start of the next call
    }
    KL(K_WBO, k_wbo<<<grid2((long long)s.max_nat * s.max_nat, 128, nmol), 128, 0, st>>>(d, any_open ? 1 : 0));
    KL(K_FINISH, k_finish<<<nmol, 256, 0, st>>>(d));
    cudaEventRecord(ev[3], st);
    XTB_CUDA_CHECK(cudaMemcpyAsync(I.h_state.data(), d.st, nmol * sizeof(MolState), cudaMemcpyDeviceToHost, st));
    XTB_CUDA_CHECK(cudaStreamSynchronize(st));
    XTB_CUDA_CHECK(cudaGetLastError());
    cudaEventElapsedTime(&I.t_setup, ev[0], ev[1]);
    cudaEventElapsedTime(&I.t_scc, ev[1], ev[2]);
    cudaEventElapsedTime(&I.t_grad, ev[2], ev[3]);
    cudaEventElapsedTime(&I.t_fact, ev[4], ev[1]);
    I.t_comm = 0;
    if (R > 1) {
        float a = 0, b = 0;
        cudaEventElapsedTime(&a, I.ev_comm[0], I.ev_comm[1]);
        cudaEventElapsedTime(&b, I.ev_comm[2], I.ev_comm[3]);   // includes k_grad_cn between the two exchanges
        I.t_comm = a + b + I.t_gather;
    }
    I.prof_collect();
    I.launches = nl;
    I.scc_passes = iters;
    // The cached wavefunction seeds the next restarted call (MD, repeated xtb_singlepoint on one results handle): keep it
    // only when no molecule broke down (error code or non-finite energy); a molecule that merely ran out of iterations
    // keeps its last charges like the reference does (src/scf_module.F90:806-808 raises after storing the results).
    int nfail = 0;
    bool wfn_ok = true;
    for (int i = 0; i < nmol; ++i) {
        const MolState& hs = I.h_state[i];
        if (hs.fail || !hs.converged) nfail++;
        if (hs.fail || !(hs.energy == hs.energy) || std::fabs(hs.energy) > 1.0e300) wfn_ok = false;
    }
    I.have_wfn = wfn_ok;
    return nfail;
}

int symmetric_eigen(int n, const double* a_host, double* w_host, double* z_host, float* ms3, int device) {
    if (n < 1 || !a_host || !w_host) throw std::runtime_error("symmetric_eigen: invalid arguments");
    if (device_count() < 1) throw std::runtime_error("xtb_b200: no CUDA device available (this engine has no CPU fallback)");
    XTB_CUDA_CHECK(cudaSetDevice(device));
    EighWork w;
    cudaStream_t st = nullptr;
    double *A = nullptr, *Z = nullptr, *ev = nullptr;
    int fail = 0;
    const size_t n2 = (size_t)n * n;
    try {
        w.init(n, device);
        XTB_CUDA_CHECK(cudaStreamCreate(&st));
        A = w.alloc<double>(n2); Z = w.alloc<double>(n2); ev = w.alloc<double>(n);
        XTB_CUDA_CHECK(cudaMemcpyAsync(A, a_host, n2 * sizeof(double), cudaMemcpyHostToDevice, st));
        w.solve(st, A, n, ev, Z, true, [](int, int, int) {});
        XTB_CUDA_CHECK(cudaMemcpyAsync(w_host, ev, n * sizeof(double), cudaMemcpyDeviceToHost, st));
        if (z_host) XTB_CUDA_CHECK(cudaMemcpyAsync(z_host, Z, n2 * sizeof(double), cudaMemcpyDeviceToHost, st));
        XTB_CUDA_CHECK(cudaMemcpyAsync(&fail, w.fail, sizeof(int), cudaMemcpyDeviceToHost, st));
        XTB_CUDA_CHECK(cudaStreamSynchronize(st));
        XTB_CUDA_CHECK(cudaGetLastError());
        if (ms3) for (int k = 0; k < 3; ++k) ms3[k] = w.ms[k];
    } catch (...) {
        w.release();
        if (st) cudaStreamDestroy(st);
        throw;
    }
    w.release();
    cudaStreamDestroy(st);
    return fail;
}

void Batch::set_profile(unsigned long long mask) {
    Impl& I = *p_;
    I.prof_mask = mask;
    for (int k = 0; k < K_COUNT; ++k) { I.k_ms[k] = 0; I.k_calls[k] = 0; I.k_launches[k] = 0; }
}
int Batch::kernel_classes() { return K_COUNT; }
const char* Batch::kernel_class_name(int id) { return (id >= 0 && id < K_COUNT) ? kKernelNames[id] : ""; }
void Batch::kernel_times(float* ms, int* calls, int* launches) const {
    for (int k = 0; k < K_COUNT; ++k) {
        if (ms) ms[k] = p_->k_ms[k];
        if (calls) calls[k] = p_->k_calls[k];
        if (launches) launches[k] = p_->k_launches[k];
    }
}
void Batch::mark_call(int end) {
    Impl& I = *p_;
    XTB_CUDA_CHECK(cudaSetDevice(I.device));
    if (!I.ev_call[0]) for (auto& e : I.ev_call) cudaEventCreate(&e);
    cudaEventRecord(I.ev_call[end ? 1 : 0], I.stream);
    if (end) {
        cudaEventSynchronize(I.ev_call[1]);
        cudaEventElapsedTime(&I.t_call, I.ev_call[0], I.ev_call[1]);
    }
}
float Batch::call_ms() const { return p_->t_call; }
int Batch::scc_passes() const { return p_->scc_passes; }

int Batch::launches() const { return p_->launches; }
void Batch::timings(float* t3) const { t3[0] = p_->t_setup; t3[1] = p_->t_scc; t3[2] = p_->t_grad; }
float Batch::factorisation_ms() const { return p_->t_fact; }

void Batch::set_iteration_log(bool on) { p_->iterlog_on = on; }

int Batch::iteration_log(double* out, int max_rows) const {
    const Impl& I = *p_;
    if (!I.iterlog_on || !I.d.iterlog) return 0;
    XTB_CUDA_CHECK(cudaSetDevice(I.device));
    XTB_CUDA_CHECK(cudaStreamSynchronize(I.stream));
    const int rows = std::min(max_rows, std::min(result(0).niter, I.iterlog_rows));
    if (rows > 0) XTB_CUDA_CHECK(cudaMemcpy(out, I.d.iterlog, 5 * (size_t)rows * sizeof(double), cudaMemcpyDeviceToHost));
    return rows;
}

void Batch::set_decision_counting(bool on) {
    Impl& I = *p_;
    XTB_CUDA_CHECK(cudaSetDevice(I.device));
    if (on && !I.cnt_buf) I.cnt_buf = I.dalloc<unsigned long long>((size_t)kNCnt * I.d.nmol);
    I.d.cnt = on ? I.cnt_buf : nullptr;
}

void Batch::decision_counts(int i, long long* out) const {
    const Impl& I = *p_;
    if (!I.d.cnt) throw std::runtime_error("decision counting is off (xtbb200_setDecisionCounting)");
    if (i < 0 || i >= I.d.nmol) throw std::runtime_error("invalid molecule index");
    XTB_CUDA_CHECK(cudaSetDevice(I.device));
    unsigned long long h[kNCnt];
    XTB_CUDA_CHECK(cudaStreamSynchronize(I.stream));
    XTB_CUDA_CHECK(cudaMemcpy(h, I.d.cnt + (size_t)kNCnt * i, sizeof(h), cudaMemcpyDeviceToHost));
    // the pair kernels visit ordered pairs (i, j != i), the ATM kernel every triple from each of its three atoms
    static const int div[kNCnt] = {1, 1, 1, 1, 2, 2, 2, 2, 2, 3, 2, 1};
    for (int k = 0; k < kNCnt; ++k) out[k] = (long long)(h[k] / div[k]);
}

MolResult Batch::result(int i) const {
    const MolState& s = p_->h_state[i];
    MolResult r;
    r.energy = s.energy; r.e_elec = s.eel + s.embd; r.e_rep = s.ep; r.e_es = s.ees; r.e_aes = s.eaes; r.e_axc = s.epol;
    r.e_disp = s.ed + s.embd; r.hl_gap = s.egap; r.gnorm = s.gnorm; r.e_atm = s.embd;
    r.e_pcem = s.epcem;
    r.niter = s.iter; r.converged = s.converged; r.fail = s.fail; r.ndp = s.ndp; r.nqp = s.nqp; r.sweeps = s.sweeps_total;
    for (int k = 0; k < 3; ++k) r.dipole[k] = s.dipole[k];
    return r;
}

void Batch::fetch(const char* name, double* out) const {
    Impl& I = *p_;
    const Dev& d = I.d;
    const size_t na = I.sys.atoms.size(), ns = I.sys.shells.size(), nao = I.sys.ao2sh.size(), nr = I.sys.refs.size();
    const size_t tm = (size_t)I.sys.tot_mat, tn = (size_t)I.sys.tot_nn;
    struct E { const char* n; const double* p; size_t c; };
    const E tab[] = {
        {"gradient", d.grad, 3 * na}, {"q", d.q, na}, {"qsh", d.qsh, ns}, {"dipm", d.dipm, 3 * na}, {"qp", d.qp, 6 * na},
        {"cn", d.cn, na}, {"cn_cov", d.cn_cov, na}, {"cn_d4", d.cn_d4, na}, {"cn_erf", d.cn_erf, na}, {"radcn", d.radcn, na},
        {"emo", d.emo, nao}, {"focc", d.focc, nao}, {"S", d.S, tm}, {"dpint", d.D, 3 * tm}, {"qpint", d.Q, 6 * tm},
        {"H0", d.H0, tm}, {"H0_noovlp", d.H0n, tm}, {"P", d.P, tm}, {"C", d.C, tm}, {"X", d.X, tm}, {"wbo", d.wbo, tn},
        {"gab3", d.gab3, tn}, {"gab5", d.gab5, tn}, {"jmat", d.jmat, (size_t)I.sys.tot_ss},
        {"dispmat", d.dispmat, (size_t)I.sys.tot_dd}, {"gw", d.gw, nr}, {"selfenergy", d.selfE, ns}, {"xyz", d.xyz, 3 * na},
        {"vs", d.vs, na}, {"vd", d.vd, 3 * na}, {"vq", d.vq, 6 * na}, {"q_out", d.q_out, na}, {"qsh_out", d.qsh_out, ns},
        {"pcgradient", d.pc_grad, 3 * (size_t)external_charge_count()},
        {"velocities", I.md.velo, 3 * na}, {"md_mass", I.md.mass, na}, {"md_out", I.md.out, 4 * (size_t)d.nmol},
        {"brad", d.sv_brad, na}, {"sasa", d.sv_sasa, na}, {"bornmat", d.sv_bmat, tn}, {"solv_mol", d.sv_mol, (size_t)SVM_STRIDE * d.nmol},
    };
    if (std::strcmp(name, "solvation") == 0) {   // per molecule: g_solv, g_born, g_sasa, g_hb, g_shift
        if (!d.solv_on) throw std::runtime_error("no solvation model set");
        std::vector<double> tmp((size_t)SVM_STRIDE * d.nmol);
        XTB_CUDA_CHECK(cudaSetDevice(I.device));
        XTB_CUDA_CHECK(cudaMemcpyAsync(tmp.data(), d.sv_mol, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost, I.stream));
        XTB_CUDA_CHECK(cudaStreamSynchronize(I.stream));
        for (int i = 0; i < d.nmol; ++i) {
            const double* t = &tmp[(size_t)SVM_STRIDE * i];
            out[5 * i] = t[SVM_GSOLV]; out[5 * i + 1] = t[SVM_GBORN]; out[5 * i + 2] = t[SVM_GSASA]; out[5 * i + 3] = t[SVM_GHB];
            out[5 * i + 4] = t[SVM_GSHIFT];
        }
        return;
    }
    if (!I.md_ready && (!std::strcmp(name, "velocities") || !std::strcmp(name, "md_mass") || !std::strcmp(name, "md_out")))
        throw std::runtime_error("molecular dynamics has not been initialised");
    if (!d.solv_on && (!std::strcmp(name, "brad") || !std::strcmp(name, "sasa") || !std::strcmp(name, "bornmat") || !std::strcmp(name, "solv_mol")))
        throw std::runtime_error("no solvation model set");
    XTB_CUDA_CHECK(cudaSetDevice(I.device));
    for (const E& e : tab)
        if (std::strcmp(e.n, name) == 0) {
            XTB_CUDA_CHECK(cudaMemcpyAsync(out, e.p, e.c * sizeof(double), cudaMemcpyDeviceToHost, I.stream));
            XTB_CUDA_CHECK(cudaStreamSynchronize(I.stream));
            return;
        }
    throw std::runtime_error(std::string("unknown array: ") + name);
}
size_t Batch::fetch_len(const char* name) const {
    Impl& I = *p_;
    const size_t na = I.sys.atoms.size(), ns = I.sys.shells.size(), nao = I.sys.ao2sh.size(), nr = I.sys.refs.size();
    const size_t tm = (size_t)I.sys.tot_mat, tn = (size_t)I.sys.tot_nn;
    std::string n(name);
    if (n == "gradient" || n == "dipm" || n == "xyz" || n == "vd") return 3 * na;
    if (n == "qp" || n == "vq") return 6 * na;
    if (n == "q" || n == "cn" || n == "cn_cov" || n == "cn_d4" || n == "cn_erf" || n == "radcn" || n == "vs" || n == "q_out") return na;
    if (n == "qsh" || n == "selfenergy" || n == "qsh_out") return ns;
    if (n == "emo" || n == "focc") return nao;
    if (n == "S" || n == "H0" || n == "H0_noovlp" || n == "P" || n == "C" || n == "X") return tm;
    if (n == "dpint") return 3 * tm;
    if (n == "qpint") return 6 * tm;
    if (n == "wbo" || n == "gab3" || n == "gab5") return tn;
    if (n == "jmat") return (size_t)I.sys.tot_ss;
    if (n == "dispmat") return (size_t)I.sys.tot_dd;
    if (n == "gw") return nr;
    if (n == "pcgradient") return 3 * (size_t)external_charge_count();
    if (n == "velocities") return 3 * na;
    if (n == "md_mass") return na;
    if (n == "md_out") return 4 * (size_t)I.d.nmol;
    if (n == "brad" || n == "sasa") return na;
    if (n == "bornmat") return tn;
    if (n == "solvation") return 5 * (size_t)I.d.nmol;
    if (n == "solv_mol") return (size_t)SVM_STRIDE * I.d.nmol;
    throw std::runtime_error(std::string("unknown array: ") + name);
}
void Batch::fetch_mask(unsigned char* out) const {
    Impl& I = *p_;
    XTB_CUDA_CHECK(cudaSetDevice(I.device));
    XTB_CUDA_CHECK(cudaMemcpyAsync(out, I.d.mask, (size_t)I.sys.tot_mat, cudaMemcpyDeviceToHost, I.stream));
    XTB_CUDA_CHECK(cudaStreamSynchronize(I.stream));
}
const double* Batch::device_ptr(const char* name) const {
    const Dev& d = p_->d;
    std::string n(name);
    if (n == "gradient") return d.grad;
    if (n == "xyz") return d.xyz;
    if (n == "q") return d.q;
    throw std::runtime_error(std::string("unknown device array: ") + name);
}

}  // namespace xtbb200
