// Public C++ interface of the batched B200 GFN2-xTB engine (no CUDA types in the signatures).
#pragma once
#include <cstddef>

namespace xtbb200 {

struct MolResult {
    double energy, e_elec, e_rep, e_es, e_aes, e_axc, e_disp, hl_gap, gnorm, e_atm;
    double dipole[3];
    double e_pcem;   // interaction with external point charges (part of e_elec)
    int niter, converged, fail, ndp, nqp, sweeps;
};

int device_count();
// multi-GPU split of one large system: NCCL unique id (128 bytes) created on rank 0, ownership rules
void shard_unique_id(char* out128);
int shard_atom_owner_host(int iatom, int nranks);
int shard_pair_cta_owner_host(long long cta, int nranks);

// Eigen-decomposition of one dense symmetric matrix (column-major, both triangles) with the large-matrix solver of the
// SCC loop (Householder tridiagonalisation, divide and conquer, back-transformation; kernels_eigh.cuh): eigenvalues
// ascending in w[n], eigenvectors in z[n*n] (may be null).  ms3: device time of the three stages.  Returns the number
// of secular equations that did not converge (0 = success).
int symmetric_eigen(int n, const double* a_host, double* w_host, double* z_host, float* ms3, int device);

class Batch {
 public:
    // numbers: concatenated atomic numbers of all molecules; charges/uhf may be null (0)
    Batch(int nmol, const int* natoms, const int* numbers, const double* charges, const int* uhf, int device = 0);
    ~Batch();
    Batch(const Batch&) = delete;
    Batch& operator=(const Batch&) = delete;

    void set_options(double accuracy, double etemp, int maxiter);
    void set_guess(int mode);                      // 0: EEQ (default), 1: uniform charges
    void set_jacobi_stop(double qtol);             // quadratic stop of the block Jacobi solver (<= 0: default 1e-7); numerical
                                                   // differentiation of the results (xtb_hessian) wants it tighter
    void set_solver(int mode);                     // 0: automatic, 1: scalar Jacobi, 2: X from eig(S), 3: two-level Jacobi, 4: batched D&C, 5: D&C one matrix at a time
    // this process becomes rank `rank` of `nranks` working on the SAME molecules: shell-pair integral / gradient blocks
    // and the per-atom gradient terms are split, NCCL allreduce of the contributions; nranks == 1 switches it off
    void set_shard(int rank, int nranks, const char* nccl_unique_id128);
    int shard_rank() const;
    int shard_nranks() const;
    float comm_ms() const;                         // device time of the exchanges of the last call
    void set_positions(const double* xyz_host);    // concatenated (3, natoms) in bohr, host memory
    void set_positions_device(const double* xyz_dev);
    void set_wavefunction(const double* qsh, const double* dipm, const double* qp);
    // external point charges (xtb_setExternalCharges, src/api/calculator.f90:470-540): npc[nmol] charges per molecule,
    // concatenated element numbers (-> chemical hardness), charges, positions [npc][3] in bohr; all null / zero: release.
    // The gradient on the charges of the last single point is fetch("pcgradient").
    void set_external_charges(const int* npc, const int* numbers, const double* charges, const double* xyz);
    void release_external_charges();
    long long external_charge_count() const;
    // implicit solvation (xtb_setSolvent, src/api/calculator.f90:341-445): one model for all molecules of the batch.
    // kernel 1 = Still (GBSA parameter sets), 2 = P16 (ALPB parameter sets); state 1 gsolv, 2 reference, 3 mol1bar;
    // throws when the solvent has no internal GFN2-xTB parameter set.  fetch("solvation") gives per molecule
    // g_solv, g_born, g_sasa, g_hb, g_shift; fetch("brad") / fetch("sasa") the Born radii and atomic surfaces.
    void set_solvent(const char* solvent, bool alpb, int kernel, int state, double temperature, double ion_strength,
                     double ion_rad);
    void release_solvent();
    bool has_solvent() const;
    // Molecular dynamics of every molecule of the batch as independent replicas (src/dynamic.f90:140-702): leap-frog
    // with Berendsen thermostat (tau = 500 fs), SHAKE (0 off, 1 X-H bonds, 2 all bonds with a Wiberg bond order > 0.5,
    // src/shake_module.f90:50-148) and removal of overall translation / rotation; hydrogen mass hmass amu (0: physical).
    // md_init runs the single point that provides the bond orders; velocities [natoms][3] in atomic units, or null:
    // mdinitu (dynamic.f90:850-890) with random signs from `seed`.  md_run propagates nsteps steps, positions, velocities
    // and forces stay on the device; epot / ekin / temp: [nsteps][nmol] (Hartree, Hartree, K), any may be null.  Returns
    // the number of steps done (fewer than nsteps when a replica blew up: energy not finite or T > 10000 K).
    void md_init(const double* xyz_host, double tstep_fs, double temperature, bool thermostat, int shake_mode, int hmass,
                 const double* velocities, unsigned long long seed, bool restart);
    int md_run(int nsteps, double* epot, double* ekin, double* temp);
    long long md_constraints() const;
    void invalidate_wavefunction();
    bool has_wavefunction() const;
    // runs SP + gradient for every molecule; returns the number of molecules that failed / did not converge
    int singlepoint(bool restart);

    int nmol() const;
    int natoms(int i) const;
    int nao(int i) const;
    int nshell(int i) const;
    int nel(int i) const;      // valence electrons (src/scf_module.F90:334)
    int nopen(int i) const;    // unpaired electrons
    int atom_offset(int i) const;
    int shell_offset(int i) const;
    int ao_offset(int i) const;
    long long mat_offset(int i) const;
    long long nn_offset(int i) const;
    int total_atoms() const;
    int total_shells() const;
    int total_ao() const;
    size_t device_bytes() const;
    void* stream() const;
    int launches() const;
    void timings(float* t3) const;   // ms: setup+integrals, SCC, gradient+properties
    float factorisation_ms() const;  // part of the setup time: S = L L^T, X = L^-T (the reference times it with the iterations)
    // per-kernel-class device timers: launches whose class bit is set in mask are bracketed by CUDA events on the
    // engine's stream; times accumulate over singlepoint calls until the next set_profile
    void set_profile(unsigned long long mask);
    static int kernel_classes();
    static const char* kernel_class_name(int id);
    void kernel_times(float* ms, int* calls, int* launches) const;
    void mark_call(int end);         // events around a whole host-buffer call (H2D ... D2H)
    float call_ms() const;
    int scc_passes() const;          // batch-wide SCC passes of the last call (= max niter over molecules)

    MolResult result(int i) const;
    // diagnostic: count the discrete decisions (cutoffs, primitive screening) where the kernels take them during the
    // next single points (common.cuh: kNCnt counters per molecule, unordered pairs / triples)
    // per-iteration log of molecule 0 (energy, energy change, rms charge change, gap, Broyden weight): the columns of the
    // reference's SCC iteration table (src/scc_core.f90:621-622); iteration_log returns the number of rows written to
    // out[rows][5] by the last single point
    void set_iteration_log(bool on);
    int iteration_log(double* out, int max_rows) const;
    void set_decision_counting(bool on);
    void decision_counts(int i, long long* out) const;
    void fetch(const char* name, double* out) const;   // device -> host copy of a named array
    size_t fetch_len(const char* name) const;
    void fetch_mask(unsigned char* out) const;
    const double* device_ptr(const char* name) const;

 private:
    struct Impl;
    Impl* p_;
};

}  // namespace xtbb200
