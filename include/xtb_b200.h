/* xtb_b200.h -- batched extension of the xtb C ABI for conformer ensembles on B200.
 *
 * The reference API (include/xtb.h) evaluates one molecule per call; callers that own an
 * ensemble (CREST-style conformer searches, the numerical Hessian loop of
 * src/type/calculator.f90:116-200) loop over xtb_singlepoint.  This extension submits the whole
 * ensemble in one call so that every kernel runs batched over molecules.  Plain pointers and
 * sizes only; no CUDA or torch types.  All functions return 0 on success, non-zero on error with
 * a message retrievable through xtbb200_lastError().
 */
#pragma once
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct _xtbb200_TBatch* xtbb200_TBatch;

/* last error message of the calling thread ("" if none) */
const char* xtbb200_lastError(void);

/* number of visible CUDA devices (0: the engine cannot run; there is no CPU fallback) */
int xtbb200_deviceCount(void);

/* Create a batch of nmol molecules on `device`.
 * natoms[nmol]; numbers: concatenated atomic numbers; charges/uhf [nmol] or NULL.
 * Replaces per-molecule xtb_newMolecule + xtb_newCalculator + xtb_loadGFN2xTB. */
xtbb200_TBatch xtbb200_newBatch(int nmol, const int* natoms, const int* numbers,
                                const double* charges, const int* uhf, int device);
void xtbb200_delBatch(xtbb200_TBatch* batch);

/* accuracy / electronic temperature [K] / max SCC iterations (xtb_setAccuracy, xtb_setElectronicTemp,
 * xtb_setMaxIter); guess: 0 = EEQ charges (reference default), 1 = uniform charges */
int xtbb200_setOptions(xtbb200_TBatch batch, double accuracy, double etemp, int maxiter, int guess);

/* Restart wavefunction of the whole batch (the reference's chk%wfn: shell charges qsh[nshell], atomic dipoles
 * dipm[natoms][3], quadrupoles qp[natoms][6], concatenated over molecules; src/type/restart.F90, used as SCC start by
 * src/scf_module.F90 when restart is requested).  NULL pointers drop the stored wavefunction (next call starts from
 * the EEQ guess).  The converged values are read back with xtbb200_getArray("qsh" | "dipm" | "qp"). */
int xtbb200_setWavefunction(xtbb200_TBatch batch, const double* qsh, const double* dipm, const double* qp);

/* External point charges for every molecule of the batch (batched xtb_setExternalCharges / xtb_releaseExternalCharges,
 * src/api/calculator.f90:470-580): npc[nmol] charges per molecule; numbers (element whose chemical hardness damps the
 * interaction, src/api/calculator.f90:533), charges, positions [npc][3] in Bohr, concatenated over molecules.
 * npc == NULL (or all zero) releases them.  After a single point the gradient on the charges is
 * xtbb200_getArray("pcgradient") (3 doubles per charge; xtb_getPCGradient). */
int xtbb200_setExternalCharges(xtbb200_TBatch batch, const int* npc, const int* numbers, const double* charges,
                               const double* positions);

/* Host-side query of the solvation model xtbb200_setSolvent / xtb_setSolvent would build (no GPU needed): out8 = dielectric
 * constant, keps = (1/eps - 1) / (1 + alpha beta), ALPB alpha * beta (0 for GBSA), free energy shift incl. the reference
 * state [Eh], Born radius scaling c1, Born offset [bohr], probe radius [bohr], 1.0 if the set has a hydrogen bond term.
 * Returns -1 (xtbb200_lastError: "solvent: '...' is not parametrized") for solvents without an internal GFN2 set. */
int xtbb200_solventParameters(const char* solvent, int alpb, int kernel, int state, double temperature, double* out8);

/* Host-side query (no GPU needed) of what xtbb200_mdInit sets up for one molecule: the SHAKE constraint pairs in the order
 * of the reference's init_shake (src/shake_module.f90:50-148; 0-based atom indices, pairs[2 * capacity]) from positions
 * [natoms][3] in bohr and the Wiberg bond orders wbo[natoms * natoms] (mode 2 only), and the masses in atomic units with
 * hydrogens at hmass amu (masses[natoms], may be NULL).  Returns the number of constraints, -1 on error. */
long long xtbb200_shakeList(int natoms, const int* numbers, const double* positions, const double* wbo, int mode, int hmass,
                            int* pairs, long long capacity, double* masses);

/* Molecular dynamics of every molecule of the batch as independent replicas (reference: src/dynamic.f90:140-702 md,
 * src/shake_module.f90 init_shake / do_shake, src/rmrottr.f90): leap-frog with Berendsen thermostat (thermostat != 0,
 * tau = 500 fs), SHAKE (shake_mode 0 off, 1 X-H bonds, 2 all bonds with a Wiberg bond order above 0.5; xtb's default is 2),
 * hydrogen mass hmass amu (0 = physical; xtb's default is 4) and removal of the overall translation / rotation every
 * step.  The potential is whatever the batch is set up with (solvent, external charges, accuracy).
 *   xtbb200_mdInit: positions [natoms][3] in bohr; runs the single point that provides the bond orders (restart != 0:
 *     from the stored wavefunction); velocities [natoms][3] in atomic units or NULL = the reference's start (every
 *     component +-sqrt(2 e / m), hydrogens doubled, signs drawn from `seed`).  xtb's defaults: tstep_fs 4, 298.15 K.
 *   xtbb200_mdRun: nsteps steps; positions, velocities and forces stay on the device, per step only 4 doubles per molecule
 *     are read back.  epot / ekin / temperature: [nsteps][nmol] (Hartree, Hartree, K), any may be NULL.  Returns the
 *     number of steps done (less than nsteps when a replica blew up), -1 on error.  Positions and velocities:
 *     xtbb200_getArray("xyz" | "velocities").
 *   xtbb200_mdConstraints: number of SHAKE constraints of the whole batch. */
int xtbb200_mdInit(xtbb200_TBatch batch, const double* positions, double tstep_fs, double temperature, int thermostat,
                   int shake_mode, int hmass, const double* velocities, unsigned long long seed, int restart);
int xtbb200_mdRun(xtbb200_TBatch batch, int nsteps, double* epot, double* ekin, double* temperature);
long long xtbb200_mdConstraints(xtbb200_TBatch batch);

/* Implicit solvation for every molecule of the batch (batched form of xtb_setSolvent / xtb_releaseSolvent,
 * reference: src/api/calculator.f90:341-476 -> src/solv/model.f90, gbsa.f90).  solvent: name as xtb accepts it ("h2o",
 * "water", "ch2cl2", "dmso" ...), NULL releases the model.  alpb: 0 = GBSA, 1 = ALPB.  kernel: 1 = Still (GBSA parameter
 * sets; what xtb_setSolvent(..., "gbsa") selects), 2 = P16 (ALPB parameter sets; xtb_setSolvent(..., "alpb")).
 * state: 1 gsolv, 2 reference, 3 mol1bar.  temperature in K (298.15).  ion_strength > 0 switches Debye-Hueckel salt
 * screening on (Still kernel), ion_rad in bohr.  Energy parts of the last single point: xtbb200_getArray("solvation") =
 * g_solv, g_born, g_sasa, g_hb, g_shift per molecule; Born radii "brad", atomic surfaces "sasa". */
int xtbb200_setSolvent(xtbb200_TBatch batch, const char* solvent, int alpb, int kernel, int state, double temperature,
                       double ion_strength, double ion_rad);

/* eigensolver selection: 0 = automatic (block Jacobi on the FP64 tensor path when the matrix fits in shared memory,
 * Householder tridiagonalisation + divide and conquer + back-transformation beyond: one matrix at a time with GPU-wide
 * kernels, or - batches of one order up to ~670 - all matrices at once with one thread-block cluster per matrix),
 * 1 = scalar one-sided Jacobi,
 * 2 = (matrices beyond shared memory) overlap orthogonaliser X = V s^-1/2 from the eigen-decomposition of S instead of
 * the blocked Cholesky X = L^-T, 3 = (matrices beyond shared memory) two-level block Jacobi instead of Householder +
 * divide and conquer, 4 = (ensembles whose matrices all have one order <= 160) batched Householder + divide and conquer
 * instead of the block Jacobi (an independent cross-check; measured 2.3x slower than the warm-started Jacobi at 144 AOs),
 * 5 = (matrices beyond shared memory) Householder + divide and conquer one matrix at a time also for batches of one
 * order (cross-check of the cluster-per-matrix path).
 * All give the same spectrum and density to rounding; used by the parity tests. */
int xtbb200_setSolver(xtbb200_TBatch batch, int mode);

/* The large-matrix symmetric eigensolver of the SCC loop on its own (what dsyevd is to the reference,
 * src/mctc/lapack/eigensolve.F90:240): a[n*n] column-major symmetric (both triangles, host memory) -> w[n] ascending,
 * z[n*n] eigenvectors (may be NULL).  ms3 (may be NULL): device time [ms] of tridiagonalisation, divide and conquer,
 * back-transformation.  Returns 0 on success, > 0 = number of unconverged secular equations, < 0 on error. */
int xtbb200_symmetricEigen(int n, const double* a, double* w, double* z, float* ms3, int device);

/* Split of ONE large system over the GPUs of a box (one process per GPU; every process creates the SAME batch on its
 * own device).  The shell-pair blocks of the integral build (build_SDQH0, src/xtb/hamiltonian.F90:261-412) and of its
 * derivative (build_dSDQH0, :575-792) are dealt to the ranks CTA by CTA, the per-atom gradient terms (AES, D4 two- and
 * three-body, Coulomb derivative, CN chain rule) in chunks of 32 atoms; the contributions are summed with NCCL
 * allreduce over NVLink: S/D/Q/H0 (12 n^2 doubles, disjoint blocks: bit-identical to one GPU), the dE/dCN vectors
 * (4 N) and the gradient (3 N).  In the SCC iterations the tridiagonalisation and the divide and conquer run replicated on
 * every rank; the column-parallel products (X^T H X, back-transformation of the eigenvectors, C = X Z, density) are split by
 * tile columns and their blocks gathered with grouped ncclBroadcast (bit-identical to one GPU), so every rank ends with the
 * complete result.
 *   xtbb200_shardUniqueId: rank 0 creates the 128-byte NCCL id; the caller ships it to the other ranks (MPI,
 *                          torch.distributed, a file ...).
 *   xtbb200_setShard:      collective over all ranks; nranks == 1 switches the split off again.
 *   xtbb200_shardOwner:    ownership rule (pure function): kind 0 = atom index, kind 1 = shell-pair CTA index. */
int xtbb200_shardUniqueId(char* id128, size_t len);
int xtbb200_setShard(xtbb200_TBatch batch, int rank, int nranks, const char* id128, size_t len);
int xtbb200_shardOwner(int kind, long long index, int nranks);
/* device time [ms] spent in the exchanges of the last call (0 when not split) */
int xtbb200_getCommTime(xtbb200_TBatch batch, float* ms);

/* GFN2-xTB single point + gradient for every molecule of the batch (batched xtb_singlepoint).
 * positions: concatenated [natoms][3] in Bohr (host memory). restart != 0 reuses the converged
 * wavefunction of the previous call as SCC start (src/api/interface.f90:82-101).
 * Outputs (host memory, may be NULL): energies[nmol] in Eh, gradients concatenated [natoms][3]
 * in Eh/Bohr, charges concatenated [natoms], niter[nmol], status[nmol] (0 = converged).
 * Returns 0 if every molecule converged, 1 if some did not, <0 on error. */
int xtbb200_singlepoint(xtbb200_TBatch batch, const double* positions, int restart,
                        double* energies, double* gradients, double* charges, int* niter, int* status);

/* same, with positions/gradients given as DEVICE pointers on the batch's device (no host copies) */
int xtbb200_singlepointDevice(xtbb200_TBatch batch, const double* positions_dev, int restart,
                              double* energies_dev_or_null);

/* sizes */
int xtbb200_getNMol(xtbb200_TBatch batch);
int xtbb200_getSizes(xtbb200_TBatch batch, int imol, int* natoms, int* nshell, int* nao);
/* number of valence electrons and of unpaired electrons (wfn%nel, wfn%nopen; header of the xtbrestart file,
 * src/restart.f90:190-192) */
int xtbb200_getElectrons(xtbb200_TBatch batch, int imol, int* nel, int* nopen);
size_t xtbb200_getDeviceBytes(xtbb200_TBatch batch);

/* per-molecule scalar results of the last call:
 * out[16] = energy, e_elec, e_rep, e_es, e_aes, e_axc, e_disp, hl_gap, gnorm, e_atm, dipole(3),
 *           niter, converged, fail */
int xtbb200_getResult(xtbb200_TBatch batch, int imol, double* out16);
/* What the reference prints but does not return through xtb.h: HOMO-LUMO gap [eV] (scc_results%hl_gap,
 * src/scf_module.F90:894) and the number of SCC iterations (src/scc_core.f90:628-638) of the single point stored on an
 * xtb.h results handle (xtb_TResults passed as void*).  Either output may be NULL. */
int xtbb200_getResultsInfo(void* xtb_results, double* hl_gap_eV, int* scc_iterations);
/* screening list lengths ndp, nqp and total Jacobi sweeps of one molecule */
int xtbb200_getCounts(xtbb200_TBatch batch, int imol, int* ndp, int* nqp, int* sweeps);
/* diagnostic (parity of the discrete decisions): when switched on, the kernels of the following
 * single points count every cutoff / screening decision where they take it.  counts12, per molecule, unordered pairs:
 * [0] shell pairs of build_SDQH0 with r^2 <= 2000 (src/intgrad.f90:904), [1] their primitive pairs with
 * alpha_i alpha_j r^2 <= intcut (alpha_i + alpha_j) (intgrad.f90:917-928), [2] off-site shell pairs of build_dSDQH0 with
 * r^2 <= 1600 (src/xtb/hamiltonian.F90:634,672), [3] their surviving primitive pairs (intgrad.f90:981-1004), [4] pairs of
 * the GFN / covalent coordination numbers (1e-12 <= r^2 <= 1600, src/disp/coordinationnumber.F90:427), [5] pairs of the
 * erf / D4 coordination numbers (r^2 <= 1600, src/disp/ncoord.f90:46,495,601), [6] repulsion pairs (1e-8 <= r^2 <= 1600,
 * src/xtb/repulsion.F90:107), [7] non-zero reference pairs of the D4 self-consistent matrix (r^2 <= 72^2 and
 * gw_i gw_j >= 1e-7, src/disp/dftd4.F90:674-713), [8] D4 gradient pairs (1e-10 <= r^2 <= 60^2, dftd4.F90:1894), [9] ATM
 * triples with all edges r^2 <= 40^2 (dftd4.F90:2126-2140), [10] bond-order pairs r^2 < 100 (src/scc_core.f90:1232),
 * [11] pairs of the Born radii whose psi term re-uses the preceding pair's 1/r (src/solv/born.f90:292-309, see
 * kernels_solv.cuh born_stale_r1; only with a solvation model).  Not defined for a batch split over several GPUs. */
int xtbb200_setDecisionCounting(xtbb200_TBatch batch, int on);
int xtbb200_getDecisionCounts(xtbb200_TBatch batch, int imol, long long* counts12);

/* copy a named device array of the whole batch to the host; names: gradient q qsh dipm qp cn emo
 * focc S dpint qpint H0 H0_noovlp P C wbo gab3 gab5 jmat dispmat gw xyz pcgradient ...; with a solvation model:
 * solvation (5 per molecule: g_solv g_born g_sasa g_hb g_shift) brad sasa bornmat; after xtbb200_mdInit: velocities
 * md_mass; len = number of doubles */
int xtbb200_getArrayLen(xtbb200_TBatch batch, const char* name, size_t* len);
int xtbb200_getArray(xtbb200_TBatch batch, const char* name, double* out, size_t len);
/* offsets of molecule imol inside the concatenated arrays: atoms, shells, AOs, nao^2, natoms^2 */
int xtbb200_getOffsets(xtbb200_TBatch batch, int imol, long long* off5);
/* dipole (bit 0) / quadrupole (bit 1) screening mask, one byte per AO pair, concatenated nao^2 */
int xtbb200_getMask(xtbb200_TBatch batch, unsigned char* out, size_t len);

/* device timings [ms] of the last call (setup+integrals, SCC, gradient+properties) and kernel launches */
int xtbb200_getTimings(xtbb200_TBatch batch, float* ms3, int* launches);
/* device time [ms] of the overlap factorisation S = L L^T, X = L^-T inside ms3[0].  The reference's "iterations" timer
 * slot (src/scf_module.F90:606-663) covers solver init + dpotrf + all SCC iterations, i.e. this time + ms3[1]. */
int xtbb200_getFactorisationTime(xtbb200_TBatch batch, float* ms);

/* device time [ms] of the last xtbb200_singlepoint call from the first host->device copy to the last device->host
 * copy (CUDA events on the engine's stream) */
int xtbb200_getCallTime(xtbb200_TBatch batch, float* ms);

/* Per-kernel-class device timers.  Every launch whose class bit is set in class_mask is bracketed by CUDA events
 * on the engine's own stream; times accumulate over xtbb200_singlepoint* calls until the next xtbb200_setProfile.
 * ms / timed_launches / launches: arrays of xtbb200_kernelClassCount() entries (may be NULL);
 * scc_passes: batch-wide SCC passes of the last call. */
int xtbb200_setProfile(xtbb200_TBatch batch, unsigned long long class_mask);
int xtbb200_kernelClassCount(void);
const char* xtbb200_kernelClassName(int id);
int xtbb200_getKernelTimes(xtbb200_TBatch batch, float* ms, int* timed_launches, int* launches, int* scc_passes);

#ifdef __cplusplus
}
#endif
