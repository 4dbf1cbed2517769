#!/usr/bin/env python3
"""Benchmark of the GFN2-xTB single point + gradient hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA engine, one rank per GPU
  python bench.py --impl reference --steps K --warmup W     # the reference algorithm on the host cores (CPU oracle)

Workload (BASELINE.json configs[2]): an ensemble of 4096 conformers per GPU of the 47-atom drug
rivaroxaban (test/unit/molstock.f90:366-424 of the reference; S and Cl carry d shells), every Cartesian coordinate
displaced by N(0, 0.05 bohr) with numpy.random.default_rng(seed=i), fresh EEQ guess per conformer.  One "step" = GFN2
single point + gradient for every conformer of the ensemble.  The ensemble shards by molecule: every rank owns 4096
conformers (weak scaling), there is no data-path collective; ranks only meet at the timing barrier.

`value`  : molecules/s with the positions already resident in HBM (xtbb200_singlepointDevice), device time from CUDA
           events on the engine's stream, max over ranks.
`e2e`    : molecules/s through xtbb200_singlepoint with pinned HOST buffers: positions H2D and energies, gradients,
           charges, iteration counts D2H inside the timed region.
`secondary`: the default run also carries, nested under this key, one line each for the other BASELINE configs, every one
           with its own roofline / cpu_baseline / e2e: "water" (configs[3], the second half of the metric: s/SCC-iteration
           at 3000 atoms; with N ranks the integral / gradient blocks are split over the N GPUs and the line reports the
           phase split, the exchange time and whether the split energy is bit-identical to one GPU), "caffeine"
           (configs[0]), "protein" (configs[1]), "md" (configs[4]).  Their step counts are fixed and small, so the default
           run stays within a few minutes; `--secondary off` drops them.
`--workload water | single | md`: the same secondary workloads on their own with --steps / --warmup as given.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "GFN2 SP+gradient/sec (ensemble)"
UNIT = "molecules/s"


# ----------------------------------------------------------------------------------------------- workloads
def load_kat(name):
    with open(os.path.join(ROOT, "tests", "golden", "reference_kat.json")) as f:
        s = json.load(f)[name]
    s["numbers"] = np.array(s["numbers"], dtype=np.int32)
    s["xyz"] = np.array(s["xyz"], dtype=np.float64).reshape(-1, 3)
    return s


def make_ensemble(s, nmol, first=0, sigma=0.05):
    """conformer i = geometry + N(0, sigma) per coordinate, rng(seed=first+i)."""
    Z = np.asarray(s["numbers"], dtype=np.int32)
    x0 = np.asarray(s["xyz"], dtype=np.float64).reshape(-1, 3)
    pos = np.empty((nmol,) + x0.shape)
    for i in range(nmol):
        pos[i] = x0 + sigma * np.random.default_rng(first + i).standard_normal(x0.shape)
    return Z, pos.reshape(-1, 3)


def make_water_cluster(nwat, seed=3000, spacing=5.86):
    """nwat H2O on a jittered cubic lattice clipped to a sphere, random orientations."""
    rng = np.random.default_rng(seed)
    mono = np.array([[0.0, 0.0, -0.7344], [0.0, 1.4327, 0.3672], [0.0, -1.4327, 0.3672]])  # ~0.96 A, 104.5 deg (bohr)
    m = int(np.ceil((nwat * 6 / np.pi) ** (1 / 3))) + 2
    g = np.arange(m) - (m - 1) / 2
    pts = np.array([(a, b, c) for a in g for b in g for c in g]) * spacing
    order = np.argsort((pts ** 2).sum(1), kind="stable")[:nwat]
    xyz = []
    for p in pts[order]:
        q = rng.standard_normal(4)
        q /= np.linalg.norm(q)
        a, b, c, d_ = q
        R = np.array([[a * a + b * b - c * c - d_ * d_, 2 * (b * c - a * d_), 2 * (b * d_ + a * c)],
                      [2 * (b * c + a * d_), a * a - b * b + c * c - d_ * d_, 2 * (c * d_ - a * b)],
                      [2 * (b * d_ - a * c), 2 * (c * d_ + a * b), a * a - b * b - c * c + d_ * d_]])
        xyz.append(mono @ R.T + p + 0.3 * rng.standard_normal(3))
    Z = np.array([8, 1, 1] * nwat, dtype=np.int32)
    return Z, np.concatenate(xyz)


def make_peptide_cluster(s, ncopy, gap=8.0):
    """Stand-in for a protein-sized system (crambin coordinates are not available offline): ncopy copies
    of the 76-atom peptide pdb-4qxx (test/unit/molstock.f90:739-909) on a cubic grid, `gap` bohr between bounding boxes."""
    Z0 = np.asarray(s["numbers"], dtype=np.int32)
    x0 = np.asarray(s["xyz"], dtype=np.float64).reshape(-1, 3)
    x0 = x0 - x0.mean(axis=0)
    ext = x0.max(axis=0) - x0.min(axis=0) + gap
    m = int(np.ceil(ncopy ** (1 / 3)))
    cells = [(a, b, c) for a in range(m) for b in range(m) for c in range(m)][:ncopy]
    xyz = np.concatenate([x0 + np.array(c) * ext for c in cells])
    return np.tile(Z0, ncopy), xyz


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock and throttle reasons with nvidia-smi while the timed region runs."""

    def __init__(self, index):
        self.index = index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None

    def _query(self, fields):
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + fields, "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=10).stdout.strip()
        return [x.strip() for x in out.split(",")]

    def _loop(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap", "hw_power_brake_slowdown"]
        fields = "clocks.sm," + ",".join("clocks_throttle_reasons." + n for n in names)
        while not self._stop.is_set():
            try:
                v = self._query(fields)
                self.samples.append(float(v[0]))
                for n, x in zip(names, v[1:]):
                    if x.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        try:
            self.max_mhz = float(self._query("clocks.max.sm")[0])
        except Exception:
            self.max_mhz = None
        self._t = threading.Thread(target=self._loop, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------- CPU arm
def _oracle_worker(args):
    Z, xyz = args
    from oracle.oracle import Oracle
    o = Oracle(Z, xyz, threads=1)
    st = o.singlepoint()
    return st, o.info["niter"], o.energies["e_total"]


def cpu_ensemble_rate(s, nmol, cores, first=0):
    """molecules/s of the CPU oracle (restatement of the reference algorithm + host OpenBLAS), one molecule per core."""
    import multiprocessing as mp
    Z, pos = make_ensemble(s, nmol, first)
    x = pos.reshape(nmol, -1, 3)
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    os.environ["OMP_NUM_THREADS"] = "1"
    with mp.get_context("fork").Pool(cores) as pool:
        pool.map(_oracle_worker, [(Z, x[i]) for i in range(min(cores, nmol))])   # page in the library
        t0 = time.perf_counter()
        res = pool.map(_oracle_worker, [(Z, x[i]) for i in range(nmol)], chunksize=1)
        dt = time.perf_counter() - t0
    assert all(r[0] == 0 for r in res)
    return nmol / dt, dt, [r[1] for r in res], [r[2] for r in res]


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    """--impl reference: the reference's CPU algorithm for this path.  The Fortran reference cannot be built in this
    image (no Fortran compiler, un-vendored mctc-lib), so this arm times the CPU oracle port."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.oracle import build
    build()
    s = load_kat(args.molecule)
    cores = host_cores()
    per_step = 64 * cores   # ~5-10 s per step
    for w in range(args.warmup):
        cpu_ensemble_rate(s, cores, cores)
    rates, times = [], []
    for k in range(args.steps):
        r, dt, _, _ = cpu_ensemble_rate(s, per_step, cores, first=k * per_step)
        rates.append(r)
        times.append(dt)
    value = per_step * args.steps / sum(times)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "ensemble of %s conformers (47 atoms, 144 AOs), sigma=0.05 bohr, EEQ guess" % args.molecule,
                   "molecules_per_step": per_step},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d conformers per step, one single-threaded oracle process per core" % per_step},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- GPU arm
class Ctx:
    """One process per GPU (torchrun env); the NCCL process group is only the timing barrier / max-over-ranks plumbing
    (and, for the split of one large system, the carrier of rank 0's NCCL unique id)."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        if not torch.cuda.is_available():
            raise RuntimeError("bench.py: no CUDA device; the engine has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
        self._peak64 = None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def reduce(self, vals, op="max"):
        t = self.torch.tensor([float(v) for v in vals], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        return [float(x) for x in t.tolist()]

    def fp64_peak(self):
        if self._peak64 is None:
            self._peak64 = measure_fp64_peak(self.torch, self.dev)
        return self._peak64

    def close(self):
        if self.world > 1:
            self.barrier()
            self.dist.destroy_process_group()


def library_eigh_calibration(torch, dev, n, batch):
    """Calibration only, never on the product path: the vendor library's symmetric eigensolver (cuSOLVER through
    torch.linalg.eigh -- the routine family the reference's own GPU build calls, src/mctc/lapack/eigensolve.F90:159-172) on
    random symmetric FP64 matrices of the workload's order, cold start (it cannot use the SCC's previous eigenvectors).
    Returns ms per matrix."""
    try:
        g = torch.Generator(device=dev)
        g.manual_seed(7)
        a = torch.randn(batch, n, n, dtype=torch.float64, device=dev, generator=g)
        a = 0.5 * (a + a.transpose(1, 2))
        torch.linalg.eigh(a[:1])
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.linalg.eigh(a)
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / batch
    except Exception as ex:            # a missing library must not take the bench line down
        return {"error": "%s: %s" % (type(ex).__name__, ex)}


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def ncu_traffic(kernel, nao):
    """DRAM bytes per molecule of a kernel from the committed `ncu --set full` capture (profiles/traffic.json, written by
    tools/ncu_summary.py from dram__bytes_read.sum + dram__bytes_write.sum); None when there is no capture of this size."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        e = t.get(kernel)
        if e and int(e.get("nao", -1)) == int(nao):
            return e
    except Exception:
        pass
    return None


def measure_fp64_peak(torch, dev):
    """cuBLAS DGEMM 4096^3 on this GPU: the FP64 (DMMA) denominator that MEASURED_PEAKS.json does not carry."""
    n = 4096
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    torch.matmul(a, b)
    torch.cuda.synchronize(dev)
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize(dev)
        best = min(best, e0.elapsed_time(e1))
    del a, b
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


def run_engine(ctx, args):
    import ctypes

    import xtb_b200
    torch, dev, rank, local, world = ctx.torch, ctx.dev, ctx.rank, ctx.local, ctx.world
    s = load_kat(args.molecule)
    nmol = args.nmol
    Z, pos = make_ensemble(s, nmol, first=rank * nmol)
    nat = len(Z)
    batch = xtb_b200.BatchCalculator([Z] * nmol, device=local)
    if args.solvent:
        batch.set_solvent(args.solvent, model=args.solvent_model)
        args.no_cpu = True     # the cpu_baseline leg of this workload is the gas-phase ensemble
    sizes = batch.sizes(0)
    nao, nsh = sizes["nao"], sizes["nshell"]
    # pinned host buffers for the end-to-end leg; device-resident positions for the kernel leg
    pos_pin = torch.from_numpy(pos.copy()).pin_memory()
    pos_dev = pos_pin.to(dev)
    torch.cuda.synchronize(dev)

    def step_device():
        batch.singlepoint_device(pos_dev.data_ptr())
        t = batch.timings()
        return t["setup_ms"] + t["scc_ms"] + t["grad_ms"], t

    # -- warm-up (also a full per-kernel-class profile pass to find the dominant kernel)
    batch.set_profile("all")
    for _ in range(max(args.warmup, 1)):
        step_device()
    prof, passes = batch.kernel_times()
    top = max(prof, key=lambda k: prof[k]["ms"])
    share_all = {k: v["ms"] for k, v in prof.items()}
    tot_all = sum(share_all.values())
    batch.set_profile([top])

    # -- timed region 1: device-resident inputs
    clocks = ClockSampler(local)
    ctx.barrier()
    clocks.start()
    dev_ms, launches = [], 0
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        ms, t = step_device()
        dev_ms.append(ms)
        launches += t["launches"]
    torch.cuda.synchronize(dev)
    wall_dev = time.perf_counter() - t_wall0
    prof_top, passes = batch.kernel_times()
    res = [batch.result(i) for i in (0, nmol - 1)]
    batch.array("gradient")  # results stay readable after the timed region (D2H not timed in this leg)
    ctx.barrier()

    # -- timed region 2: end to end through the host-buffer C ABI
    e_pin = torch.empty(nmol, dtype=torch.float64).pin_memory()
    g_pin = torch.empty(nmol * nat * 3, dtype=torch.float64).pin_memory()
    q_pin = torch.empty(nmol * nat, dtype=torch.float64).pin_memory()
    it_pin = torch.empty(nmol, dtype=torch.int32).pin_memory()
    st_pin = torch.empty(nmol, dtype=torch.int32).pin_memory()
    L = xtb_b200.load_library()
    dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)

    def step_host():
        rc = L.xtbb200_singlepoint(batch._b, ctypes.cast(pos_pin.data_ptr(), dp), 0, ctypes.cast(e_pin.data_ptr(), dp),
                                   ctypes.cast(g_pin.data_ptr(), dp), ctypes.cast(q_pin.data_ptr(), dp),
                                   ctypes.cast(it_pin.data_ptr(), ip), ctypes.cast(st_pin.data_ptr(), ip))
        if rc < 0:
            raise RuntimeError(L.xtbb200_lastError().decode())
        return batch.call_ms()

    batch.set_profile(None)
    step_host()
    ctx.barrier()
    e2e_ms = [step_host() for _ in range(args.steps)]
    ctx.barrier()
    clk = clocks.stop()
    n_conv = int((st_pin.numpy() == 0).sum())
    mean_iter = float(it_pin.numpy().mean())

    # -- max over ranks
    t_dev, t_e2e, t_wall = ctx.reduce([sum(dev_ms), sum(e2e_ms), wall_dev * 1e3], "max")
    value = world * nmol * args.steps / (t_dev * 1e-3)
    e2e = world * nmol * args.steps / (t_e2e * 1e-3)
    if rank != 0:
        del batch
        return None

    # -- roofline of the dominant kernel
    k = prof_top[top]
    k_ms = k["ms"] / max(k["timed_launches"], 1)
    roof = kernel_roofline(ctx, top, k_ms, nmol, nao, nsh, nat, batch, load_peaks())
    roof["kernel"] = top
    roof["avg_launch_ms"] = k_ms
    roof["share_of_step"] = k["ms"] / sum(dev_ms)
    roof["share_all_kernels_warmup"] = {kk: round(v / tot_all, 4) for kk, v in sorted(share_all.items(), key=lambda x: -x[1])
                                        if v / tot_all >= 0.005}

    # -- CPU baseline on a bounded sample of the same workload
    cpu = None
    if not args.no_cpu and world == 1:   # reported at N=1 only
        from oracle.oracle import build
        build()
        cores = host_cores()
        n_cpu = min(nmol, 96 * cores)   # ~10-20 s of CPU work
        rate, dt, its, es = cpu_ensemble_rate(s, n_cpu, cores, first=0)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "first %d conformers of the same ensemble, one single-threaded oracle process per core, %.1f s" % (n_cpu, dt),
               "max_abs_dE_vs_gpu": float(np.max(np.abs(np.array(es) - e_pin.numpy()[:n_cpu]))),
               "niter_identical": bool(np.array_equal(np.array(its), it_pin.numpy()[:n_cpu]))}

    h2d = int(pos_pin.numel() * 8)
    d2h = int((e_pin.numel() + g_pin.numel() + q_pin.numel()) * 8 + (it_pin.numel() + st_pin.numel()) * 4)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "ensemble of %d %s conformers per GPU (%d atoms, %d shells, %d AOs), sigma=0.05 bohr, "
                               "EEQ guess, SP+gradient%s" % (nmol, args.molecule, nat, nsh, nao,
                                                             (", %s(%s)" % (args.solvent_model.upper(), args.solvent)) if args.solvent else ""),
                   "molecules_per_gpu": nmol, "parallelism": "molecule-sharded x%d, no collective" % world,
                   "l2": "per-step working set %.1f GB of HBM-resident matrices >> 126 MB L2 (no flush needed)" % (batch.device_bytes() / 1e9),
                   "mean_scc_iterations": mean_iter, "scc_passes": passes, "converged": "%d/%d" % (n_conv, nmol)},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": t_e2e / args.steps},
        "gpu_launches": launches, "wall_ms_per_step_device_leg": t_wall / args.steps,
        "roofline": roof, "cpu_baseline": cpu, "clocks": clk,
        "sample_energy_Eh": res[0]["energy"],
    }
    del batch
    if rank == 0:
        torch.cuda.empty_cache()
        ms = library_eigh_calibration(torch, dev, nao, 256)
        line["calibration"] = {
            "what": "cuSOLVER symmetric eigensolver (torch.linalg.eigh) on 256 random symmetric FP64 matrices of %d x %d, cold "
                    "start; calibration only, not on the product path" % (nao, nao),
            "library_ms_per_matrix": ms,
            "engine_ms_per_matrix_and_diagonalisation": (roof.get("avg_launch_ms") / nmol) if roof.get("avg_launch_ms") else None}
    return line


def kernel_roofline(ctx, top, k_ms, nmol, nao, nsh, nat, batch, peaks):
    """Algorithmic bytes / flops per launch of the dominant kernel class (DESIGN.md, section "Kernels")."""
    hbm = peaks.get("hbm_gbs")
    n = nao
    if "jacobi" in top or "gemm" in top or "density" in top:
        # FP64 arithmetic: eigensolver = reduction to standard form is separate; Jacobi counted as the textbook
        # 4/3 n^3 (tridiagonalisation) + 2 n^3 (back-transformation) flop of the reference's dsyevd path
        sweeps = np.mean([batch.counts(i)["sweeps"] for i in range(0, nmol, max(1, nmol // 64))])
        if "jacobi" in top:
            flops = nmol * (10.0 / 3.0) * n ** 3
            note = "algorithmic flop = (4/3+2) n^3 per matrix (dsyevd equivalent); Jacobi executes ~%.1f sweeps x 8 n^3 DMMA flop per single point" % sweeps
        elif "X^T H X" in top:
            flops = nmol * 2 * 2.0 * n ** 3 / 2   # two GEMM launches per timed pair; per launch 2 n^3
            note = "2 n^3 flop per launch and matrix"
        else:
            flops = nmol * 2.0 * n ** 3
            note = "2 n^3 flop per launch and matrix (upper bound; density uses occupied columns only)"
        peak = ctx.fp64_peak()
        ach = flops / (k_ms * 1e-3) / 1e12
        # DRAM traffic per launch from the committed ncu capture of this kernel at this matrix order (None: no capture)
        tr = ncu_traffic("k_jacobi_blk", n) if "jacobi" in top else None
        out = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
               "traffic": tr["bytes_per_molecule"] * nmol if tr else None,
               "peak_source": "cuBLAS DGEMM 4096^3 measured in this run (FP64 DMMA; MEASURED_PEAKS.json has no FP64 entry)",
               "note": note}
        if tr:
            out["traffic_source"] = tr.get("source")
            out["algorithmic_bytes"] = 3 * 8 * n * n * nmol   # T and Z in, Z out
        return out
    per_mol = {"k_sdqh0": 88 * n * n, "k_fock": 8 * (2.5 * n * n), "populations": 8 * 11 * n * n,
               "k_grad_h0": 8 * 4 * n * n}
    b = next((v for kk, v in per_mol.items() if kk in top), 8 * n * n)
    ach = nmol * b / (k_ms * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm if hbm else None, "traffic": None,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst)"}


def run_water(ctx, args, steps=None, warmup=None):
    """BASELINE configs[3]: one large single system, (H2O)_n cluster (default n = 1000, 3000 atoms, 6000 AOs);
    metric = seconds per SCC iteration as the reference's "iterations" timer slot defines it (src/scf_module.F90:606-663:
    solver init + Cholesky of S + all SCC iterations) divided by the iteration count.  With N > 1 ranks the integral /
    gradient blocks are split over the GPUs (xtb_b200/sharded.py, NCCL allreduce of the contributions); the SCC iterations
    - the eigensolve - run replicated on every GPU (they stay on one GPU by design), so the per-iteration metric does not
    scale and the line carries the phase split (setup+integrals / SCC / gradient / exchanges) that shows what does, and
    the result of the split's parity check against the same system on one GPU (energy bit-identical)."""
    from xtb_b200.sharded import ShardedCalculator
    import xtb_b200
    torch, dev, rank, local, world = ctx.torch, ctx.dev, ctx.rank, ctx.local, ctx.world
    steps = args.steps if steps is None else steps
    warmup = args.warmup if warmup is None else warmup
    Z, xyz = make_water_cluster(args.nwat)
    sc = ShardedCalculator(Z, rank, world, local)
    b = sc.calc
    sz = b.sizes(0)
    clocks = ClockSampler(local)
    for _ in range(max(warmup, 0)):
        b.singlepoint(xyz)
    ctx.barrier()
    clocks.start()
    b.set_profile("all")
    per_iter, per_iter_ref, tot_ms, phases, launches, call_ms = [], [], [], [], 0, []
    for _ in range(steps):
        r = b.singlepoint(xyz)
        t = b.timings()
        prof, passes = b.kernel_times()
        per_iter.append(t["scc_ms"] * 1e-3 / max(passes, 1))
        per_iter_ref.append((t["scc_ms"] + t["factorisation_ms"]) * 1e-3 / max(passes, 1))
        tot_ms.append(t["setup_ms"] + t["scc_ms"] + t["grad_ms"])
        phases.append([t["setup_ms"], t["scc_ms"], t["grad_ms"], b.comm_ms(), t["factorisation_ms"]])
        call_ms.append(b.call_ms())
        launches += t["launches"]
    ctx.barrier()
    clk = clocks.stop()
    e_split = float(r["energies"][0])
    g_split = r["gradients"].copy()
    # max over ranks of the device times
    agg = ctx.reduce([float(np.mean(per_iter_ref)), float(np.mean(tot_ms))] + list(np.mean(np.array(phases), axis=0))
                     + [float(np.mean(per_iter)), float(np.mean(call_ms))], "max")
    value, step_ms, ph_setup, ph_scc, ph_grad, ph_comm, ph_fact, value_pure, e2e_ms = agg
    # parity of the split inside the bench run (pytest on a 1-GPU box can only skip it): the same system on ONE GPU
    shard_check = None
    if world > 1:
        if rank == 0:
            b1 = xtb_b200.BatchCalculator([Z], device=local)
            r1 = b1.singlepoint(xyz)
            shard_check = {"ranks": world, "energy_identical": bool(float(r1["energies"][0]) == e_split),
                           "dE": float(r1["energies"][0]) - e_split,
                           "max_abs_dgradient": float(np.max(np.abs(r1["gradients"] - g_split))),
                           "niter_identical": bool(int(r1["niter"][0]) == int(r["niter"][0])),
                           "ndp_nqp_identical": bool(b1.counts(0)["ndp"] == b.counts(0)["ndp"] and b1.counts(0)["nqp"] == b.counts(0)["nqp"])}
            del b1
        ctx.barrier()
    if rank != 0:
        del sc, b
        return None
    n = sz["nao"]
    ndiag = steps * max(passes, 1)
    peak = ctx.fp64_peak()
    hbm = load_peaks().get("hbm_gbs")
    eig = {kk: v for kk, v in prof.items() if kk.startswith("eig:") and v["launches"]}
    stage_ms = {kk: v["ms"] / ndiag for kk, v in eig.items()}
    topk = max(prof, key=lambda kk: prof[kk]["ms"])
    roofline = eig_roofline(prof, topk, n, ndiag, steps, step_ms, stage_ms, peak, hbm)
    share = {kk: round(v["ms"] / steps / step_ms, 4) for kk, v in sorted(prof.items(), key=lambda x: -x[1]["ms"])
             if v["ms"] / steps / step_ms >= 0.005}
    cpu = None
    if not args.no_cpu and world == 1:
        from oracle.oracle import Oracle, build
        build()
        cores = host_cores()
        # one oracle run limited to two SCC iterations; its timer slot 5 is the reference's "iterations" slot
        # (src/scf_module.F90:606-663: dpotrf + all iterations), the same quantity `value` reports for the GPU
        o = Oracle(Z, xyz, maxiter=2, threads=cores)
        t0 = time.perf_counter(); o.singlepoint(); t_all = time.perf_counter() - t0
        eo = o.energies["e_total"]
        t_it = float(o.get("timers")[5])
        # the same two iterations on the GPU: the energy after iteration 2 is a deterministic function of the algorithm
        b.set_options(maxiter=2)
        r2 = b.singlepoint(xyz)
        b.set_options(maxiter=250)
        cpu = {"value": t_it / 2.0, "unit": "s/SCC-iteration", "cores": cores, "kind": "port",
               "sample": "the first two SCC iterations of the same system: the oracle's 'iterations' timer slot (dpotrf + iterations) / 2; "
                         "whole oracle call incl. integrals and gradient %.1f s" % t_all,
               "dE_vs_gpu_after_2_iterations": float(r2["energies"][0]) - eo,
               "note": "full converged parity (E, gradient, charges, niter, ndp/nqp) of this system: tests/test_gpu_large_parity.py"}
    line = {
        "metric": "s/SCC-iter at 3k atoms", "value": value, "unit": "s/SCC-iteration", "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": step_ms, "higher_is_better": False, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "(H2O)_%d cluster single point + gradient (%d atoms, %d AOs), seed 3000" % (args.nwat, sz["natoms"], n),
                   "scc_iterations": int(r["niter"][0]), "converged": bool(r["status"][0] == 0),
                   "parallelism": ("integral/gradient blocks split x%d + NCCL allreduce; SCC (eigensolve) replicated" % world)
                   if world > 1 else "one GPU",
                   "value_definition": "(overlap factorisation + all SCC iterations) / iterations = the reference's 'iterations' "
                                       "timer slot (src/scf_module.F90:606-663) per iteration; iterations alone: value_scc_only",
                   "l2": "matrices of %.2f GB each >> 126 MB L2" % (8.0 * n * n / 1e9)},
        "value_scc_only": value_pure,
        "phases_ms": {"setup+integrals+S factorisation": ph_setup, "of which S factorisation": ph_fact, "scc": ph_scc,
                      "gradient+properties": ph_grad, "of which exchanges (NCCL allreduce)": ph_comm},
        "comm_ms": ph_comm, "shard_check": shard_check,
        "e2e": {"value": value, "unit": "s/SCC-iteration", "h2d_bytes_per_step": int(xyz.size * 8),
                "d2h_bytes_per_step": int(8 * (1 + 4 * sz["natoms"]) + 8), "ms_per_call": e2e_ms,
                "note": "xtbb200_singlepoint with host buffers: ms_per_call is the whole call incl. H2D/D2H; the per-iteration "
                        "metric is taken from the same calls (all timed steps go through the host-buffer ABI)"},
        "gpu_launches": launches,
        "roofline": roofline, "share_all_kernels": share,
        "cpu_baseline": cpu, "clocks": clk, "sample_energy_Eh": e_split,
    }
    del sc, b
    torch.cuda.empty_cache()
    ms = library_eigh_calibration(torch, dev, n, 1)
    line["calibration"] = {"what": "cuSOLVER symmetric eigensolver (torch.linalg.eigh) on one random symmetric FP64 matrix of "
                                   "%d x %d; calibration only, not on the product path" % (n, n),
                           "library_ms_per_matrix": ms,
                           "engine_ms_per_diagonalisation": sum(roofline.get("eigensolver_ms_per_diagonalisation", {}).values()) or None}
    return line


def eig_roofline(prof, topk, n, ndiag, steps, step_ms, stage_ms, peak, hbm):
    """Roofline of the dominant kernel class of a large single system."""
    # trailing update: the lower triangle of the symmetric rank-2k update (what the symmetric panel kernel needs; the
    # two-triangle form used for odd orders executes twice that)
    gemm_flops = {"eig:trailing update(k_eg_gemm)": 2.0 * n ** 3 / 3.0, "eig:back-transformation": 2.0 * n ** 3,
                  "eig:band reduction(DMMA)": 4.0 * n ** 3 / 3.0}
    common = {"eigensolver_ms_per_diagonalisation": stage_ms,
              "dmma_tflops": {kk: gemm_flops[kk] / (stage_ms[kk] * 1e-3) / 1e12 for kk in gemm_flops if stage_ms.get(kk)},
              "fp64_tensor_peak_tflops(cuBLAS DGEMM 4096^3, this run)": peak}
    if stage_ms:
        common["dsyevd_equivalent_tflops"] = (10.0 / 3.0) * n ** 3 / (sum(stage_ms.values()) * 1e-3) / 1e12
    if topk == "eig:k_sytrd_panel":
        # Householder tridiagonalisation: the trailing-matrix product A v is HBM-bound.  Algorithmic bytes (DESIGN.md
        # section 4): the lower triangle of the trailing matrix once per column, 8 (n-j-1)^2 / 2, summed over the columns
        # = 4 (n-1) n (2n-1) / 6 per diagonalisation.
        pk = prof[topk]
        alg = 4.0 * (n - 1) * n * (2 * n - 1) / 6.0
        sec_panel = pk["ms"] * 1e-3 / ndiag
        npan = (n + 63) // 64
        r = {"bound": "hbm", "kernel": "k_sytrd_panel (Householder tridiagonalisation: column update, reflector, trailing-matrix product)",
             "achieved": alg / sec_panel / 1e9, "peak": hbm, "unit": "GB/s", "frac": (alg / sec_panel / 1e9 / hbm) if hbm else None,
             "traffic": None, "algorithmic_bytes_per_launch": alg / npan, "avg_launch_ms": 1e3 * sec_panel / npan,
             "launches_per_diagonalisation": npan, "peak_source": "MEASURED_PEAKS.json hbm_gbs (kernel timed inside a long step)",
             "note": "achieved = sum of algorithmic bytes of the %d panel launches / sum of their durations" % npan,
             "share_of_step": pk["ms"] / steps / step_ms}
        tr = ncu_traffic("k_sytrd_panel", n)
        if tr:
            r["traffic"] = tr["ratio_to_algorithmic"] * alg / npan
            r["traffic_source"] = tr.get("source")
        r.update(common)
        return r
    k = prof[topk]
    flops = gemm_flops.get(topk, (10.0 / 3.0) * n ** 3)
    sec = k["ms"] * 1e-3 / ndiag
    r = {"bound": "tensor", "kernel": topk, "achieved": flops / sec / 1e12, "peak": peak, "unit": "TFLOP/s",
         "frac": flops / sec / 1e12 / peak, "traffic": None, "avg_ms_per_diagonalisation": 1e3 * sec,
         "note": "algorithmic flop of this stage per diagonalisation / its device time",
         "peak_source": "cuBLAS DGEMM 4096^3 measured in this run", "share_of_step": k["ms"] / steps / step_ms}
    r.update(common)
    return r


def single_system(name):
    if name == "protein" and not os.path.exists(os.path.join(ROOT, "xtb_b200", "polypeptide.py")):
        name = "taxol6"
    if name == "protein":
        from xtb_b200.polypeptide import make_polypeptide
        Z, xyz, info = make_polypeptide()
        return Z, xyz, ("synthetic polypeptide %s, %d atoms, seed %d: covalently connected random coil with crambin's sequence, "
                        "free thiols (crambin C202H315N55O64S6 + 6 H; SURVEY 8(d) C2)" % (info["formula"], len(Z), info["seed"]))
    if name == "taxol6":
        Z, xyz = make_peptide_cluster(load_kat("taxol"), 6, gap=20.0)
        return Z, xyz, "6 x taxol, 20 bohr apart"
    s = load_kat(name)
    return s["numbers"], s["xyz"], name


def run_single(ctx, args, molecule=None, steps=None, warmup=None):
    """BASELINE configs[0] / configs[1]: wall time of ONE GFN2-xTB single point + gradient through the xtb.h C ABI
    (xtb_newMolecule / xtb_loadGFN2xTB / xtb_singlepoint / xtb_getEnergy / xtb_getGradient), fresh results handle (EEQ
    guess) per step.  --molecule caffeine: the 24-atom molecule of `xtb caffeine.xyz --gfn 2 --grad`; --molecule
    protein: a seeded, covalently connected 648-atom polypeptide with crambin's sequence, C202H321N55O64S6 = crambin's
    composition with free thiols (xtb_b200/polypeptide.py; crambin's coordinates are not available offline)."""
    import xtb_b200
    torch = ctx.torch
    if ctx.rank != 0:
        return None                                       # does not shard: one molecule, one GPU
    molecule = molecule or args.molecule
    steps = args.steps if steps is None else steps
    warmup = args.warmup if warmup is None else warmup
    Z, xyz, label = single_system(molecule)
    calc = xtb_b200.Calculator(xtb_b200.Param.GFN2xTB, Z, xyz)
    calc.set_verbosity(0)
    for _ in range(max(warmup, 1)):
        res = calc.singlepoint()
    clocks = ClockSampler(ctx.local)
    clocks.start()
    torch.cuda.synchronize()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        res = calc.singlepoint()                          # new results handle: no restart
        e = res.get_energy()
        g = res.get_gradient()
        ts.append(time.perf_counter() - t0)
    clk = clocks.stop()
    # device-side split of the same calculation through the batched ABI (positions resident on the device)
    b = xtb_b200.BatchCalculator([Z], device=ctx.local)
    pos_dev = torch.from_numpy(np.ascontiguousarray(xyz)).to(ctx.dev)
    b.set_profile("all")
    dev_ms = []
    for _ in range(max(warmup, 1) + steps):
        b.singlepoint_device(pos_dev.data_ptr())
        t = b.timings()
        dev_ms.append(t["setup_ms"] + t["scc_ms"] + t["grad_ms"])
    dev_ms = dev_ms[-steps:]
    prof, passes = b.kernel_times()
    tot = sum(v["ms"] for v in prof.values())
    share = {kk: round(v["ms"] / tot, 4) for kk, v in sorted(prof.items(), key=lambda x: -x[1]["ms"]) if v["ms"] / tot >= 0.01}
    r = b.result(0)
    sz = b.sizes(0)
    n = sz["nao"]
    topk = max(prof, key=lambda k: prof[k]["ms"])
    if n > 152:     # Householder + divide and conquer path: same roofline bookkeeping as the water workload
        ncall = max(warmup, 1) + steps
        ndiag = ncall * max(passes, 1)
        stage_ms = {kk: v["ms"] / ndiag for kk, v in prof.items() if kk.startswith("eig:") and v["launches"]}
        roof = eig_roofline(prof, topk, n, ndiag, ncall, float(np.mean(dev_ms)), stage_ms, ctx.fp64_peak(), load_peaks().get("hbm_gbs"))
    else:
        roof = {"bound": "latency", "kernel": topk, "achieved": None, "peak": None, "unit": None, "frac": None, "traffic": None,
                "note": "one small molecule cannot fill the GPU; see the ensemble workload for throughput"}
    cpu = None
    if not args.no_cpu:
        from oracle.oracle import Oracle, build
        build()
        cores = host_cores()
        best = None
        for nth in sorted({1, cores}) if len(Z) < 200 else [cores]:   # tiny molecules run fastest on one thread
            for _ in range(2 if len(Z) < 200 else 1):
                o = Oracle(Z, xyz, threads=nth)
                t0 = time.perf_counter(); st = o.singlepoint(); dt = time.perf_counter() - t0
                if best is None or dt < best[0]:
                    best = (dt, nth)
        cpu = {"value": 1e3 * best[0], "unit": "ms/single point", "cores": best[1], "kind": "port",
               "sample": "the same molecule, OpenMP + OpenBLAS, best of the runs with 1 / %d threads" % cores,
               "dE_vs_gpu": float(e - o.energies["e_total"]), "niter": int(o.info["niter"]),
               "niter_identical": bool(int(o.info["niter"]) == int(r["niter"])),
               "max_abs_dgradient_vs_gpu": float(np.max(np.abs(g - o.get("gradient").reshape(-1, 3))))}
    line = {
        "metric": "wall time per GFN2 SP+gradient (single molecule)", "value": float(np.mean(dev_ms)), "unit": "ms/single point",
        "n_gpus": 1, "steps": steps, "warmup": warmup, "ms_per_step": float(np.mean(dev_ms)), "higher_is_better": False,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s: %d atoms, %d AOs, xtb.h C API, EEQ guess every call" % (label, sz["natoms"], sz["nao"]),
                   "scc_iterations": int(r["niter"]), "l2": "single system: working set %.0f MB" % (b.device_bytes() / 1e6)},
        "e2e": {"value": 1e3 * float(np.mean(ts)), "unit": "ms/single point", "h2d_bytes_per_step": int(np.asarray(xyz).size * 8),
                "d2h_bytes_per_step": int(8 * (1 + 3 * sz["natoms"])),
                "note": "xtb_singlepoint + xtb_getEnergy + xtb_getGradient, host buffers, wall clock around the blocking calls"},
        "gpu_launches": int(b.timings()["launches"]) * steps,
        "roofline": roof,
        "share_all_kernels": share, "cpu_baseline": cpu, "clocks": clk, "sample_energy_Eh": float(e),
    }
    del b, calc
    return line


def run_md(ctx, args, steps=None, warmup=None):
    """BASELINE configs[4]: GFN2-xTB molecular dynamics of a ~200-atom peptide in ALPB(water), MD steps per second.
    System: 3 copies of the 76-atom peptide pdb-4qxx (test/unit/molstock.f90:739-909) 8 bohr apart = 228 atoms (SURVEY 8(d)
    C5); xtb's default $md settings (4 fs, Berendsen 298.15 K, SHAKE on all bonds, hydrogen mass 4 amu).  A step = one
    restarted single point + gradient and one propagation kernel.  MD does not shard (replicas only): with
    --gpus N every rank runs its own replica (other start velocities) and the value is the sum over ranks.
    value: xtbb200_mdRun, positions / velocities / forces resident on the device; e2e: the loop a user of the reference's
    API drives from the host (positions H2D, energy + gradient D2H every step, leap-frog + thermostat on the host)."""
    import ctypes

    import xtb_b200
    torch, dev, rank, local, world = ctx.torch, ctx.dev, ctx.rank, ctx.local, ctx.world
    steps_arg = args.steps if steps is None else steps
    warmup = args.warmup if warmup is None else warmup
    Z, xyz = make_peptide_cluster(load_kat("pdb_4qxx"), 3, gap=8.0)
    nat = len(Z)
    solvent = dict(solvent="water", model="alpb")
    b = xtb_b200.BatchCalculator([Z], device=local)
    b.set_solvent(**solvent)
    ncons = b.md_init(xyz, seed=1 + rank)
    sz = b.sizes(0)
    steps = max(steps_arg, 1) * 20                        # a "step" of the bench contract = 20 MD steps
    b.set_profile("all")
    b.md_run(max(warmup, 3) * 5)
    prof, passes = b.kernel_times()
    tot = sum(v["ms"] for v in prof.values())
    share = {kk: round(v["ms"] / tot, 4) for kk, v in sorted(prof.items(), key=lambda x: -x[1]["ms"]) if v["ms"] / tot >= 0.01}
    top = max(prof, key=lambda k: prof[k]["ms"])
    b.set_profile([])
    clocks = ClockSampler(local)
    ctx.barrier()
    clocks.start()
    t0 = time.perf_counter()
    r = b.md_run(steps)
    torch.cuda.synchronize(dev)
    wall = time.perf_counter() - t0                       # md_run blocks on the engine's stream
    clk = clocks.stop()
    launches = int(b.timings()["launches"]) * steps
    ctx.barrier()
    # ---- end to end: the same dynamics (same thermostat, SHAKE, time step) driven from the host one step per call, with what
    # a user who keeps a trajectory reads back every step: energies (3 doubles) from xtbb200_mdRun and the frame
    # (positions + velocities, device -> host) through xtbb200_getArray
    nh = max(steps_arg, 1) * 10
    t0 = time.perf_counter()
    frames = []
    for _ in range(nh):
        r1 = b.md_run(1)
        frames.append((b.array("xyz"), b.array("velocities"), r1["epot"][0, 0]))
    wall_host = time.perf_counter() - t0
    rate, rate_host = steps / wall, nh / wall_host
    if world > 1:
        rate, rate_host = ctx.reduce([rate, rate_host], "sum")
        wall = ctx.reduce([wall], "max")[0]
    if rank != 0:
        del b
        return None
    cpu = None
    if not args.no_cpu and world == 1:
        from oracle.md_oracle import MDOracle
        from oracle.oracle import build
        build()
        cores = host_cores()
        sg = np.where(np.random.default_rng(1).random((nat, 3)) > 0.5, -1.0, 1.0)
        mo = MDOracle(Z, xyz, signs=sg, solvent=solvent, threads=cores)
        mo.run(1)
        t0 = time.perf_counter(); mo.run(4); dtc = time.perf_counter() - t0
        cpu = {"value": 4 / dtc, "unit": "MD steps/s", "cores": cores, "kind": "port",
               "sample": "4 MD steps of the same system and settings with the CPU oracle (OpenMP + OpenBLAS)"}
    line = {
        "metric": "GFN2-xTB MD steps/sec (228-atom peptide cluster, ALPB water)", "value": rate, "unit": "MD steps/s",
        "n_gpus": world, "steps": steps_arg, "warmup": warmup, "ms_per_step": 1e3 * wall / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "MD, 3 x pdb-4qxx (%d atoms, %d AOs), ALPB(water), 4 fs, NVT 298.15 K, SHAKE on %d bonds, H mass 4; "
                               "one replica per GPU (replicas only, SURVEY 8e); %d MD steps timed" % (nat, sz["nao"], ncons, steps),
                   "scc_iterations_last_step": int(b.result(0)["niter"]),
                   "l2": "one small system: latency bound, working set in L2"},
        "e2e": {"value": rate_host, "unit": "MD steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": int(8 * (3 + 6 * nat)),
                "note": "host-driven: one xtbb200_mdRun step per call, energies and the frame (positions, velocities) read back "
                        "every step, %d steps; same thermostat / SHAKE / time step as `value`" % nh},
        "gpu_launches": launches,
        "roofline": {"bound": "latency", "kernel": top, "achieved": None, "peak": None, "unit": None, "frac": None, "traffic": None,
                     "note": "a 228-atom system cannot fill the GPU; kernel-class shares below"},
        "share_all_kernels": share, "cpu_baseline": cpu, "clocks": clk,
        "trajectory": {"epot_first": float(r["epot"][0, 0]), "epot_last": float(r["epot"][-1, 0]),
                       "T_mean": float(r["temperature"][:, 0].mean()), "steps_done": int(r["steps"])},
    }
    del b
    return line


def run_secondary(ctx, args):
    """The other BASELINE configs, bounded: every entry is a complete line of its own workload (or {"error": ...})."""
    out = {}
    plan = [("water", lambda: run_water(ctx, args, steps=2, warmup=1)),
            ("caffeine", lambda: run_single(ctx, args, molecule="caffeine", steps=20, warmup=3)),
            ("protein", lambda: run_single(ctx, args, molecule="protein", steps=3, warmup=1)),
            ("md", lambda: run_md(ctx, args, steps=3, warmup=3))]
    for name, fn in plan:
        t0 = time.perf_counter()
        try:
            line = fn()
        except Exception as ex:   # a secondary workload must never take the headline line down with it
            line = {"error": "%s: %s" % (type(ex).__name__, ex)}
            if ctx.world > 1:
                raise             # ... but ranks must not diverge inside collectives
        ctx.barrier()
        if ctx.rank == 0:
            if line is not None:
                line["bench_wall_s"] = time.perf_counter() - t0
            out[name] = line
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--nmol", type=int, default=4096, help="conformers per GPU")
    ap.add_argument("--molecule", default="rivaroxaban")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--workload", default="ensemble", choices=["ensemble", "water", "single", "md"],
                    help="ensemble: BASELINE configs[2] (default, the headline metric, with the other configs nested under "
                         "`secondary`); water: configs[3], s/SCC-iteration; single: configs[0]/[1], one molecule (--molecule "
                         "caffeine | protein) through the xtb.h C API; md: configs[4], MD steps/s of a 228-atom peptide cluster "
                         "in ALPB water")
    ap.add_argument("--secondary", default="auto", choices=["auto", "on", "off"],
                    help="nest the water / caffeine / protein / md lines into the ensemble line (auto: for the default workload)")
    ap.add_argument("--solvent", default=None, help="ensemble workload: implicit solvent (e.g. water) instead of the gas phase")
    ap.add_argument("--solvent-model", default="alpb", choices=["gbsa", "alpb"])
    ap.add_argument("--nwat", type=int, default=1000)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    ctx = Ctx()
    if args.workload == "water":
        line = run_water(ctx, args)
    elif args.workload == "md":
        line = run_md(ctx, args)
    elif args.workload == "single":
        if args.molecule == "rivaroxaban":
            args.molecule = "caffeine"
        line = run_single(ctx, args)
    else:
        line = run_engine(ctx, args)
        default_wl = args.molecule == "rivaroxaban" and not args.solvent
        if args.secondary == "on" or (args.secondary == "auto" and default_wl):
            ctx.torch.cuda.empty_cache()
            sec = run_secondary(ctx, args)
            if line is not None:
                line["secondary"] = sec
    if ctx.rank == 0 and line is not None:
        print(json.dumps(line), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
