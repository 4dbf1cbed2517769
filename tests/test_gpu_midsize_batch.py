"""Batches of mid-size matrices of one order (conformers of a 60-150 atom molecule, 153-472 AOs): the batched Householder +
divide and conquer path with one thread-block cluster per matrix in the tridiagonalisation (k_sytrd_smem<2, true>) against
the one-matrix-at-a-time GPU-wide solver (set_solver(5)) and against the CPU oracle."""
import os
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def xb():
    import xtb_b200
    if xtb_b200.device_count() < 1:
        pytest.fail("no CUDA device: the engine has no CPU fallback")
    return xtb_b200


@pytest.mark.parametrize("name,nconf", [("taxol", 6), ("remdesivir", 5)])
def test_cluster_batch_matches_sequential_solver_and_oracle(xb, kat, name, nconf):
    from oracle.oracle import Oracle
    s = kat[name]
    rng = np.random.default_rng(3)
    geos = [s["xyz"] + (rng.normal(0.0, 0.03, s["xyz"].shape) if k else 0.0) for k in range(nconf)]
    pos = np.concatenate(geos)
    res = {}
    for mode in ("cluster", "sequential"):
        b = xb.BatchCalculator([s["numbers"]] * nconf)
        if mode == "sequential":
            b.set_solver(5)                         # Householder + D&C, one matrix at a time through the GPU-wide kernels
        assert 152 < b.sizes(0)["nao"] <= 472
        b.singlepoint(pos)                          # warm-up (allocations)
        t0 = time.perf_counter()
        r = b.singlepoint(pos)
        res[mode] = (r, time.perf_counter() - t0, [int(b.result(i)["niter"]) for i in range(nconf)])
    rc, rs = res["cluster"][0], res["sequential"][0]
    assert rc["all_converged"] and rs["all_converged"]
    assert np.max(np.abs(rc["energies"] - rs["energies"])) < 1e-8
    assert np.max(np.abs(rc["gradients"] - rs["gradients"])) < 1e-7
    assert res["cluster"][2] == res["sequential"][2]
    print("\n%s x %d: batched cluster path %.1f ms, sequential GPU-wide solver %.1f ms" % (
        name, nconf, 1e3 * res["cluster"][1], 1e3 * res["sequential"][1]))
    o = Oracle(s["numbers"], geos[1], threads=8)
    assert o.singlepoint() == 0
    n = len(s["numbers"])
    assert abs(rc["energies"][1] - o.energies["e_total"]) < 1e-7
    assert np.max(np.abs(rc["gradients"][n:2 * n] - o.get("gradient").reshape(-1, 3))) < 1e-6
    assert res["cluster"][2][1] == o.info["niter"]


def test_cluster16_batch(xb, kat):
    """Orders between 473 and ~670 AOs with clusters of 16 CTAs (non-portable cluster size): four replicas of the 228-atom peptide cluster (552 AOs) in ALPB water against the one-matrix-at-a-time
    solver."""
    import bench
    Z, xyz = bench.make_peptide_cluster(bench.load_kat("pdb_4qxx"), 3, gap=8.0)
    rng = np.random.default_rng(9)
    nrep = 4
    pos = np.concatenate([xyz + (rng.normal(0.0, 0.02, xyz.shape) if k else 0.0) for k in range(nrep)])
    res = {}
    for mode in ("cluster16", "sequential"):
        b = xb.BatchCalculator([Z] * nrep)
        if mode == "sequential":
            b.set_solver(5)
        b.set_solvent("water", model="alpb")
        b.singlepoint(pos)
        t0 = time.perf_counter()
        r = b.singlepoint(pos)
        res[mode] = (r, time.perf_counter() - t0, [int(b.result(i)["niter"]) for i in range(nrep)])
    rc, rs = res["cluster16"][0], res["sequential"][0]
    assert rc["all_converged"] and rs["all_converged"]
    print("\npeptide cluster x %d (552 AOs): cluster16 path %.1f ms, sequential %.1f ms" % (
        nrep, 1e3 * res["cluster16"][1], 1e3 * res["sequential"][1]))
    assert np.max(np.abs(rc["energies"] - rs["energies"])) < 1e-7
    assert np.max(np.abs(rc["gradients"] - rs["gradients"])) < 1e-6
    assert res["cluster16"][2] == res["sequential"][2]
