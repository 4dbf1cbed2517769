"""Molecule sharding + result gather of the ensemble driver, world_size 2 on CPU with gloo.
The compute function is the CPU oracle here (test infrastructure); on a GPU box it is the CUDA engine."""
import os
import socket

import numpy as np
import pytest

from xtb_b200.ensemble import EnsembleRunner, shard_bounds


def test_shard_bounds_cover_and_balance():
    for nmol in (0, 1, 7, 4096, 4099):
        for world in (1, 2, 3, 8):
            b = shard_bounds(nmol, world)
            assert b[0][0] == 0 and b[-1][1] == nmol
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [y - x for x, y in b]
            assert max(sizes) - min(sizes) <= 1


def _oracle_compute(numbers_list, positions):
    from oracle.oracle import Oracle
    e, g, q, it, st = [], [], [], [], []
    off = 0
    for z in numbers_list:
        n = len(z)
        o = Oracle(z, positions[off:off + n], threads=1)
        st.append(o.singlepoint())
        e.append(o.energies["e_total"]); g.append(o.get("gradient").reshape(-1, 3)); q.append(o.get("q"))
        it.append(o.info["niter"])
        off += n
    return dict(energies=np.array(e), gradients=np.concatenate(g), charges=np.concatenate(q),
                niter=np.array(it, dtype=np.int32), status=np.array(st, dtype=np.int32))


def _ensemble():
    rng = np.random.default_rng(7)
    h2o = np.array([[0.0, 0.0, -0.74], [0.0, 1.43, 0.37], [0.0, -1.43, 0.37]])
    nh3 = np.array([[0.0, 0.0, 0.2], [0.0, 1.77, -0.5], [1.53, -0.89, -0.5], [-1.53, -0.89, -0.5]])
    numbers, pos = [], []
    for i in range(5):   # ragged: alternating 3- and 4-atom molecules, odd count over 2 ranks
        z, x = ([8, 1, 1], h2o) if i % 2 == 0 else ([7, 1, 1, 1], nh3)
        numbers.append(np.array(z, dtype=np.int32))
        pos.append(x + 0.05 * rng.standard_normal(x.shape))
    return numbers, np.concatenate(pos)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    numbers, pos = _ensemble()
    runner = EnsembleRunner(numbers, rank=rank, world=world, compute=_oracle_compute)
    res = runner.singlepoint(pos)
    if rank == 0:
        q.put({k: np.asarray(v) for k, v in res.items()})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_matches_serial():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    numbers, pos = _ensemble()
    ref = EnsembleRunner(numbers, compute=_oracle_compute).singlepoint(pos)
    assert got["energies"].shape == (5,) and got["gradients"].shape == (17, 3)
    assert np.array_equal(got["energies"], ref["energies"])
    assert np.array_equal(got["gradients"], ref["gradients"])
    assert np.array_equal(got["charges"], ref["charges"])
    assert np.array_equal(got["niter"], ref["niter"])
    assert bool(got["all_converged"])
