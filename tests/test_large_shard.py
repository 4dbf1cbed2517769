"""Split of one large system over GPUs (second row): ownership rules, the exchange of the NCCL id over
a process group (world_size 2, gloo, CPU) and, on a box with >= 2 GPUs, the sharded engine against the single-GPU
engine."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import xtb_b200
from xtb_b200 import sharded

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ownership_rules_tile_the_work_and_match_the_library():
    lib = xtb_b200.load_library()
    for nranks in (1, 2, 3, 4, 8):
        for natoms in (1, 31, 32, 33, 192, 3000):
            owners = np.array([sharded.atom_owner(i, nranks) for i in range(natoms)])
            assert owners.min() >= 0 and owners.max() < nranks
            assert [lib.xtbb200_shardOwner(0, i, nranks) for i in range(natoms)] == owners.tolist()
            cover = np.concatenate([sharded.owned_atoms(natoms, r, nranks) for r in range(nranks)])
            assert sorted(cover.tolist()) == list(range(natoms))          # every atom exactly once
            if natoms >= 32 * nranks * 8:                                  # large systems are balanced to one chunk
                cnt = np.bincount(owners, minlength=nranks)
                assert cnt.max() - cnt.min() <= sharded.ATOM_CHUNK
        ctas = np.array([sharded.pair_cta_owner(c, nranks) for c in range(1000)])
        assert [lib.xtbb200_shardOwner(1, c, nranks) for c in range(1000)] == ctas.tolist()
        cnt = np.bincount(ctas, minlength=nranks)
        assert cnt.max() - cnt.min() <= 1
    assert lib.xtbb200_shardOwner(0, -1, 2) == -1 and lib.xtbb200_shardOwner(7, 0, 2) == -1


def _id_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    calls = []

    def make_id():
        calls.append(rank)
        return bytes((7 * i + 3) % 256 for i in range(sharded.UNIQUE_ID_BYTES))

    uid = sharded.exchange_unique_id(make_id, rank, world)
    q.put((rank, uid, calls))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_unique_id_exchange_two_ranks_gloo():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_id_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=240) for _ in range(2)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    expect = bytes((7 * i + 3) % 256 for i in range(sharded.UNIQUE_ID_BYTES))
    for rank, uid, calls in got:
        assert uid == expect
        assert calls == ([0] if rank == 0 else [])       # only rank 0 creates the id


@pytest.mark.gpu
def test_set_shard_validates_arguments():
    b = xtb_b200.BatchCalculator([np.array([8, 1, 1], dtype=np.int32)])
    with pytest.raises(xtb_b200.XTBException):
        b.set_shard(2, 2, bytes(128))
    with pytest.raises(xtb_b200.XTBException):
        b.set_shard(0, 2, None)
    b.set_shard(0, 1)       # switching the split off is always allowed


@pytest.mark.gpu
@pytest.mark.timeout(900)
def test_two_gpu_split_matches_single_gpu():
    if xtb_b200.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tools", "shard_check.py"), "water", "64"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=800)
    assert "SHARD OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
    # the same with the column-parallel products of the SCC iterations split over the two ranks as well (by default only
    # from 2500 AOs on): X^T H X, back-transformation, C = X Z and density by tile columns, gathered by NCCL broadcasts
    env = dict(os.environ, XTBB200_TUNING="1", XTBB200_GEMM_SPLIT_MIN="0")
    cmd[cmd.index(str(port))] = str(port + 1 if port < 65000 else port - 1)
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=800, env=env)
    assert "SHARD OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
