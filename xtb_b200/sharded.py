"""One large system split over the GPUs of a box (second row).

Every rank (one process per GPU, torch.distributed) builds the SAME molecule on its own device.  The engine deals the
shell-pair blocks of the integral build and of its derivative, and the per-atom gradient terms, to the ranks and sums
the contributions with NCCL allreduce on its own stream (xtbb200_setShard, include/xtb_b200.h); the SCC iterations --
the eigensolver -- run replicated, so every rank returns the complete result and no broadcast is needed.

Host-side logic in this file: the ownership rules (mirrors of the device functions, used by the tests), and the
exchange of rank 0's NCCL unique id over the process group (works with the gloo backend as well, which is how the
CPU test exercises it).
"""
import ctypes

import numpy as np

from .interface import BatchCalculator, load_library

ATOM_CHUNK = 32     # atoms are owned in warp-sized chunks (common.cuh: shard_atom_owner)
UNIQUE_ID_BYTES = 128


def atom_owner(iatom, nranks):
    """Rank that evaluates the per-atom gradient terms of atom `iatom` (k_grad_aes, k_grad_pairs, k_grad_atm,
    k_grad_cn)."""
    return 0 if nranks <= 1 else (int(iatom) // ATOM_CHUNK) % int(nranks)


def pair_cta_owner(cta, nranks):
    """Rank that evaluates shell-pair CTA `cta` of k_sdqh0 / k_grad_h0 (CTAs are interleaved over the ranks)."""
    return 0 if nranks <= 1 else int(cta) % int(nranks)


def owned_atoms(natoms, rank, nranks):
    return np.array([i for i in range(natoms) if atom_owner(i, nranks) == rank], dtype=np.int64)


def exchange_unique_id(make_id, rank, world, group=None):
    """Rank 0 calls make_id() -> 128 bytes; every rank returns the same bytes (broadcast over the process group)."""
    import torch
    import torch.distributed as dist

    if world == 1:
        return bytes(make_id())
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    buf = torch.zeros(UNIQUE_ID_BYTES, dtype=torch.uint8)
    if rank == 0:
        uid = bytes(make_id())
        if len(uid) != UNIQUE_ID_BYTES:
            raise ValueError("NCCL unique id must have %d bytes" % UNIQUE_ID_BYTES)
        buf = torch.tensor(list(uid), dtype=torch.uint8)
    buf = buf.to(dev)
    dist.broadcast(buf, src=0, group=group)
    return bytes(buf.cpu().numpy().tobytes())


def new_unique_id():
    buf = ctypes.create_string_buffer(UNIQUE_ID_BYTES)
    lib = load_library()
    if lib.xtbb200_shardUniqueId(buf, UNIQUE_ID_BYTES) != 0:
        raise RuntimeError(lib.xtbb200_lastError().decode())
    return buf.raw


class ShardedCalculator:
    """GFN2-xTB SP+gradient of one large molecule on `world` GPUs.  Collective: construct and call on every rank."""

    def __init__(self, numbers, rank, world, device, charge=0.0, uhf=0, group=None, **options):
        self.rank, self.world = int(rank), int(world)
        self.calc = BatchCalculator([np.asarray(numbers, dtype=np.int32)], charges=[charge], uhf=[uhf], device=device,
                                    **options)
        if self.world > 1:
            uid = exchange_unique_id(new_unique_id, self.rank, self.world, group)
            self.calc.set_shard(self.rank, self.world, uid)

    def singlepoint(self, positions, restart=False):
        return self.calc.singlepoint(positions, restart=restart)
