"""Conformer-ensemble driver: shards an ensemble by molecule over the ranks of a torch.distributed job.

SURVEY.md 8(e): ensembles shard naturally (independent molecules), so every rank owns a contiguous block of molecules,
runs them through its own BatchCalculator on its own GPU, and the only communication is the final result gather
(energies, gradients, charges, iteration counts).  No collective sits on the compute path.

The compute function is injectable so that the sharding / gather logic can be exercised on CPU with the gloo
backend (tests/test_ensemble_sharding.py); in production it is BatchCalculator.singlepoint.
"""
import numpy as np


def shard_bounds(nmol, world):
    """Contiguous, balanced partition of nmol molecules over world ranks: list of (first, last+1)."""
    base, rem = divmod(int(nmol), int(world))
    bounds, first = [], 0
    for r in range(world):
        cnt = base + (1 if r < rem else 0)
        bounds.append((first, first + cnt))
        first += cnt
    return bounds


def shard_range(nmol, rank, world):
    return shard_bounds(nmol, world)[rank]


class EnsembleRunner:
    """GFN2-xTB SP+gradient for an ensemble, one rank per GPU.

    numbers_list : atomic numbers of every molecule of the WHOLE ensemble (same on all ranks)
    compute      : None -> BatchCalculator on `device` (CUDA engine); otherwise a callable
                   compute(numbers_list_local, positions_local) -> dict(energies, gradients, charges, niter, status)
    solvent      : None (gas phase) or the keyword arguments of BatchCalculator.set_solvent, e.g.
                   dict(solvent="water", model="alpb"): every rank sets the same implicit solvation model
    """

    def __init__(self, numbers_list, rank=0, world=1, device=0, compute=None, solvent=None, **options):
        self.rank, self.world = int(rank), int(world)
        self.numbers_list = [np.asarray(z, dtype=np.int32) for z in numbers_list]
        self.nmol = len(self.numbers_list)
        self.natoms = np.array([len(z) for z in self.numbers_list], dtype=np.int64)
        self.atom_offsets = np.concatenate([[0], np.cumsum(self.natoms)])
        self.bounds = shard_bounds(self.nmol, self.world)
        self.first, self.last = self.bounds[self.rank]
        self._compute = compute
        self._calc = None
        if compute is None and self.last > self.first:
            from .interface import BatchCalculator
            self._calc = BatchCalculator(self.numbers_list[self.first:self.last], device=device, **options)
            if solvent:
                self._calc.set_solvent(**solvent)

    def local_positions(self, positions):
        """Slice of the concatenated (total_atoms, 3) positions that belongs to this rank."""
        pos = np.asarray(positions, dtype=np.float64).reshape(-1, 3)
        return pos[self.atom_offsets[self.first]:self.atom_offsets[self.last]]

    def singlepoint_local(self, positions, restart=False):
        """Run this rank's shard; positions = concatenated positions of the whole ensemble."""
        loc = self.local_positions(positions)
        if self.last == self.first:
            return dict(energies=np.zeros(0), gradients=np.zeros((0, 3)), charges=np.zeros(0),
                        niter=np.zeros(0, dtype=np.int32), status=np.zeros(0, dtype=np.int32))
        if self._compute is not None:
            return self._compute(self.numbers_list[self.first:self.last], loc)
        return self._calc.singlepoint(loc, restart=restart)

    def gather(self, local, dst=0):
        """Result gather: returns the whole-ensemble result dict on rank dst (None elsewhere)."""
        if self.world == 1:
            return dict(local)
        import torch
        import torch.distributed as dist
        backend = dist.get_backend()
        dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
        max_mol = max(b[1] - b[0] for b in self.bounds)
        max_at = max(int(self.atom_offsets[b[1]] - self.atom_offsets[b[0]]) for b in self.bounds)
        nloc = self.last - self.first
        nat_loc = int(self.atom_offsets[self.last] - self.atom_offsets[self.first])
        # one padded float64 buffer per rank: [energies | gradients | charges | niter | status]
        width = max_mol + 3 * max_at + max_at + 2 * max_mol
        buf = torch.zeros(width, dtype=torch.float64)
        o = 0
        buf[o:o + nloc] = torch.from_numpy(np.asarray(local["energies"], dtype=np.float64)); o = max_mol
        buf[o:o + 3 * nat_loc] = torch.from_numpy(np.asarray(local["gradients"], dtype=np.float64).reshape(-1)); o += 3 * max_at
        buf[o:o + nat_loc] = torch.from_numpy(np.asarray(local["charges"], dtype=np.float64)); o += max_at
        buf[o:o + nloc] = torch.from_numpy(np.asarray(local["niter"], dtype=np.float64)); o += max_mol
        buf[o:o + nloc] = torch.from_numpy(np.asarray(local["status"], dtype=np.float64))
        buf = buf.to(dev)
        out = [torch.zeros_like(buf) for _ in range(self.world)]
        dist.all_gather(out, buf)
        if self.rank != dst:
            return None
        e, g, q, it, st = [], [], [], [], []
        for r, (a, b) in enumerate(self.bounds):
            v = out[r].cpu().numpy()
            nm, na = b - a, int(self.atom_offsets[b] - self.atom_offsets[a])
            o = 0
            e.append(v[o:o + nm]); o = max_mol
            g.append(v[o:o + 3 * na].reshape(-1, 3)); o += 3 * max_at
            q.append(v[o:o + na]); o += max_at
            it.append(v[o:o + nm].astype(np.int32)); o += max_mol
            st.append(v[o:o + nm].astype(np.int32))
        status = np.concatenate(st)
        return dict(energies=np.concatenate(e), gradients=np.concatenate(g), charges=np.concatenate(q),
                    niter=np.concatenate(it), status=status, all_converged=bool((status == 0).all()))

    def singlepoint(self, positions, restart=False, dst=0):
        return self.gather(self.singlepoint_local(positions, restart), dst)
