#!/usr/bin/env python3
"""Multi-GPU split of one large system: sharded result vs the single-GPU result of the same engine.
Run on the GPU box:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
    --master-port 29511 tools/shard_check.py water 64
Prints 'SHARD OK ...' on rank 0 when the energies are bit-identical and the gradients agree to 1e-11."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import xtb_b200  # noqa: E402
from bench import load_kat, make_peptide_cluster, make_water_cluster  # noqa: E402
from xtb_b200.sharded import ShardedCalculator  # noqa: E402

kind, k = sys.argv[1], int(sys.argv[2])
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
Z, xyz = make_water_cluster(k) if kind == "water" else make_peptide_cluster(load_kat("pdb_4qxx"), k)
sc = ShardedCalculator(Z, rank, world, local)
r = sc.singlepoint(xyz)
t = sc.calc.timings()
comm = sc.calc.comm_ms()
res = np.concatenate([[r["energies"][0], float(r["niter"][0])], r["gradients"].ravel(), r["charges"]])
if world > 1:
    all_res = [torch.zeros(res.size, dtype=torch.float64, device="cuda") for _ in range(world)]
    dist.all_gather(all_res, torch.tensor(res, device="cuda"))
    all_res = [a.cpu().numpy() for a in all_res]
else:
    all_res = [res]
if rank == 0:
    ref = xtb_b200.BatchCalculator([Z], device=local)
    rr = ref.singlepoint(xyz)
    tr = ref.timings()
    ref_res = np.concatenate([[rr["energies"][0], float(rr["niter"][0])], rr["gradients"].ravel(), rr["charges"]])
    ok = True
    for w, a in enumerate(all_res):
        de = abs(a[0] - ref_res[0])
        dn = abs(a[1] - ref_res[1])
        dg = np.max(np.abs(a[2:2 + 3 * len(Z)] - ref_res[2:2 + 3 * len(Z)]))
        dq = np.max(np.abs(a[2 + 3 * len(Z):] - ref_res[2 + 3 * len(Z):]))
        print("rank %d: dE %.3e dniter %d dG %.3e dq %.3e" % (w, de, int(dn), dg, dq))
        ok &= (de == 0.0 and dn == 0 and dg < 1e-11 and dq == 0.0)
    print("sharded   x%d: setup %.1f ms scc %.1f ms grad %.1f ms  (exchanges %.2f ms)" % (world, t["setup_ms"], t["scc_ms"], t["grad_ms"], comm))
    print("single GPU  : setup %.1f ms scc %.1f ms grad %.1f ms" % (tr["setup_ms"], tr["scc_ms"], tr["grad_ms"]))
    print(("SHARD OK" if ok else "SHARD MISMATCH") + " %s %d natoms %d world %d E %.10f" % (kind, k, len(Z), world, rr["energies"][0]))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
