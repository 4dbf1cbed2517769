#!/bin/bash
set -x
O=gpurun_out
N=${1:-2}
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --workload water --steps 2 --warmup 1 --no-cpu > $O/w_water$N.json 2> $O/w_water$N.err ) 2>&1 | tail -3
python - $O/w_water$N.json <<'PY'
import json,sys
w=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k:w.get(k) for k in ('value','n_gpus','phases_ms','comm_ms','shard_check','sample_energy_Eh')})
print(w['roofline'].get('eigensolver_ms_per_diagonalisation'))
print(w['share_all_kernels'])
PY
tail -5 $O/w_water$N.err
