#!/bin/bash
set -x
O=gpurun_out
nvidia-smi -L | wc -l
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 3 --warmup 3 > $O/e_bench2.json 2> $O/e_bench2.err ) 2>&1 | tail -3
python - $O/e_bench2.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d['value'],d['n_gpus'],d['ms_per_step'])
w=d['secondary']['water']; print({k:w[k] for k in ('value','n_gpus','phases_ms','comm_ms','shard_check')})
for k in ('caffeine','protein','md'):
    v=d['secondary'][k]; print(k, v.get('value'), v.get('n_gpus'), v.get('error'))
PY
tail -3 $O/e_bench2.err
timeout 600 python -m pytest tests/test_large_shard.py tests/test_gpu_large_parity.py -x -q -m gpu 2>&1 | tail -3
