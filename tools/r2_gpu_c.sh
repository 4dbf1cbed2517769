#!/bin/bash
set -x
O=gpurun_out
timeout 600 python bench.py --workload ensemble --steps 3 --warmup 3 --secondary off > $O/c_bench.json 2> $O/c_bench.err
python - $O/c_bench.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
cb=d.get('cpu_baseline') or {}
print(sys.argv[1], round(d['value'],1), round(d['ms_per_step'],1), d['config'].get('converged'), cb.get('max_abs_dE_vs_gpu'), cb.get('niter_identical'), d['roofline'].get('avg_launch_ms'), d['roofline'].get('note')[-70:])
PY
timeout 900 python -m pytest tests/test_gpu_hessian.py tests/test_gpu_parity.py -x -q 2>&1 | tail -3
