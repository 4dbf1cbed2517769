#!/bin/bash
set -x
O=gpurun_out
XTBB200_TUNING=1 XTBB200_LIB=$PWD/xtb_b200/lib/libxtb_b200_timing.so XTBB200_JB_VARIANT=1 timeout 300 python tools/jb_timing.py 148 > $O/c_timing_v1.log 2>&1
grep "sweeps 3" $O/c_timing_v1.log | head -6; tail -2 $O/c_timing_v1.log
timeout 600 python bench.py --workload ensemble --steps 3 --warmup 3 --secondary off > $O/c_bench.json 2> $O/c_bench.err
python - $O/c_bench.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
cb=d.get('cpu_baseline') or {}
print(sys.argv[1], round(d['value'],1), round(d['ms_per_step'],1), d['config'].get('converged'), cb.get('max_abs_dE_vs_gpu'), cb.get('niter_identical'), d['roofline'].get('avg_launch_ms'), d['roofline']['note'][-60:])
print(d['roofline']['share_all_kernels_warmup'])
PY
timeout 900 python -m pytest tests/test_gpu_hessian.py tests/test_gpu_round2.py tests/test_gpu_parity.py -x -q 2>&1 | tail -5
