#!/bin/bash
set -x
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_eigh.py -x -q 2>&1 | tail -4
for v in 1 0; do
  if [ $v = 0 ]; then export XTBB200_NO_TMA_GEMM=1; fi
  timeout 600 python bench.py --workload water --steps 2 --warmup 1 --no-cpu > $O/h_water_tma$v.json 2> $O/h_water_tma$v.err
  python - $O/h_water_tma$v.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], d['value'], d['roofline'].get('dmma_tflops'), d['roofline'].get('eigensolver_ms_per_diagonalisation'), d['sample_energy_Eh'])
PY
done
