#!/bin/bash
set -x
O=gpurun_out
timeout 600 python bench.py --workload water --steps 2 --warmup 1 --no-cpu > $O/n_water.json 2> $O/n_water.err
python - $O/n_water.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d['value'], d['roofline'].get('dmma_tflops'), d['roofline'].get('eigensolver_ms_per_diagonalisation'), d['sample_energy_Eh'], d['roofline']['frac'], d['phases_ms'])
print(d['share_all_kernels'])
PY
tail -3 $O/n_water.err
timeout 900 python bench.py --workload single --molecule protein --steps 3 --warmup 1 --no-cpu > $O/n_protein.json 2> $O/n_protein.err
python - $O/n_protein.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['roofline'].get('eigensolver_ms_per_diagonalisation'), d['sample_energy_Eh'])
PY
