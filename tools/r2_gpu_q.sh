#!/bin/bash
set -x
O=gpurun_out
( EIG_ONLY=random timeout 600 python tools/eig_check.py 6000 3006 ) 2>&1 | tail -3
timeout 600 python bench.py --workload water --steps 2 --warmup 1 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['sample_energy_Eh'], d['config']['scc_iterations'], d['roofline']['eigensolver_ms_per_diagonalisation'], d['roofline']['dmma_tflops'])"
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --secondary off 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['sample_energy_Eh'], d['roofline']['share_all_kernels_warmup'])"
