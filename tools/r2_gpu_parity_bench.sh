#!/bin/bash
set -x
O=gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_decisions.py -x -q ) 2>&1 | tail -4
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --secondary off 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['sample_energy_Eh'], d['ms_per_step'], d['roofline']['share_all_kernels_warmup'])"
