#!/bin/bash
set -x
O=gpurun_out
timeout 900 python bench.py --workload single --molecule protein --steps 3 --warmup 1 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['sample_energy_Eh'], d['share_all_kernels'].get('eeq_guess'))"
timeout 600 python bench.py --workload water --steps 1 --warmup 1 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['sample_energy_Eh'], d['config']['scc_iterations'], d['phases_ms']['setup+integrals+S factorisation'], d['share_all_kernels'].get('eeq_guess'))"
