#!/bin/bash
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --secondary off 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['roofline']['share_all_kernels_warmup']; print(d['value'], d['sample_energy_Eh'], d['ms_per_step'], {k:s[k] for k in s if 'h0' in k or 'fock' in k or 'pop' in k})"
