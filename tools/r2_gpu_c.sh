#!/bin/bash
set -x
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py -x -q 2>&1 | tail -5
( time timeout 900 python bench.py --steps 3 --warmup 3 > $O/c_bench_full.json 2> $O/c_bench_full.err ) 2>&1 | tail -3
python - $O/c_bench_full.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(round(d['value'],1), d['e2e']['value'], d.get('calibration'))
for k,v in d['secondary'].items(): print(k, v.get('value'), v.get('unit'), v.get('e2e',{}).get('value'), v.get('calibration'), v.get('error'))
PY
tail -3 $O/c_bench_full.err
