#!/bin/bash
# round 2, late: full GPU test suite, smoke, default bench line (with the secondary workloads) and the reference arm
set -x
O=gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/s_tests.log 2>&1; tail -6 $O/s_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/s_smoke.log 2>&1; tail -2 $O/s_smoke.log
( time timeout 900 python bench.py ) > $O/s_bench.json 2> $O/s_bench.err; tail -3 $O/s_bench.err
python - $O/s_bench.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('ensemble', d['value'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline'])
for k,v in d['secondary'].items():
    print(k, v.get('value'), v.get('unit'), (v.get('e2e') or {}).get('value'), (v.get('roofline') or {}).get('frac'), v.get('error'))
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/s_bench_ref.json 2> $O/s_bench_ref.err; tail -c 600 $O/s_bench_ref.json
