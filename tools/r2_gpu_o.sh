#!/bin/bash
set -x
O=gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/o_tests.log 2>&1; tail -6 $O/o_tests.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --secondary off > $O/o_bench.json 2> $O/o_bench.err
python - $O/o_bench.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['share_all_kernels_warmup'])
PY
tail -3 $O/o_bench.err
