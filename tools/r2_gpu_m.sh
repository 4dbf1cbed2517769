#!/bin/bash
set -x
O=gpurun_out
export XTBB200_TUNING=1
for v in 0 1; do
  if [ $v = 1 ]; then export XTBB200_CLUSTER_SINGLE=1; fi
  timeout 600 python bench.py --workload md --steps 3 --warmup 3 --no-cpu > $O/m_md$v.json 2> $O/m_md$v.err
  python - $O/m_md$v.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], d['value'], d['e2e']['value'], d['trajectory'], d['share_all_kernels'])
PY
  tail -2 $O/m_md$v.err
done
