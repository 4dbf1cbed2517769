#!/bin/bash
# session 5, call c: one rotation per lane in the simultaneous inner solve (default) against four per lane (rot4), sleeping poll
# loops, no DMMAs for row tiles beyond nr in the apply (tail); finer phase counters
O=gpurun_out
mkdir -p $O
( time timeout 300 python -m pytest tests/test_gpu_parity.py -x -q ) > $O/c_tests.log 2>&1; tail -4 $O/c_tests.log
timeout 200 python bench.py --secondary off --steps 3 --warmup 3 > $O/c_bench_default.json 2> $O/c_bench_default.err
for v in rot4 poll40 poll100 tail tailpoll; do
  XTBB200_LIB=$PWD/xtb_b200/lib/libxtb_b200_$v.so timeout 120 python bench.py --secondary off --no-cpu --steps 3 --warmup 3 > $O/c_bench_$v.json 2> $O/c_bench_$v.err
done
python - <<'PY'
import json
for v in ['default','rot4','poll40','poll100','tail','tailpoll']:
    try:
        d=json.loads(open('gpurun_out/c_bench_%s.json'%v).read().strip().splitlines()[-1])
        print(v, round(d['value'],1), round(d['e2e']['value'],1), 'jb_ms', round(d['roofline']['avg_launch_ms'],3), 'frac', round(d['roofline']['frac'],4), 'E', repr(d['sample_energy_Eh']), d['config'].get('mean_scc_iterations'), (d.get('cpu_baseline') or {}).get('max_abs_dE_vs_gpu'), (d.get('cpu_baseline') or {}).get('niter_identical'), d['roofline']['note'][-60:-40])
    except Exception as e:
        print(v, 'failed', e)
PY
XTBB200_LIB=$PWD/xtb_b200/lib/libxtb_b200_timing.so timeout 120 python tools/jb_timing.py 148 > $O/c_jbtiming.log 2>&1; tail -10 $O/c_jbtiming.log
