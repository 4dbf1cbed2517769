#!/bin/bash
# session 5, call d: rotation evaluated ahead of the votes (default) against after them (late), 6 or 3 row tiles per pass of the apply, 104 registers
O=gpurun_out
mkdir -p $O
( time timeout 300 python -m pytest tests/test_gpu_parity.py -x -q ) > $O/d_tests.log 2>&1; tail -4 $O/d_tests.log
timeout 200 python bench.py --secondary off --steps 3 --warmup 3 > $O/d_bench_default.json 2> $O/d_bench_default.err
for v in late apply6 apply3 reg104 reg104a6; do
  XTBB200_LIB=$PWD/xtb_b200/lib/libxtb_b200_$v.so timeout 120 python bench.py --secondary off --no-cpu --steps 3 --warmup 3 > $O/d_bench_$v.json 2> $O/d_bench_$v.err
done
python - <<'PY'
import json
for v in ['default','late','apply6','apply3','reg104','reg104a6']:
    try:
        d=json.loads(open('gpurun_out/d_bench_%s.json'%v).read().strip().splitlines()[-1])
        print(v, round(d['value'],1), round(d['e2e']['value'],1), 'jb_ms', round(d['roofline']['avg_launch_ms'],3), 'frac', round(d['roofline']['frac'],4), 'E', repr(d['sample_energy_Eh']), d['config'].get('mean_scd_iterations'), (d.get('cpu_baseline') or {}).get('max_abs_dE_vs_gpu'), (d.get('cpu_baseline') or {}).get('niter_identical'), d['roofline']['note'][-60:-40])
    except Exception as e:
        print(v, 'failed', e)
PY
XTBB200_LIB=$PWD/xtb_b200/lib/libxtb_b200_timing.so timeout 120 python tools/jb_timing.py 148 > $O/d_jbtiming.log 2>&1; tail -10 $O/d_jbtiming.log
