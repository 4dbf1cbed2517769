#!/bin/bash
# session 5, call e: GPU test suite without the two large parity systems (their path is untouched by this session), apply with
# 2 / 1 row tiles per pass against the default 3, smoke, default bench line with the secondary workloads
O=gpurun_out
mkdir -p $O
( time timeout 420 python -m pytest tests -m gpu -q --durations=8 --deselect tests/test_gpu_large_parity.py ) > $O/e_tests.log 2>&1; tail -16 $O/e_tests.log
for v in apply2 apply1; do
  XTBB200_LIB=$PWD/xtb_b200/lib/libxtb_b200_$v.so timeout 100 python bench.py --secondary off --no-cpu --steps 3 --warmup 3 > $O/e_bench_$v.json 2> $O/e_bench_$v.err
done
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > $O/e_smoke.log 2>&1; tail -1 $O/e_smoke.log
( time timeout 330 python bench.py ) > $O/e_bench_default.json 2> $O/e_bench_default.err; tail -3 $O/e_bench_default.err
python - <<'PY'
import json
for v in ['default','apply2','apply1']:
    try:
        d=json.loads(open('gpurun_out/e_bench_%s.json'%v).read().strip().splitlines()[0])
        print(v, round(d['value'],1), round(d['e2e']['value'],1), 'jb_ms', round(d['roofline']['avg_launch_ms'],3), 'frac', round(d['roofline']['frac'],4), 'E', repr(d['sample_energy_Eh']), (d.get('cpu_baseline') or {}).get('max_abs_dE_vs_gpu'), (d.get('cpu_baseline') or {}).get('niter_identical'))
        for k,x in (d.get('secondary') or {}).items():
            print('   ', k, x.get('value'), x.get('unit'), (x.get('e2e') or {}).get('value'), x.get('error'))
    except Exception as e:
        print(v, 'failed', e)
PY
