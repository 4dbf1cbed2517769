#!/bin/bash
# session 5, call a: parity tests of the small-molecule path with the batched prologue of k_jacobi_blk, ensemble bench of the
# default build and of three variant builds (launch bound 608, pipelined apply), phase counters of the timing build
O=gpurun_out
mkdir -p $O
( time timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -x -q ) > $O/a_tests.log 2>&1; tail -4 $O/a_tests.log
timeout 200 python bench.py --secondary off --steps 3 --warmup 3 > $O/a_bench_default.json 2> $O/a_bench_default.err
for v in varB varC varD; do
  XTBB200_LIB=$PWD/xtb_b200/lib/libxtb_b200_$v.so timeout 120 python bench.py --secondary off --no-cpu --steps 3 --warmup 3 > $O/a_bench_$v.json 2> $O/a_bench_$v.err
done
python - <<'PY'
import json
for v in ['default','varB','varC','varD']:
    try:
        d=json.loads(open('gpurun_out/a_bench_%s.json'%v).read().strip().splitlines()[-1])
        print(v, round(d['value'],1), round(d['e2e']['value'],1), 'jb_ms', round(d['roofline']['avg_launch_ms'],3), 'frac', round(d['roofline']['frac'],4), 'E', repr(d['sample_energy_Eh']), d['config'].get('mean_scc_iterations'), (d.get('cpu_baseline') or {}).get('max_abs_dE_vs_gpu'), (d.get('cpu_baseline') or {}).get('niter_identical'))
    except Exception as e:
        print(v, 'failed', e)
PY
XTBB200_LIB=$PWD/xtb_b200/lib/libxtb_b200_timing.so timeout 120 python tools/jb_timing.py 148 > $O/a_jbtiming.log 2>&1; tail -40 $O/a_jbtiming.log
