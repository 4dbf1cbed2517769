#!/bin/bash
# session 5, call b: simultaneous-angle inner solve of k_jacobi_blk (default 0.1) against variants 0 (off), 0.03, 0.3
O=gpurun_out
mkdir -p $O
( time timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_hessian.py tests/test_gpu_decisions.py -x -q ) > $O/b_tests.log 2>&1; tail -4 $O/b_tests.log
timeout 200 python bench.py --secondary off --steps 3 --warmup 3 > $O/b_bench_default.json 2> $O/b_bench_default.err
for v in tau0.0 tau0.03 tau0.3; do
  XTBB200_LIB=$PWD/xtb_b200/lib/libxtb_b200_$v.so timeout 120 python bench.py --secondary off --no-cpu --steps 3 --warmup 3 > $O/b_bench_$v.json 2> $O/b_bench_$v.err
done
python - <<'PY'
import json
for v in ['default','tau0.0','tau0.03','tau0.3']:
    try:
        d=json.loads(open('gpurun_out/b_bench_%s.json'%v).read().strip().splitlines()[-1])
        print(v, round(d['value'],1), round(d['e2e']['value'],1), 'jb_ms', round(d['roofline']['avg_launch_ms'],3), 'frac', round(d['roofline']['frac'],4), 'E', repr(d['sample_energy_Eh']), d['config'].get('mean_scc_iterations'), (d.get('cpu_baseline') or {}).get('max_abs_dE_vs_gpu'), (d.get('cpu_baseline') or {}).get('niter_identical'), d['roofline']['note'][-90:])
    except Exception as e:
        print(v, 'failed', e)
PY
XTBB200_LIB=$PWD/xtb_b200/lib/libxtb_b200_timing.so timeout 120 python tools/jb_timing.py 148 > $O/b_jbtiming.log 2>&1; tail -12 $O/b_jbtiming.log
