#!/bin/bash
# session 5, call f (final): parity tests of the small-molecule path, default bench line with the secondary workloads, ncu launch
# list of one ensemble step and one full capture of k_jacobi_blk in its final state, pipelined single-tile apply for comparison
O=gpurun_out
mkdir -p $O
( time timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_hessian.py -x -q ) > $O/f_tests.log 2>&1; tail -4 $O/f_tests.log
( time timeout 200 python bench.py ) > $O/f_bench_default.json 2> $O/f_bench_default.err; tail -2 $O/f_bench_default.err
timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/f_launches_ensemble592.csv python tools/profile_step.py 592 > $O/f_ncu_l.log 2>&1; tail -1 $O/f_ncu_l.log
timeout 120 ncu --set full --clock-control none --import-source on -k regex:k_jacobi_blk -s 5 -c 1 -o $O/f_jacobi_blk -f python tools/profile_step.py 592 > $O/f_ncu_a.log 2>&1; tail -1 $O/f_ncu_a.log
XTBB200_LIB=$PWD/xtb_b200/lib/libxtb_b200_pipe.so timeout 60 python bench.py --secondary off --no-cpu --steps 3 --warmup 3 > $O/f_bench_pipe.json 2> $O/f_bench_pipe.err
python - <<'PY'
import json
for v in ['default','pipe']:
    try:
        d=json.loads(open('gpurun_out/f_bench_%s.json'%v).read().strip().splitlines()[0])
        print(v, round(d['value'],1), round(d['e2e']['value'],1), 'jb_ms', round(d['roofline']['avg_launch_ms'],3), 'frac', round(d['roofline']['frac'],4), 'E', repr(d['sample_energy_Eh']), (d.get('cpu_baseline') or {}).get('max_abs_dE_vs_gpu'), (d.get('cpu_baseline') or {}).get('niter_identical'))
        for k,x in (d.get('secondary') or {}).items():
            print('   ', k, x.get('value'), x.get('unit'), (x.get('e2e') or {}).get('value'), x.get('error'))
    except Exception as e:
        print(v, 'failed', e)
PY
ls -la $O/*.ncu-rep
