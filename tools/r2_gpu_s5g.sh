#!/bin/bash
# session 5, call g: phase counters of the final k_jacobi_blk (timing build), cross Gram matrix in bursts of one DMMA (gram1)
O=gpurun_out
mkdir -p $O
XTBB200_LIB=$PWD/xtb_b200/lib/libxtb_b200_timing.so timeout 60 python tools/jb_timing.py 148 > $O/g_jbtiming.log 2>&1; tail -6 $O/g_jbtiming.log
timeout 40 python bench.py --secondary off --no-cpu --steps 3 --warmup 3 > $O/g_bench_default.json 2> $O/g_bench_default.err
XTBB200_LIB=$PWD/xtb_b200/lib/libxtb_b200_gram1.so timeout 40 python bench.py --secondary off --no-cpu --steps 3 --warmup 3 > $O/g_bench_gram1.json 2> $O/g_bench_gram1.err
python - <<'PY'
import json
for v in ['default','gram1']:
    try:
        d=json.loads(open('gpurun_out/g_bench_%s.json'%v).read().strip().splitlines()[0])
        print(v, round(d['value'],1), 'jb_ms', round(d['roofline']['avg_launch_ms'],3), 'E', repr(d['sample_energy_Eh']))
    except Exception as e:
        print(v, 'failed', e)
PY
XTBB200_LIB=$PWD/xtb_b200/lib/libxtb_b200_gram1.so timeout 80 python -m pytest tests/test_gpu_parity.py -x -q > $O/g_tests.log 2>&1; tail -2 $O/g_tests.log
