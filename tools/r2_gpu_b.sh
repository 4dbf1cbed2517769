#!/bin/bash
# round 2, batch B: block Jacobi variants A/B (ensemble bench, no secondary lines) + the full GPU test suite
set -x
O=gpurun_out
export XTBB200_TUNING=1
XTBB200_JB_VARIANT=1 timeout 600 python bench.py --workload ensemble --steps 3 --warmup 3 --secondary off > $O/b_bench_v1.json 2> $O/b_bench_v1.err
tail -c 1200 $O/b_bench_v1.json; tail -3 $O/b_bench_v1.err
XTBB200_JB_VARIANT=1 XTBB200_JB_INNER=2 timeout 600 python bench.py --workload ensemble --steps 3 --warmup 3 --secondary off --no-cpu > $O/b_bench_v1i2.json 2> $O/b_bench_v1i2.err
XTBB200_JB_VARIANT=1 XTBB200_JB_QTOL=1e-6 timeout 600 python bench.py --workload ensemble --steps 3 --warmup 3 --secondary off > $O/b_bench_v1q6.json 2> $O/b_bench_v1q6.err
unset XTBB200_TUNING
timeout 1500 python -m pytest tests -m gpu -x -q > $O/b_tests.log 2>&1; tail -5 $O/b_tests.log
