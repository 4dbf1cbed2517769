#!/bin/bash
# round 2, batch A: new regression tests, large parity test, the default bench line with its secondary lines
set -x
O=gpurun_out
nproc
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_midsize_batch.py -x -q > $O/a_tests_new.log 2>&1; tail -5 $O/a_tests_new.log
timeout 900 python -m pytest tests/test_gpu_large_parity.py -x -q -s > $O/a_tests_large.log 2>&1; tail -5 $O/a_tests_large.log
( time timeout 900 python bench.py --steps 5 --warmup 3 > $O/a_bench.json 2> $O/a_bench.err ) 2>&1 | tail -3; tail -c 600 $O/a_bench.json; tail -5 $O/a_bench.err
