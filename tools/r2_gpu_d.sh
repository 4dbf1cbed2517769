#!/bin/bash
set -x
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/d_tests.log 2>&1; tail -5 $O/d_tests.log
( time timeout 900 python bench.py --steps 5 --warmup 3 > $O/d_bench.json 2> $O/d_bench.err ) 2>&1 | tail -3; tail -c 300 $O/d_bench.json; tail -5 $O/d_bench.err
