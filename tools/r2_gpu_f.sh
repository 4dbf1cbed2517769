#!/bin/bash
set -x
O=gpurun_out
timeout 1200 python -m pytest tests/test_gpu_decisions.py "tests/test_gpu_large_parity.py::test_polypeptide_full_single_point_matches_oracle" -x -q -s > $O/f_tests.log 2>&1; tail -15 $O/f_tests.log
timeout 600 python bench.py --workload single --molecule protein --steps 3 --warmup 1 > $O/f_protein.json 2> $O/f_protein.err; tail -c 1500 $O/f_protein.json; tail -3 $O/f_protein.err
