#!/bin/bash
set -x
O=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_solvation.py tests/test_gpu_md.py tests/test_cli_formats.py tests/test_gpu_decisions.py -x -q -m gpu > $O/g_tests.log 2>&1; tail -15 $O/g_tests.log
