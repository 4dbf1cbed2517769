#!/bin/bash
# symmetric trailing-matrix product of the panel kernel: accuracy and stage times
set -x
O=gpurun_out
export XTBB200_TUNING=1
( EIG_ONLY=random timeout 600 python tools/eig_check.py 2500 4000 6000 6000 ) 2>&1 | tail -8 > $O/i_eig_sym.log
cat $O/i_eig_sym.log
( timeout 900 python -m pytest tests/test_gpu_eigh.py -x -q ) 2>&1 | tail -4
