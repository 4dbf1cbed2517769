#!/bin/bash
set -x
O=gpurun_out
export EIG_ONLY=random
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/r02_launches_eigh6000.csv python tools/eig_check.py 6000 > $O/ncu_l2.log 2>&1
tail -2 $O/ncu_l2.log
