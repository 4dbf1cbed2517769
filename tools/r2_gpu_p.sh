#!/bin/bash
set -x
O=gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/r02_launches_caffeine.csv python tools/profile_step.py 1 caffeine 2 > $O/ncu_p.log 2>&1
tail -1 $O/ncu_p.log
