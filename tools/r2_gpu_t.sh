#!/bin/bash
# ncu evidence of the ensemble step after the integral-kernel rewrite (592 conformers)
set -x
O=gpurun_out
CMD="python tools/jb_timing.py 592"
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/r02b_launches_ensemble592.csv $CMD > $O/ncu_t0.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:k_sdqh0|k_grad_h0" -c 2 -o $O/r02b_integrals -f $CMD > $O/ncu_t1.log 2>&1
tail -2 $O/ncu_t1.log
