#!/bin/bash
# end of round 2: ncu evidence of the kernels that changed after the first captures
set -x
O=gpurun_out
CMD="python tools/jb_timing.py 592"
ncu --set full --clock-control none --import-source on -k "regex:k_sdqh0|k_grad_h0" -c 2 -o $O/r02c_integrals -f $CMD > $O/ncu_c1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_gemm -s 12 -c 4 -o $O/r02c_gemm -f $CMD > $O/ncu_c2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/r02c_launches_ensemble592.csv $CMD > $O/ncu_c0.log 2>&1
EIG_ONLY=random ncu --set full --clock-control none --import-source on -k regex:k_sytrd_panel -s 2 -c 1 -o $O/r02c_sytrd_panel_sym -f python tools/eig_check.py 6000 > $O/ncu_c3.log 2>&1
EIG_ONLY=random ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/r02c_launches_eigh6000.csv python tools/eig_check.py 6000 > $O/ncu_c4.log 2>&1
tail -1 $O/ncu_c1.log $O/ncu_c3.log
