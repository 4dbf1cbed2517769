#!/bin/bash
# round 2: ncu evidence for the ensemble step (592 rivaroxaban conformers = 4 waves of the one-CTA-per-SM kernels)
set -x
O=gpurun_out
CMD="python tools/jb_timing.py 592"
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/r02_launches_ensemble592.csv $CMD > $O/ncu_l.log 2>&1
tail -2 $O/ncu_l.log
ncu --set full --clock-control none --import-source on -k regex:k_jacobi_blk -s 5 -c 1 -o $O/r02_jacobi_blk -f $CMD > $O/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:k_sdqh0|k_grad_h0" -c 2 -o $O/r02_integrals -f $CMD > $O/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:k_pop_rows|k_fock|k_scc_finish" -s 9 -c 3 -o $O/r02_scc_kernels -f $CMD > $O/ncu_c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_gemm -s 12 -c 4 -o $O/r02_gemm -f $CMD > $O/ncu_d.log 2>&1
tail -3 $O/ncu_a.log $O/ncu_d.log
ls -la $O/*.ncu-rep
