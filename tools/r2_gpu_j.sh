#!/bin/bash
# ncu of the large-matrix eigensolver at n = 6000: symmetric panel kernel (third panel), trailing update and back-transformation GEMMs
set -x
O=gpurun_out
export EIG_ONLY=random
CMD="python tools/eig_check.py 6000"
ncu --set full --clock-control none --import-source on -k regex:k_sytrd_panel -s 2 -c 1 -o $O/r02_sytrd_panel_sym -f $CMD > $O/ncu_j1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_eg_gemm -s 2 -c 1 -o $O/r02_eg_gemm_update -f $CMD > $O/ncu_j2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_eg_gemm -s 200 -c 6 -o $O/r02_eg_gemm_bt -f $CMD > $O/ncu_j3.log 2>&1
tail -2 $O/ncu_j1.log $O/ncu_j2.log $O/ncu_j3.log
ls -la $O/*.ncu-rep | tail -4
