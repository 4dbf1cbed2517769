#!/bin/bash
# Measurement batch for the solvation (f2) and MD (f3) rows: parity tests, bench lines, ncu launch list and full captures of
# the two largest solvation kernels.  Outputs go to gpurun_out/s_*.
set -x
O=gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q > $O/s_tests.log 2>&1; tail -3 $O/s_tests.log
timeout 300 python bench.py --workload md --steps 3 --warmup 3 > $O/s_bench_md.json 2> $O/s_bench_md.err; tail -c 600 $O/s_bench_md.json; tail -3 $O/s_bench_md.err
timeout 300 python bench.py > $O/s_bench_ensemble.json 2> $O/s_bench_ensemble.err; tail -c 300 $O/s_bench_ensemble.json
timeout 300 python bench.py --solvent water --solvent-model alpb --no-cpu > $O/s_bench_ensemble_alpb.json 2> $O/s_bench_ensemble_alpb.err; tail -c 300 $O/s_bench_ensemble_alpb.json
SOLVENT=water timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/s_ncu_launches_ensemble592_alpb.csv python tools/profile_step.py 592 > $O/s_ncu1.log 2>&1; tail -1 $O/s_ncu1.log
SOLVENT=water timeout 300 ncu --set full --clock-control none --import-source on --kernel-name k_solv_sasa --launch-count 1 -o $O/s_solv_sasa -f python tools/profile_step.py 592 > $O/s_ncu2.log 2>&1; tail -1 $O/s_ncu2.log
SOLVENT=water timeout 300 ncu --set full --clock-control none --import-source on --kernel-name k_solv_matvec --launch-skip 3 --launch-count 1 -o $O/s_solv_matvec -f python tools/profile_step.py 592 > $O/s_ncu3.log 2>&1; tail -1 $O/s_ncu3.log
ls -la $O/s_*.ncu-rep
