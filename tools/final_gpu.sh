#!/bin/bash
# Round-end measurement batch (run on the GPU box through gpurun): parity tests, the bench lines of every workload and the
# ncu evidence that profiles/ summarises.  Outputs go to gpurun_out/f_*.
set -x
O=gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q > $O/f_tests.log 2>&1; tail -3 $O/f_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/f_smoke.log 2>&1; tail -1 $O/f_smoke.log
timeout 300 python bench.py > $O/f_bench_ensemble.json 2> $O/f_bench_ensemble.err; tail -c 400 $O/f_bench_ensemble.json
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > $O/f_bench_reference.json 2> $O/f_bench_reference.err; tail -c 300 $O/f_bench_reference.json
timeout 400 python bench.py --workload water --nwat 1000 --steps 2 --warmup 1 > $O/f_bench_water1000.json 2> $O/f_bench_water.err; tail -c 400 $O/f_bench_water1000.json
timeout 200 python bench.py --workload single --molecule caffeine --steps 5 > $O/f_bench_single_caffeine.json 2> $O/f_single.err; tail -c 300 $O/f_bench_single_caffeine.json
timeout 400 python bench.py --workload single --molecule protein --steps 2 --warmup 1 > $O/f_bench_single_protein.json 2>> $O/f_single.err; tail -c 300 $O/f_bench_single_protein.json
# ncu: launch lists (serialised, cold cache: shares only) and full captures of the dominant kernels
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/f_ncu_launches_ensemble592.csv python tools/profile_step.py 592 > $O/f_ncu1.log 2>&1; tail -1 $O/f_ncu1.log
EIG_ONLY=random timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/f_ncu_launches_eigh6000.csv python tools/eig_check.py 6000 > $O/f_ncu2.log 2>&1; tail -2 $O/f_ncu2.log
EIG_ONLY=random timeout 400 ncu --set full --clock-control none --import-source on --kernel-name k_sytrd_panel --launch-skip 2 --launch-count 1 -o $O/f_sytrd_head -f python tools/eig_check.py 6000 > $O/f_ncu3.log 2>&1; tail -1 $O/f_ncu3.log
EIG_ONLY=random timeout 400 ncu --set full --clock-control none --import-source on --kernel-name k_sytrd_smem --launch-count 1 -o $O/f_sytrd_smem -f python tools/eig_check.py 1472 > $O/f_ncu4.log 2>&1; tail -1 $O/f_ncu4.log
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name k_jacobi_blk --launch-skip 6 --launch-count 1 -o $O/f_jacobi_blk -f python tools/profile_step.py 592 > $O/f_ncu5.log 2>&1; tail -1 $O/f_ncu5.log
ls -la $O/*.ncu-rep
