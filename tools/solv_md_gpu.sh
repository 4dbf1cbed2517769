#!/bin/bash
# Measurement batch for the solvation (f2) and MD (f3) rows and the mid-size eigensolver: parity tests, smoke, bench lines,
# ncu launch list of MD steps.  Outputs go to gpurun_out/s_*.
set -x
O=gpurun_out
timeout 700 python -m pytest tests -m gpu -q > $O/s_tests.log 2>&1; tail -3 $O/s_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/s_smoke.log 2>&1; tail -1 $O/s_smoke.log
timeout 300 python bench.py --workload md --steps 3 --warmup 3 > $O/s_bench_md.json 2> $O/s_bench_md.err; head -c 200 $O/s_bench_md.json; tail -3 $O/s_bench_md.err
timeout 200 python bench.py --workload single --molecule caffeine --steps 5 --no-cpu > $O/s_bench_single_caffeine.json 2> $O/s_single.err; head -c 200 $O/s_bench_single_caffeine.json
timeout 400 python bench.py --workload single --molecule protein --steps 2 --warmup 1 --no-cpu > $O/s_bench_single_protein.json 2>> $O/s_single.err; head -c 200 $O/s_bench_single_protein.json
timeout 300 python bench.py > $O/s_bench_ensemble.json 2> $O/s_bench_ensemble.err; head -c 200 $O/s_bench_ensemble.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/s_ncu_launches_md.csv python tools/md_profile_step.py 2 > $O/s_ncu_md.log 2>&1; tail -1 $O/s_ncu_md.log
