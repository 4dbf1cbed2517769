#!/bin/bash
# Measurement batch at the end of round 1: parity tests, smoke, bench lines (ensemble headline, MD, mid-size ensemble with and
# without the cluster path).  Outputs go to gpurun_out/s_*.
set -x
O=gpurun_out
timeout 700 python -m pytest tests -m gpu -q > $O/s_tests.log 2>&1; tail -3 $O/s_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/s_smoke.log 2>&1; tail -1 $O/s_smoke.log
timeout 300 python bench.py > $O/s_bench_ensemble.json 2> $O/s_bench_ensemble.err; head -c 200 $O/s_bench_ensemble.json
timeout 300 python bench.py --workload md --steps 3 --warmup 3 > $O/s_bench_md.json 2> $O/s_bench_md.err; head -c 200 $O/s_bench_md.json; tail -3 $O/s_bench_md.err
timeout 200 python bench.py --molecule taxol --nmol 128 --steps 2 --warmup 1 --no-cpu > $O/s_bench_taxol128_cluster.json 2> $O/s_bench_taxol.err; head -c 200 $O/s_bench_taxol128_cluster.json; tail -2 $O/s_bench_taxol.err
XTBB200_NO_CLUSTER_EIG=1 timeout 200 python bench.py --molecule taxol --nmol 128 --steps 1 --warmup 1 --no-cpu > $O/s_bench_taxol128_sequential.json 2>> $O/s_bench_taxol.err; head -c 200 $O/s_bench_taxol128_sequential.json
