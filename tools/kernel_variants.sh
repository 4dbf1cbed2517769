#!/bin/bash
# A/B of compile-time variants of the engine on the ensemble workload in ONE GPU call (how the k_jacobi_blk steps of the end of
# round 2 were measured: every variant is a separate library next to the default one, selected through XTBB200_LIB).
#   here (no GPU):   tools/kernel_variants.sh build name1 "-DMACRO_A=1" name2 "-DMACRO_B=2 -DMACRO_C" ...
#   on the GPU box:  gpurun -- 'bash tools/kernel_variants.sh run name1 name2 ...'   -> gpurun_out/v_bench_<name>.json, v_tests.log
set -u
ROOT=$(cd "$(dirname "$0")/.." && pwd)
LIB=$ROOT/xtb_b200/lib
NV="nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
mode=$1; shift
if [ "$mode" = build ]; then
  make -C $ROOT/xtb_b200/csrc > /dev/null || exit 1
  while [ $# -ge 2 ]; do
    name=$1; flags=$2; shift 2
    ( cd $ROOT/xtb_b200/csrc && $NV $flags -c engine.cu -o $LIB/engine_$name.o &&
      nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $LIB/libxtb_b200_$name.so $LIB/engine_$name.o $LIB/capi.o -Xlinker --exclude-libs,ALL &&
      cuobjdump -res-usage $LIB/libxtb_b200_$name.so 2>/dev/null | grep -A1 "k_jacobi_blkILi1" | tail -1 | sed "s/^/$name: /" ) &
  done
  wait
else
  O=$ROOT/gpurun_out; mkdir -p $O; cd $ROOT
  ( time timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -x -q ) > $O/v_tests.log 2>&1; tail -4 $O/v_tests.log
  timeout 120 python bench.py --secondary off --steps 3 --warmup 3 > $O/v_bench_default.json 2> $O/v_bench_default.err
  for v in "$@"; do
    XTBB200_LIB=$LIB/libxtb_b200_$v.so timeout 120 python bench.py --secondary off --no-cpu --steps 3 --warmup 3 > $O/v_bench_$v.json 2> $O/v_bench_$v.err
  done
  python - default "$@" <<'PY'
import json, sys
for v in sys.argv[1:]:
    try:
        d = json.loads(open('gpurun_out/v_bench_%s.json' % v).read().strip().splitlines()[0])
        cb = d.get('cpu_baseline') or {}
        print(v, round(d['value'], 1), 'molecules/s, top kernel', round(d['roofline']['avg_launch_ms'], 3), 'ms per launch, frac',
              round(d['roofline']['frac'], 4), 'E', repr(d['sample_energy_Eh']), 'dE vs oracle', cb.get('max_abs_dE_vs_gpu'),
              'niter identical', cb.get('niter_identical'))
    except Exception as e:
        print(v, 'failed', e)
PY
fi
