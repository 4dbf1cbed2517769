#!/bin/bash
# session 5, call h (last seconds of the budget): division-free pair indices + warp-per-column epilogue of k_jacobi_blk against the build before
O=gpurun_out
mkdir -p $O
timeout 22 python bench.py --secondary off --no-cpu --steps 3 --warmup 3 > $O/h_bench_default.json 2> $O/h_bench_default.err
XTBB200_LIB=$PWD/xtb_b200/lib/libxtb_b200_prev.so timeout 22 python bench.py --secondary off --no-cpu --steps 3 --warmup 3 > $O/h_bench_prev.json 2> $O/h_bench_prev.err
python - <<'PY'
import json
for v in ['default','prev']:
    try:
        d=json.loads(open('gpurun_out/h_bench_%s.json'%v).read().strip().splitlines()[0])
        print(v, round(d['value'],1), 'jb_ms', round(d['roofline']['avg_launch_ms'],3), 'E', repr(d['sample_energy_Eh']), d['config'].get('mean_scc_iterations'), d['config'].get('converged'))
    except Exception as e:
        print(v, 'failed', e)
PY
timeout 15 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
