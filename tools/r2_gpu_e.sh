#!/bin/bash
# 2 GPUs: default bench line under torchrun (ensemble + secondary incl. the water split with its in-bench shard check), split parity test
set -x
O=gpurun_out
nvidia-smi -L
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > $O/e_bench2.json 2> $O/e_bench2.err ) 2>&1 | tail -3
tail -c 400 $O/e_bench2.json; tail -5 $O/e_bench2.err
timeout 900 python -m pytest tests/test_large_shard.py tests/test_gpu_large_parity.py -x -q -m gpu 2>&1 | tail -5
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/e_ref2.json 2> $O/e_ref2.err ) 2>&1 | tail -3
tail -c 600 $O/e_ref2.json
