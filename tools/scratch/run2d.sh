#!/bin/bash
python -m pytest tests/test_gpu_sharded.py -q -k "exchange" 2>&1 | tail -2
export SMORE_VERBOSE=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
COMMON="--gpus 2 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e"
$TR bench.py $COMMON --parallelism sharded-exchange > gpurun_out/x2d_c2.json 2> gpurun_out/x2d_c2.err
grep -h "exchange rank 0" gpurun_out/x2d_c2.err | tail -2; cut -c1-120 gpurun_out/x2d_c2.json
$TR bench.py $COMMON --parallelism sharded-exchange --grow-graph --scale 3.5 > gpurun_out/x2d_grown.json 2> gpurun_out/x2d_grown.err
grep -h "exchange rank 0" gpurun_out/x2d_grown.err | tail -2; cut -c1-120 gpurun_out/x2d_grown.json
