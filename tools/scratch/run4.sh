#!/bin/bash
export SMORE_VERBOSE=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511"
COMMON="--gpus 4 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e"
timeout 70 $TR bench.py $COMMON > gpurun_out/x4_c2_peer.json 2> gpurun_out/x4_c2_peer.err
cut -c1-120 gpurun_out/x4_c2_peer.json
timeout 70 $TR bench.py $COMMON --parallelism sharded-exchange > gpurun_out/x4_c2_exchange.json 2> gpurun_out/x4_c2_exchange.err
grep -h "exchange rank 0" gpurun_out/x4_c2_exchange.err | tail -1 | cut -c1-330; cut -c1-120 gpurun_out/x4_c2_exchange.json
