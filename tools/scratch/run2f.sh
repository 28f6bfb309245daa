#!/bin/bash
export SMORE_VERBOSE=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
COMMON="--gpus 2 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e"
$TR bench.py $COMMON --parallelism sharded-exchange --grow-graph --scale 3.5 > gpurun_out/x2f_grown_exchange.json 2> gpurun_out/x2f_grown_exchange.err
grep -h "exchange rank 0" gpurun_out/x2f_grown_exchange.err | tail -1; cut -c1-120 gpurun_out/x2f_grown_exchange.json
$TR bench.py $COMMON --parallelism sharded-exchange --superbatch 4194304 > gpurun_out/x2f_c2_sb4m.json 2> gpurun_out/x2f_c2_sb4m.err
grep -h "exchange rank 0" gpurun_out/x2f_c2_sb4m.err | tail -1; cut -c1-120 gpurun_out/x2f_c2_sb4m.json
$TR bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/x2f_c2_peer.json 2> gpurun_out/x2f_c2_peer.err
cut -c1-120 gpurun_out/x2f_c2_peer.json
