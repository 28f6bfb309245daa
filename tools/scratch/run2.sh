#!/bin/bash
# 2-GPU comparison: peer-access mode vs bulk-exchange mode, configs[1] graph and a grown graph
set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
COMMON="--gpus 2 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e"
$TR bench.py $COMMON --parallelism sharded-exchange > gpurun_out/x2_c2_exchange.json 2> gpurun_out/x2_c2_exchange.err
$TR bench.py $COMMON --parallelism sharded-exchange --hot-threshold -1 > gpurun_out/x2_c2_exchange_nohot.json 2> gpurun_out/x2_c2_exchange_nohot.err
$TR bench.py $COMMON --parallelism sharded-exchange --grow-graph --scale 3.5 > gpurun_out/x2_grown_exchange.json 2> gpurun_out/x2_grown_exchange.err
$TR bench.py $COMMON --parallelism sharded --grow-graph --scale 3.5 > gpurun_out/x2_grown_peer.json 2> gpurun_out/x2_grown_peer.err
tail -3 gpurun_out/x2_*.err
cat gpurun_out/x2_*.json | cut -c1-260
