#!/bin/bash
export SMORE_VERBOSE=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
COMMON="--gpus 2 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e"
for CFG in "8 16" "16 32" "32 32"; do
  set -- $CFG
  NCCL_MIN_P2P_NCHANNELS=$1 NCCL_MAX_P2P_NCHANNELS=$1 SMORE_EXCH_RESERVE_SMS=$2 $TR bench.py $COMMON --parallelism sharded-exchange > gpurun_out/x2_c2_ch$1_r$2.json 2> gpurun_out/x2_c2_ch$1_r$2.err
  echo "channels $1 reserve $2: $(cut -c1-120 gpurun_out/x2_c2_ch$1_r$2.json)"
done
NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32 SMORE_EXCH_RESERVE_SMS=32 $TR bench.py $COMMON --parallelism sharded-exchange --grow-graph --scale 3.5 > gpurun_out/x2_grown_ch32_r32.json 2> gpurun_out/x2_grown_ch32_r32.err
echo "grown ch32 r32: $(cut -c1-120 gpurun_out/x2_grown_ch32_r32.json)"
