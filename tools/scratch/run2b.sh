#!/bin/bash
export SMORE_VERBOSE=1
python -m pytest tests/test_gpu_sharded.py -q -k "exchange" 2>&1 | tail -3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
COMMON="--gpus 2 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e"
for R in 16 32 0; do
  SMORE_EXCH_RESERVE_SMS=$R $TR bench.py $COMMON --parallelism sharded-exchange > gpurun_out/x2_c2_exchange_r$R.json 2> gpurun_out/x2_c2_exchange_r$R.err
  grep -h "smore_b200\]" gpurun_out/x2_c2_exchange_r$R.err | head -1
  cut -c1-140 gpurun_out/x2_c2_exchange_r$R.json
done
SMORE_EXCH_RESERVE_SMS=16 $TR bench.py $COMMON --parallelism sharded-exchange --grow-graph --scale 3.5 > gpurun_out/x2_grown_exchange_r16.json 2> gpurun_out/x2_grown_exchange_r16.err
cut -c1-140 gpurun_out/x2_grown_exchange_r16.json
SMORE_EXCH_RESERVE_SMS=16 $TR bench.py $COMMON --parallelism sharded-exchange --superbatch 4194304 > gpurun_out/x2_c2_exchange_r16_sb4m.json 2> gpurun_out/x2_c2_exchange_r16_sb4m.err
cut -c1-140 gpurun_out/x2_c2_exchange_r16_sb4m.json
tail -n 3 gpurun_out/x2_c2_exchange_r16.err
