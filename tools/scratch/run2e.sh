#!/bin/bash
export SMORE_VERBOSE=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
COMMON="--gpus 2 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e"
for R in 0 32; do
SMORE_EXCH_RESERVE_SMS=$R $TR bench.py $COMMON --parallelism sharded-exchange > gpurun_out/x2e_c2_r$R.json 2> gpurun_out/x2e_c2_r$R.err
grep -h "exchange rank 0" gpurun_out/x2e_c2_r$R.err | tail -1; cut -c1-120 gpurun_out/x2e_c2_r$R.json
done
