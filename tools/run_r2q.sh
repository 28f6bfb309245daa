#!/bin/bash
# Round-2 GPU session q (final kernels): per-trainer throughput (incl. CPR / TPR), then BASELINE.json configs[2] and
# configs[3] at their named sizes. Run under gpurun.
mkdir -p gpurun_out
timeout 500 python tools/bench_models.py --steps 3 --warmup 1 > gpurun_out/r2q_bench_models.jsonl 2> gpurun_out/r2q_bench_models.err; echo "models rc=$?"
cat gpurun_out/r2q_bench_models.jsonl
timeout 400 python tools/bench_models.py --scale 5 --only deepwalk,walklets --steps 3 --warmup 1 > gpurun_out/r2q_c2.jsonl 2> gpurun_out/r2q_c2.err; echo "c2 rc=$?"
cat gpurun_out/r2q_c2.jsonl
timeout 600 python tools/bench_models.py --c4-full --only warp,bpr_cpp --steps 3 --warmup 1 > gpurun_out/r2q_c3_directed.jsonl 2> gpurun_out/r2q_c3_directed.err; echo "c3 directed rc=$?"
cat gpurun_out/r2q_c3_directed.jsonl; tail -3 gpurun_out/r2q_c3_directed.err
timeout 700 python tools/bench_models.py --c4-full --only hoprec --steps 3 --warmup 1 > gpurun_out/r2q_c3_hoprec.jsonl 2> gpurun_out/r2q_c3_hoprec.err; echo "c3 hoprec rc=$?"
cat gpurun_out/r2q_c3_hoprec.jsonl; tail -3 gpurun_out/r2q_c3_hoprec.err
