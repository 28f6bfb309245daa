#!/bin/bash
# Round-2 GPU session g: block-kernel variants at the configs[4] per-GPU footprint, then the ncu captures of the round
# (launch list + --set full of k_line at both footprints and of k_bpr_go on the Zipf catalogue). Run under gpurun.
mkdir -p gpurun_out
for v in default pf4 pf6 mb4 mb4pf4; do
  if [ $v = default ]; then unset SMORE_B200_LIB; else export SMORE_B200_LIB=$PWD/smore_b200/lib_$v/libsmore_b200.so; fi
  timeout 200 python tools/bench_block.py >> gpurun_out/r2g_block_variants.jsonl 2>> gpurun_out/r2g_block_variants.err
done
unset SMORE_B200_LIB
timeout 200 python tools/bench_block.py --split >> gpurun_out/r2g_block_variants.jsonl 2>> gpurun_out/r2g_block_variants.err
timeout 200 python tools/bench_block.py --scale 0.25 >> gpurun_out/r2g_block_variants.jsonl 2>> gpurun_out/r2g_block_variants.err
cat gpurun_out/r2g_block_variants.jsonl
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 200 $B > gpurun_out/r2g_b.log 2>&1 || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2g_launches.csv $B > gpurun_out/r2g_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_line -s 3 -c 1 -f -o gpurun_out/r2g_kline $B > gpurun_out/r2g_ncu2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_line -s 6 -c 1 -f -o gpurun_out/r2g_kline_block python tools/bench_block.py --episodes 4 > gpurun_out/r2g_ncu3.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_bpr_go -s 1 -c 1 -f -o gpurun_out/r2g_bprgo python tools/bench_models.py --only bpr_go_big --steps 1 --warmup 1 > gpurun_out/r2g_ncu4.log 2>&1
ls -la gpurun_out/*.ncu-rep
