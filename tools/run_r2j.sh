#!/bin/bash
# Round-2 GPU session j: A/B of the item-grouped BPR kernel and of 3 resident CTAs for the walk kernels; MF vs concurrency.
mkdir -p gpurun_out
fmt='import json,sys
for l in sys.stdin:
    d=json.loads(l); print("%-22s %8.1f M/s  frac %.3f" % (d["model"], d["units_per_s"]/1e6, d["frac_of_measured_hbm"]))'
for v in default nogrp; do
  if [ $v = default ]; then unset SMORE_B200_LIB; else export SMORE_B200_LIB=$PWD/smore_b200/lib_$v/libsmore_b200.so; fi
  echo "== $v" >> gpurun_out/r2j_ab.txt
  timeout 400 python tools/bench_models.py --only bpr_go,bpr_go_big 2>> gpurun_out/r2j.err | python -c "$fmt" >> gpurun_out/r2j_ab.txt
done
for v in default walk3; do
  if [ $v = default ]; then unset SMORE_B200_LIB; else export SMORE_B200_LIB=$PWD/smore_b200/lib_$v/libsmore_b200.so; fi
  echo "== $v" >> gpurun_out/r2j_ab.txt
  timeout 400 python tools/bench_models.py --only deepwalk,walklets,hpe 2>> gpurun_out/r2j.err | python -c "$fmt" >> gpurun_out/r2j_ab.txt
done
unset SMORE_B200_LIB
cat gpurun_out/r2j_ab.txt
timeout 300 python tools/mf_occupancy_probe.py > gpurun_out/r2j_mf_occupancy.txt 2>> gpurun_out/r2j.err; cat gpurun_out/r2j_mf_occupancy.txt
timeout 300 python -m pytest tests/test_gpu_parity_f32.py tests/test_gpu_parity.py -q -p no:cacheprovider -k "bpr" 2>&1 | tail -3
timeout 300 python -m pytest tests/test_zz_gpu_quality_gates.py -q -p no:cacheprovider -k "bpr" 2>&1 | tail -3
