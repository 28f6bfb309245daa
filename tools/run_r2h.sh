#!/bin/bash
# Round-2 GPU session h: L2 prefetch modes (0 none, 1 per 128 B line, 2 bulk per row [default build], 3 per 32 B sector) on the
# block kernel at the configs[4] footprint, on bench.py's configs[1] and on the ranking kernels.
mkdir -p gpurun_out
for v in default pm0 pm1 pm3; do
  if [ $v = default ]; then unset SMORE_B200_LIB; else export SMORE_B200_LIB=$PWD/smore_b200/lib_$v/libsmore_b200.so; fi
  echo "== $v" >> gpurun_out/r2h_prefetch_modes.txt
  timeout 200 python tools/bench_block.py >> gpurun_out/r2h_prefetch_modes.txt 2>> gpurun_out/r2h.err
  timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e 2>> gpurun_out/r2h.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(json.dumps({'bench_value': d['value'], 'frac': d['roofline']['frac'], 'clocks': d['clocks']}))" >> gpurun_out/r2h_prefetch_modes.txt
  timeout 300 python tools/bench_models.py --only bpr_go_big,bpr_cpp_big,mf --steps 2 --warmup 1 2>> gpurun_out/r2h.err | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(json.dumps({'model': d['model'], 'units_per_s': d['units_per_s'], 'frac': d['frac_of_measured_hbm']}))" >> gpurun_out/r2h_prefetch_modes.txt
done
cat gpurun_out/r2h_prefetch_modes.txt
