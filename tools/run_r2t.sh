#!/bin/bash
# A/B of the cp.async row pipeline in k_line (SMORE_LINE_PIPE): bench graph (bench.py kernel-only + bench_models line_cpp)
# and the configs[4] per-GPU footprint (tools/bench_block.py), default build vs lib_nopipe.
mkdir -p gpurun_out
for v in default nopipe; do
  if [ $v = default ]; then unset SMORE_B200_LIB; else export SMORE_B200_LIB=$PWD/smore_b200/lib_$v/libsmore_b200.so; fi
  echo "== $v"
  timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'bench_value': d['value'], 'frac': d['roofline']['frac'], 'clocks': d['clocks']}))"
  timeout 200 python tools/bench_models.py --only line_cpp,mf --steps 3 --warmup 1 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(json.dumps({'model': d['model'], 'units_per_s': d['units_per_s'], 'frac': d['frac_of_measured_hbm']}))"
  timeout 200 python tools/bench_block.py 2>/dev/null | tail -1
done
unset SMORE_B200_LIB
timeout 300 python -m pytest tests/test_gpu_parity_f32.py "tests/test_zz_gpu_quality_gates.py::test_line_cpp" "tests/test_zz_gpu_quality_gates.py::test_full_occupancy_dim128_trained_to_convergence" tests/test_gpu_rotation.py -m gpu -q -p no:cacheprovider 2>&1 | tail -4
