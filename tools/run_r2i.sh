#!/bin/bash
# Round-2 GPU session i: full GPU suite on the final kernels, block kernel + bench with 3 vs 4 resident CTAs, per-model table.
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log)
tail -15 gpurun_out/r2i_pytest.log
for v in default mb4; do
  if [ $v = default ]; then unset SMORE_B200_LIB; else export SMORE_B200_LIB=$PWD/smore_b200/lib_$v/libsmore_b200.so; fi
  echo "== $v" >> gpurun_out/r2i_variants.txt
  timeout 200 python tools/bench_block.py >> gpurun_out/r2i_variants.txt 2>> gpurun_out/r2i.err
  timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e 2>> gpurun_out/r2i.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(json.dumps({'bench_value': d['value'], 'frac': d['roofline']['frac'], 'clocks': d['clocks']}))" >> gpurun_out/r2i_variants.txt
done
unset SMORE_B200_LIB
cat gpurun_out/r2i_variants.txt
timeout 700 python tools/bench_models.py > gpurun_out/r2i_bench_models.jsonl 2>> gpurun_out/r2i.err
python -c "
import json
for l in open('gpurun_out/r2i_bench_models.jsonl'):
    d=json.loads(l); print('%-22s %8.1f M/s  frac %.3f  tries %.2f' % (d['model'], d['units_per_s']/1e6, d['frac_of_measured_hbm'], d['mean_tries']))"
timeout 300 python bench.py > gpurun_out/r2i_bench.json 2>> gpurun_out/r2i.err; cat gpurun_out/r2i_bench.json
