#!/bin/bash
# Round-2 GPU session y: per-table ordering of the asynchronous row copies (the e2e leg's context-table upload now runs next
# to the previous step's vertex-table read-back): its test, then the default bench line.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_edge_cases.py -m gpu -q -p no:cacheprovider > gpurun_out/r2y_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2y_pytest.log
timeout 400 python bench.py > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2y_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], d['clocks'])
PY
