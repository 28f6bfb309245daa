#!/bin/bash
# Round-2 GPU session z: train calls without copy-engine work (state initialised / reduced by kernels) + per-table ordering
# of the asynchronous row copies: the default bench line first, then the whole GPU suite (every trainer goes through
# init_state / collect_stats).
mkdir -p gpurun_out
timeout 400 python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2z_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], d['clocks'])
PY
(timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2z_pytest.log)
tail -4 gpurun_out/r2z_pytest.log
