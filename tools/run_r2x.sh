#!/bin/bash
# Round-2 GPU session x (re-entry check of HEAD): smoke(), the default bench line, then the whole GPU suite the way the
# driver runs it (without -x so that every failure is listed). Run under gpurun.
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2x_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2x_smoke.log
timeout 400 python bench.py > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err; echo "bench rc=$?"
cut -c1-400 gpurun_out/r2x_bench.json
(timeout 1150 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=15 > gpurun_out/r2x_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2x_pytest.log)
tail -25 gpurun_out/r2x_pytest.log
