#!/bin/bash
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_final2.log 2>&1; tail -4 gpurun_out/pytest_gpu_final2.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1_final2.json 2> gpurun_out/bench_r1_final2.err; cut -c1-200 gpurun_out/bench_r1_final2.json
