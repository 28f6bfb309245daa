#!/bin/bash
python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/pytest_gpu_final.log 2>&1; tail -14 gpurun_out/pytest_gpu_final.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err; tail -3 gpurun_out/bench_r1_final.err; cut -c1-300 gpurun_out/bench_r1_final.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1_final_ref.json 2> gpurun_out/bench_r1_final_ref.err; cut -c1-200 gpurun_out/bench_r1_final_ref.json
python tools/bench_exchange_local.py --world 2 --steps 1 > gpurun_out/xlocal_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1d_exchange.csv python tools/bench_exchange_local.py --world 2 --steps 0 --total 8388608 > gpurun_out/ncu_xlocal.log 2>&1
tail -2 gpurun_out/xlocal_plain.log
