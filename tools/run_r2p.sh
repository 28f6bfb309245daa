#!/bin/bash
# Round-2 GPU session p (final kernels): the whole GPU suite, both bench arms, and the ncu captures of the bench command
# (launch list + --set full of k_line), each ncu pass only after its command exited 0 without ncu. Run under gpurun.
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=8 > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2p_pytest.log)
tail -4 gpurun_out/r2p_pytest.log
timeout 600 python bench.py --impl reference > gpurun_out/r2p_bench_reference.json 2> gpurun_out/r2p_bench_reference.err; echo "ref rc=$?"
timeout 600 python bench.py > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; echo "bench rc=$?"
cat gpurun_out/r2p_bench_reference.json gpurun_out/r2p_bench.json
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 200 $B > gpurun_out/r2p_b.log 2>&1 || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2p_launches.csv $B > gpurun_out/r2p_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_line -s 3 -c 1 -f -o gpurun_out/r2p_kline $B > gpurun_out/r2p_ncu2.log 2>&1
ls -la gpurun_out/*.ncu-rep
