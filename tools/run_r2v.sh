#!/bin/bash
# Round-2 GPU session v (FINAL build): the whole GPU suite, both bench arms, the ncu captures of the bench command (launch
# list + --set full of k_line) and one --set full capture of every other hot kernel on its bench_models.py workload. Each
# ncu pass only after the same command exited 0 without ncu. Run under gpurun.
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=6 > gpurun_out/r2v_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2v_pytest.log)
tail -3 gpurun_out/r2v_pytest.log
timeout 600 python bench.py --impl reference > gpurun_out/r2v_bench_reference.json 2> gpurun_out/r2v_bench_reference.err; echo "ref rc=$?"
timeout 600 python bench.py > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; echo "bench rc=$?"
cat gpurun_out/r2v_bench.json | cut -c1-300
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 200 $B > gpurun_out/r2v_b.log 2>&1 || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2v_launches.csv $B > gpurun_out/r2v_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_line -s 3 -c 1 -f -o gpurun_out/r2v_kline $B > gpurun_out/r2v_ncu2.log 2>&1
timeout 500 python tools/bench_models.py --only bpr_go_big,bpr_cpp_big,warp,hoprec,hpe,deepwalk --steps 2 --warmup 1 > gpurun_out/r2v_models.jsonl 2> gpurun_out/r2v_models.err || exit 1
cat gpurun_out/r2v_models.jsonl | cut -c1-200
for kv in "bpr_go_big k_bpr_go" "bpr_cpp_big k_bpr_cpp" "warp k_warp_fast" "hoprec k_hoprec_fast" "hpe k_hpe_fast" "deepwalk k_walk"; do
  set -- $kv
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$2 -s 1 -c 1 -f -o gpurun_out/r2v_$2 python tools/bench_models.py --only $1 --steps 1 --warmup 1 > gpurun_out/r2v_ncu_$2.log 2>&1
done
ls -la gpurun_out/r2v_*.ncu-rep
