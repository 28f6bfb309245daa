#!/bin/bash
# Round-2 GPU session w (FINAL build): both bench arms, the ncu captures of the bench command (launch list + --set full of
# k_line) and one --set full capture of every other hot kernel on its bench_models.py workload -- each ncu pass only after
# the same command exited 0 without ncu. The .ncu-rep files are summarised ON THE BOX (profiles/summarize.py) and deleted:
# gpurun brings back at most 64 MiB.
mkdir -p gpurun_out/sum
export SUMMARIZE_OUT=$PWD/gpurun_out/sum
timeout 600 python bench.py --impl reference > gpurun_out/r2w_bench_reference.json 2> gpurun_out/r2w_bench_reference.err; echo "ref rc=$?"
timeout 600 python bench.py > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; echo "bench rc=$?"
cut -c1-200 gpurun_out/r2w_bench.json
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 200 $B > gpurun_out/r2w_b.log 2>&1 || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2w_launches.csv $B > gpurun_out/r2w_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_line -s 3 -c 1 -f -o gpurun_out/r2w_kline $B > gpurun_out/r2w_ncu2.log 2>&1
python profiles/summarize.py r2w gpurun_out/r2w_launches.csv gpurun_out/r2w_kline.ncu-rep k_line; rm -f gpurun_out/r2w_kline.ncu-rep
timeout 500 python tools/bench_models.py --only bpr_go_big,bpr_cpp_big,warp,hoprec,hpe,deepwalk --steps 2 --warmup 1 > gpurun_out/r2w_models.jsonl 2> gpurun_out/r2w_models.err || exit 1
for kv in "bpr_go_big k_bpr_go" "bpr_cpp_big k_bpr_cpp" "warp k_warp_fast" "hoprec k_hoprec_fast" "hpe k_hpe_fast" "deepwalk k_walk"; do
  set -- $kv
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$2 -s 1 -c 1 -f -o gpurun_out/r2w_$2 python tools/bench_models.py --only $1 --steps 1 --warmup 1 > gpurun_out/r2w_ncu_$2.log 2>&1
  python profiles/summarize.py r2w - gpurun_out/r2w_$2.ncu-rep $2; rm -f gpurun_out/r2w_$2.ncu-rep
done
ls -la gpurun_out/sum; du -sh gpurun_out
