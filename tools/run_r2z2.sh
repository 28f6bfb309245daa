#!/bin/bash
# Round-2 GPU session z2 (FINAL build, after the copy-free train calls): the ncu launch list of the bench command and one
# --set full capture of k_line, each only after the same command exited 0 without ncu; summarised on the box.
mkdir -p gpurun_out/sum
export SUMMARIZE_OUT=$PWD/gpurun_out/sum
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 200 $B > gpurun_out/r2z_b.log 2>&1 || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2z_launches.csv $B > gpurun_out/r2z_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_line -s 3 -c 1 -f -o gpurun_out/r2z_kline $B > gpurun_out/r2z_ncu2.log 2>&1
python profiles/summarize.py r2z gpurun_out/r2z_launches.csv gpurun_out/r2z_kline.ncu-rep k_line; rm -f gpurun_out/r2z_kline.ncu-rep
ls -la gpurun_out/sum
