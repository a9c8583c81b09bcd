#!/bin/bash
# developer helper run under gpurun (one GPU): plain bench first, then the ncu launch list and one --set full capture of the
# three DH kernels per algorithm.   usage: bash tools/profile_run.sh TAG
TAG=${1:-cur}
mkdir -p gpurun_out
set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_endo.json 2> gpurun_out/bench_${TAG}_endo.err || exit 1
python bench.py --algorithm windowed --steps 10 --warmup 3 --cpu-sample 0 > gpurun_out/bench_${TAG}_win.json 2> gpurun_out/bench_${TAG}_win.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --cpu-sample 0 --verify-rows 0 > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:k_dh_(prep|ladder)" -s 6 -c 2 -f -o gpurun_out/prof_${TAG}_endo \
    python bench.py --steps 2 --warmup 3 --cpu-sample 0 --verify-rows 0 > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:k_dh_(prep|ladder)" -s 6 -c 2 -f -o gpurun_out/prof_${TAG}_win \
    python bench.py --algorithm windowed --steps 2 --warmup 3 --cpu-sample 0 --verify-rows 0 > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out/*.ncu-rep
