#!/bin/bash
# one B200, final build: refresh the ncu evidence (launch list, --set full of the three DH kernels, SourceCounters of every variant) and the compare tables
mkdir -p gpurun_out
timeout 600 ncu --section SourceCounters -f -o gpurun_out/r16_compare python tools/compare_ops.py --rows 65536 --launch-only > gpurun_out/r16_ncu_compare.log 2>&1; echo "ncu compare rc=$?"
python tools/ncu_kernels_opmix.py gpurun_out/r16_compare.ncu-rep 65536 > gpurun_out/r16_compare_opmix.json; echo "opmix rc=$?"
timeout 300 python tools/compare_ops.py --opmix gpurun_out/r16_compare_opmix.json > gpurun_out/r16_compare.txt 2> gpurun_out/r16_compare.err; echo "compare rc=$?"
timeout 300 python bench.py --algorithm windowed --steps 10 --warmup 3 --cpu-sample 0 --no-configs > gpurun_out/r16_bench_win.json 2> gpurun_out/r16_bench_win.err; echo "bench win rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r16_launches.csv \
    python bench.py --steps 2 --warmup 3 --cpu-sample 0 --verify-rows 0 --no-configs > gpurun_out/r16_ncu_l.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_dh_(prep|ladder|finish)" -s 9 -c 3 -f -o gpurun_out/r16_prof_endo \
    python bench.py --steps 2 --warmup 3 --cpu-sample 0 --verify-rows 0 --no-configs > gpurun_out/r16_ncu1.log 2>&1; echo "ncu full rc=$?"
