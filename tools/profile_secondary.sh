#!/bin/bash
# developer helper run under gpurun: ncu --set full of the fixed-base comb kernel, the X25519 ladder and the strict-scan DH ladder
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k "regex:^k_comb$" -s 6 -c 1 -f -o gpurun_out/prof_comb \
    python tests/checks/bench_configs.py --quick > gpurun_out/ncu_c.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:^k_x25519$" -s 2 -c 1 -f -o gpurun_out/prof_x25519 \
    python tests/checks/bench_configs.py --quick > gpurun_out/ncu_x.log 2>&1
FQ_STRICT_SELECT=1 ncu --set full --clock-control none --import-source on -k "regex:k_dh_ladder" -s 3 -c 1 -f -o gpurun_out/prof_ladder_strict \
    python bench.py --steps 2 --warmup 3 --cpu-sample 0 --verify-rows 0 > gpurun_out/ncu_s.log 2>&1
ls -la gpurun_out/prof_comb.ncu-rep gpurun_out/prof_x25519.ncu-rep gpurun_out/prof_ladder_strict.ncu-rep
