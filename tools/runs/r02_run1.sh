#!/bin/bash
# round 2, GPU run 1 (one B200): parity tests, bench, ladder register-cap experiments, table-selection timing, ncu
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/r1_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r1_pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r1_bench_endo.json 2> gpurun_out/r1_bench_endo.err; echo "bench rc=$?"; tail -3 gpurun_out/r1_bench_endo.err
timeout 300 python bench.py --algorithm windowed --steps 10 --warmup 3 --cpu-sample 0 --no-configs > gpurun_out/r1_bench_win.json 2> gpurun_out/r1_bench_win.err; echo "bench win rc=$?"
FQ_STRICT_SELECT=0 timeout 300 python bench.py --steps 10 --warmup 3 --cpu-sample 0 --verify-rows 0 --no-configs > gpurun_out/r1_bench_endo_masked.json 2> gpurun_out/r1_bench_endo_masked.err; echo "bench masked rc=$?"
for v in masked255 strict255 strict224 strict200 strict184 strict168 strictfence; do echo "== $v"; timeout 120 ./tools/kexp/ml_$v; done > gpurun_out/r1_ladder_variants.txt 2>&1
cat gpurun_out/r1_ladder_variants.txt
timeout 300 python tools/ct_timing.py > gpurun_out/r1_ct_timing_strict.jsonl 2> gpurun_out/r1_ct.err; echo "ct rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches.csv \
    python bench.py --steps 2 --warmup 3 --cpu-sample 0 --verify-rows 0 --no-configs > gpurun_out/r1_ncu_l.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_dh_(prep|ladder|finish)" -s 9 -c 3 -f -o gpurun_out/r1_prof_endo \
    python bench.py --steps 2 --warmup 3 --cpu-sample 0 --verify-rows 0 --no-configs > gpurun_out/r1_ncu1.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/ | tail -20
python - <<'PY'
import json
for f in ("gpurun_out/r1_bench_endo.json","gpurun_out/r1_bench_win.json","gpurun_out/r1_bench_endo_masked.json"):
    try:
        d=json.load(open(f)); r=d["roofline"]
        print(f, "%.2f Mrows/s  %.3f ms  ladder frac %.4f step frac %.4f  e2e %.2f  pageable %.2f  kernels %s" % (d["value"]/1e6, d["ms_per_step"], r["frac"], r["step"]["frac"], d["e2e"]["value"]/1e6, d["e2e_pageable"]["value"]/1e6, r["kernel_ms"]))
        if d.get("configs"): print(json.dumps(d["configs"]))
    except Exception as e: print(f, "ERR", e)
PY
