#!/bin/bash
# one B200, final evidence run of the round: parity tests, bench (both algorithms + reference arm), compare tool, ct timing, ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r10_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r10_pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r10_bench_endo.json 2> gpurun_out/r10_bench_endo.err; echo "bench rc=$?"; tail -3 gpurun_out/r10_bench_endo.err
timeout 300 python bench.py --algorithm windowed --steps 10 --warmup 3 --cpu-sample 0 --no-configs > gpurun_out/r10_bench_win.json 2> gpurun_out/r10_bench_win.err; echo "bench win rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r10_bench_reference.json 2> gpurun_out/r10_bench_reference.err; echo "ref rc=$?"
timeout 300 python tools/compare_ops.py --opmix profiles/r02_compare_opmix.json > gpurun_out/r10_compare.txt 2> gpurun_out/r10_compare.err; echo "compare rc=$?"
timeout 300 python tools/ct_timing.py > gpurun_out/r10_ct_timing.jsonl 2> gpurun_out/r10_ct.err; echo "ct rc=$?"
timeout 200 python tools/pageable_check.py > gpurun_out/r10_pageable.txt 2>&1; cat gpurun_out/r10_pageable.txt
python -c "import __graft_entry__ as g; g.smoke()"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r10_launches.csv \
    python bench.py --steps 2 --warmup 3 --cpu-sample 0 --verify-rows 0 --no-configs > gpurun_out/r10_ncu_l.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_dh_(prep|ladder|finish)" -s 9 -c 3 -f -o gpurun_out/r10_prof_endo \
    python bench.py --steps 2 --warmup 3 --cpu-sample 0 --verify-rows 0 --no-configs > gpurun_out/r10_ncu1.log 2>&1; echo "ncu full rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r10_bench_endo.json","gpurun_out/r10_bench_win.json"):
    d=json.load(open(f)); r=d["roofline"]
    print(f, "%.2f Mrows/s  %.3f ms  ladder frac %.4f step %.4f e2e %.2f  pageable %.2f (%.3f)  kernels %s" % (d["value"]/1e6, d["ms_per_step"], r["frac"], r["step"]["frac"], d["e2e"]["value"]/1e6, d["e2e_pageable"]["value"]/1e6, d["e2e_pageable"]["frac_of_e2e"], r["kernel_ms"]))
print(open("gpurun_out/r10_bench_reference.json").read()[:400])
PY
echo "== affine-table ladder experiment"; timeout 120 ./tools/kexp/ml_affine
echo "== strict ladder reference"; timeout 120 ./tools/kexp/ml_strict255
echo "== compute-sanitizer attempt"; timeout 300 compute-sanitizer --tool memcheck python tests/checks/allkernels_check.py > gpurun_out/r10_sanitizer.log 2>&1; echo "sanitizer rc=$?"; tail -5 gpurun_out/r10_sanitizer.log
timeout 120 python tests/checks/allkernels_check.py
