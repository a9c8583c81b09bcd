#!/bin/bash
# round 2, GPU run 2 (one B200): sign-tracking ladder, 256-bit row I/O, staging copies with 1/2/3 threads + trace
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_endo.json 2> gpurun_out/r2_bench_endo.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_endo.err
timeout 300 python bench.py --algorithm windowed --steps 10 --warmup 3 --cpu-sample 0 --no-configs > gpurun_out/r2_bench_win.json 2> gpurun_out/r2_bench_win.err; echo "bench win rc=$?"
for t in 1 2 3 4; do
  FQ_COPY_THREADS=$t FQ_TRACE=1 timeout 300 python bench.py --steps 6 --warmup 3 --cpu-sample 0 --verify-rows 0 --no-configs > gpurun_out/r2_bench_copy$t.json 2> gpurun_out/r2_bench_copy$t.err; echo "copy$t rc=$?"
done
python - <<'PY'
import json
for f in ["gpurun_out/r2_bench_endo.json","gpurun_out/r2_bench_win.json"]+["gpurun_out/r2_bench_copy%d.json"%t for t in (1,2,3,4)]:
    try:
        d=json.load(open(f)); r=d["roofline"]
        print(f, "%.2f Mrows/s  %.3f ms  ladder frac %.4f step frac %.4f  e2e %.2f  pageable %.2f (%.3f)  kernels %s" % (d["value"]/1e6, d["ms_per_step"], r["frac"], r["step"]["frac"], d["e2e"]["value"]/1e6, d["e2e_pageable"]["value"]/1e6, d["e2e_pageable"]["frac_of_e2e"], r["kernel_ms"]))
        if d.get("configs"): print(json.dumps(d["configs"]["cfg2"]), json.dumps(d["configs"]["cfg4"]))
    except Exception as e: print(f, "ERR", e)
PY
for t in 1 2; do echo "== trace copy threads $t (last 6 lines)"; grep "fq trace" gpurun_out/r2_bench_copy$t.err | tail -6; done
