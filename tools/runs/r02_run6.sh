#!/bin/bash
# one B200: by-value prep variant (stack 896 -> 352 B), decode with shared on-curve subexpressions, engine defaults
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r6_bench_endo.json 2> gpurun_out/r6_bench_endo.err; echo "bench rc=$?"; tail -3 gpurun_out/r6_bench_endo.err
timeout 300 python bench.py --algorithm windowed --steps 10 --warmup 3 --cpu-sample 0 --no-configs > gpurun_out/r6_bench_win.json 2> gpurun_out/r6_bench_win.err; echo "bench win rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r6_bench_endo.json","gpurun_out/r6_bench_win.json"):
    d=json.load(open(f)); r=d["roofline"]
    print(f, "%.2f Mrows/s  %.3f ms  ladder frac %.4f step %.4f e2e %.2f  pageable %.2f (%.3f)  kernels %s" % (d["value"]/1e6, d["ms_per_step"], r["frac"], r["step"]["frac"], d["e2e"]["value"]/1e6, d["e2e_pageable"]["value"]/1e6, d["e2e_pageable"]["frac_of_e2e"], r["kernel_ms"]))
PY
