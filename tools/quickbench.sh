#!/bin/bash
# developer helper: both algorithms, 10 steps, no CPU sample; prints rows/s, ms, roofline fraction, e2e
python bench.py --steps 10 --warmup 3 --cpu-sample 0 --verify-rows 0 > gpurun_out/b_endo.json
python bench.py --algorithm windowed --steps 10 --warmup 3 --cpu-sample 0 --verify-rows 0 > gpurun_out/b_win.json
python - <<PY
import json
for f in ("gpurun_out/b_endo.json","gpurun_out/b_win.json"):
    d=json.load(open(f)); print(f, "%.2f Mrows/s  %.3f ms  frac %.4f  e2e %.2f" % (d["value"]/1e6, d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"]/1e6))
PY
