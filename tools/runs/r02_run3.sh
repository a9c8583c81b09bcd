#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --steps 10 --warmup 3 --cpu-sample 0 --verify-rows 0 --no-configs > gpurun_out/r3_bench_endo.json 2> gpurun_out/r3_bench_endo.err; echo "bench rc=$?"
for t in 1 3; do echo "== copy threads $t"; FQ_COPY_THREADS=$t FQ_TRACE=1 timeout 200 python tools/pageable_check.py 2> gpurun_out/r3_pageable_$t.err; grep "inside the call" gpurun_out/r3_pageable_$t.err | awk '{print $(NF-4)}' | tr '\n' ' '; echo; done
for r in 0; do echo "== copy threads 3, no ramp"; FQ_PIPELINE_RAMP=0 FQ_DH_CHUNK_ROWS=113664 FQ_COPY_THREADS=3 timeout 200 python tools/pageable_check.py 2>/dev/null; done
echo "== compare_ops"; timeout 300 python tools/compare_ops.py > gpurun_out/r3_compare.txt 2> gpurun_out/r3_compare.err; echo "rc=$?"; cat gpurun_out/r3_compare.txt
timeout 600 ncu --section SourceCounters -f -o gpurun_out/r3_compare python tools/compare_ops.py --rows 65536 --launch-only > gpurun_out/r3_ncu_compare.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r3_bench_endo.json")); r=d["roofline"]
print("%.2f Mrows/s  %.3f ms  ladder frac %.4f e2e %.2f  pageable %.2f (%.3f)  kernels %s" % (d["value"]/1e6, d["ms_per_step"], r["frac"], d["e2e"]["value"]/1e6, d["e2e_pageable"]["value"]/1e6, d["e2e_pageable"]["frac_of_e2e"], r["kernel_ms"]))
PY
