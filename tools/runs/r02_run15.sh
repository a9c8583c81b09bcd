#!/bin/bash
# one B200: the committed state exactly as the driver will run it at round end (build check, GPU tests, smoke, both bench arms)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); print('build ok')"
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r15_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r15_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()"
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r15_bench_reference.json 2> gpurun_out/r15_ref.err; echo "ref rc=$?"
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/r15_bench.json 2> gpurun_out/r15_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r15_bench.json")); r=d["roofline"]
print("%.2f Mrows/s  %.3f ms  ladder frac %.4f step %.4f e2e %.2f  pageable %.2f (%.3f)  kernels %s  select %s" % (d["value"]/1e6, d["ms_per_step"], r["frac"], r["step"]["frac"], d["e2e"]["value"]/1e6, d["e2e_pageable"]["value"]/1e6, d["e2e_pageable"]["frac_of_e2e"], r["kernel_ms"], d["config"]["table_select"]))
print(sorted(d.keys()))
PY
