#!/bin/bash
# 8 GPUs: in-process scaling with adaptive slices / work stealing (and without, for the A/B), the driver's torchrun contract at N = 8
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "multi_gpu" 2>&1 | tail -2
FQ_TRACE=1 timeout 300 python tools/inproc_scaling.py > gpurun_out/r9_inproc_scaling.jsonl 2> gpurun_out/r9_inproc_trace.err; echo "inproc rc=$?"; cat gpurun_out/r9_inproc_scaling.jsonl
FQ_ADAPT=0 FQ_STEAL=0 timeout 300 python tools/inproc_scaling.py > gpurun_out/r9_inproc_scaling_static.jsonl 2>/dev/null; echo "static rc=$?"; cat gpurun_out/r9_inproc_scaling_static.jsonl
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r9_bench_n8.json 2> gpurun_out/r9_bench_n8.err; echo "bench n8 rc=$?"; tail -3 gpurun_out/r9_bench_n8.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r9_bench_n8.json"))
print("N=8: %.2f Mrows/s  %.3f ms  e2e %.2f  pageable %.2f (%.3f)" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, d["e2e_pageable"]["value"]/1e6, d["e2e_pageable"]["frac_of_e2e"]))
print(json.dumps(d["inproc"]))
PY
grep "op 28 dev" gpurun_out/r9_inproc_trace.err | tail -8
grep "op 23 dev" gpurun_out/r9_inproc_trace.err | tail -8
