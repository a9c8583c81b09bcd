#!/bin/bash
# N GPUs (2 or 8): multi-GPU tests, in-process scaling with adaptive slices + stealing (and without), torchrun contract at N
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "multi_gpu or sliced" 2>&1 | tail -2
FQ_TRACE=1 timeout 300 python tools/inproc_scaling.py > gpurun_out/r12_inproc_scaling_n$N.jsonl 2> gpurun_out/r12_inproc_trace_n$N.err; echo "inproc rc=$?"; cat gpurun_out/r12_inproc_scaling_n$N.jsonl
FQ_ADAPT=0 FQ_STEAL=0 timeout 300 python tools/inproc_scaling.py > gpurun_out/r12_inproc_scaling_static_n$N.jsonl 2>/dev/null; echo "static rc=$?"; cat gpurun_out/r12_inproc_scaling_static_n$N.jsonl
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r12_bench_n$N.json 2> gpurun_out/r12_bench_n$N.err; echo "bench n$N rc=$?"; tail -3 gpurun_out/r12_bench_n$N.err
python - <<PY
import json
d=json.load(open("gpurun_out/r12_bench_n$N.json"))
print("N=$N: %.2f Mrows/s  %.3f ms  e2e %.2f  pageable %.2f (%.3f)" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, d["e2e_pageable"]["value"]/1e6, d["e2e_pageable"]["frac_of_e2e"]))
print(json.dumps(d["inproc"]))
PY
grep "op 28 dev" gpurun_out/r12_inproc_trace_n$N.err | tail -8
grep "op 23 dev" gpurun_out/r12_inproc_trace_n$N.err | tail -8
