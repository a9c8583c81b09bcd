#!/bin/bash
# 8 GPUs: the multi-GPU parity test, in-process scaling (with trace), the driver's torchrun contract at N = 8 and 4
mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc; free -g | head -2
timeout 300 python -m pytest tests -m gpu -x -q -k "multi_gpu" 2>&1 | tail -2
FQ_TRACE=1 timeout 300 python tools/inproc_scaling.py > gpurun_out/r8_inproc_scaling.jsonl 2> gpurun_out/r8_inproc_trace.err; echo "inproc rc=$?"; cat gpurun_out/r8_inproc_scaling.jsonl
for N in 8 4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r8_bench_n$N.json 2> gpurun_out/r8_bench_n$N.err; echo "bench n$N rc=$?"; tail -3 gpurun_out/r8_bench_n$N.err
done
python - <<'PY'
import json
for N in (8,4):
    try:
        d=json.load(open("gpurun_out/r8_bench_n%d.json"%N)); r=d["roofline"]
        print("N=%d: %.2f Mrows/s  %.3f ms  e2e %.2f  pageable %.2f (%.3f)" % (N, d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, d["e2e_pageable"]["value"]/1e6, d["e2e_pageable"]["frac_of_e2e"]))
        print(json.dumps(d["inproc"]))
    except Exception as e: print(N, "ERR", e)
PY
grep "inside the call" gpurun_out/r8_inproc_trace.err | tail -12
