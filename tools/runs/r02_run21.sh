#!/bin/bash
# N GPUs, final build: full GPU test suite, then the driver's torchrun contract (both arms) at N and at the powers of two below it
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -1
M=$N
while [ $M -ge 2 ]; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $M --master-addr 127.0.0.1 --master-port 2954$M bench.py --impl reference --gpus $M --steps 3 --warmup 1 > gpurun_out/r21_ref_n$M.json 2> gpurun_out/r21_ref_n$M.err; echo "ref n$M rc=$?"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $M --master-addr 127.0.0.1 --master-port 2955$M bench.py --gpus $M --steps 10 --warmup 3 > gpurun_out/r21_bench_n$M.json 2> gpurun_out/r21_bench_n$M.err; echo "bench n$M rc=$?"; tail -2 gpurun_out/r21_bench_n$M.err
  M=$((M / 2))
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r21_bench_n*.json")):
    d = json.load(open(f)); i = d.get("inproc") or {}
    print("%s: %.2f Mrows/s  %.3f ms  e2e %.2f  pageable %.2f | inproc cfg4 x%s cfg3 %s" % (f, d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, d["e2e_pageable"]["value"]/1e6,
          i.get("cfg4", {}).get("speedup"), i.get("cfg3", {}).get("rows_per_s")))
for f in sorted(glob.glob("gpurun_out/r21_ref_n*.json")):
    print(f, open(f).read()[:200])
PY
