#!/bin/bash
# 4 GPUs, final build: the driver's torchrun contract at N = 4 and N = 2
mkdir -p gpurun_out
for N in 4 2; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r17_bench_n$N.json 2> gpurun_out/r17_bench_n$N.err; echo "bench n$N rc=$?"; tail -2 gpurun_out/r17_bench_n$N.err
done
python - <<'PY'
import json
for N in (4,2):
    d=json.load(open("gpurun_out/r17_bench_n%d.json"%N)); i=d["inproc"]
    print("N=%d: %.2f Mrows/s  %.3f ms  e2e %.2f  pageable %.2f (%.3f) | inproc cfg4 x%.2f cfg3 %.1f M (%.3f of e2e) copy %.0f GB/s" % (N, d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, d["e2e_pageable"]["value"]/1e6, d["e2e_pageable"]["frac_of_e2e"], i["cfg4"]["speedup"], i["cfg3"]["rows_per_s"]/1e6, i["cfg3"]["rows_per_s"]/d["e2e"]["value"], i["copy_probe"]["ndevN_host_gbs"]))
PY
