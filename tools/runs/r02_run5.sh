#!/bin/bash
# 2 GPUs: the driver's torchrun contract at N=2 (with the in-process phase), the multi-GPU parity test, pageable check
mkdir -p gpurun_out
nvidia-smi -L
timeout 120 python /tmp/pc3.py 2>&1 | tail -1 || true
cat > /tmp/pc3.py <<'PY'
import os, sys, time, numpy as np
sys.path.insert(0, os.getcwd())
import fourq_b200 as fq
n = 1 << 20
rng = np.random.default_rng(1)
k = rng.integers(0, 256, (n, 32), np.uint8)
pub = fq.MUL_base(rng.integers(0, 256, (n, 32), np.uint8))
pk = fq.pinned_empty((n, 32)); pk[:] = k
pp = fq.pinned_empty((n, 32)); pp[:] = pub
po = fq.pinned_empty((n, 32)); ps = fq.pinned_empty((n,))
o = np.zeros((n, 32), np.uint8); s = np.zeros((n,), np.uint8)
res = []
for name, args, kw in (("pin", (pk, pp), dict(out=po, status=ps)), ("page/fresh", (k, pub), {}), ("page/reused", (k, pub), dict(out=o, status=s))):
    for _ in range(2): fq.DH(*args, **kw)
    ts = []
    for _ in range(7):
        t = time.perf_counter(); fq.DH(*args, **kw); ts.append(time.perf_counter() - t)
    res.append("%s %.2f" % (name, float(np.median(ts)) * 1e3))
print("  ".join(res), flush=True)
PY
echo -n "defaults: "; timeout 120 python /tmp/pc3.py 2>&1 | tail -1
echo -n "no populate: "; FQ_POPULATE=0 timeout 120 python /tmp/pc3.py 2>&1 | tail -1
timeout 300 python -m pytest tests -m gpu -x -q -k "multi_gpu or pageable or concurrent" 2>&1 | tail -2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r5_bench_n2.json 2> gpurun_out/r5_bench_n2.err; echo "bench n2 rc=$?"; tail -5 gpurun_out/r5_bench_n2.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r5_bench_n2.json")); r=d["roofline"]
print("N=2: %.2f Mrows/s  %.3f ms  ladder frac %.4f e2e %.2f  pageable %.2f (%.3f)" % (d["value"]/1e6, d["ms_per_step"], r["frac"], d["e2e"]["value"]/1e6, d["e2e_pageable"]["value"]/1e6, d["e2e_pageable"]["frac_of_e2e"]))
print(json.dumps(d["inproc"], indent=1))
PY
