#!/bin/bash
mkdir -p gpurun_out
python - <<'PY' > /tmp/pc.py
PY
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
for slots in 3 4 5 6; do for chunk in 454656 227328 151552; do
  echo -n "slots $slots chunk $chunk copy 3:  "; FQ_SLOTS=$slots FQ_DH_CHUNK_ROWS=$chunk FQ_COPY_THREADS=3 timeout 120 python /tmp/pc3.py 2>&1 | tail -1
done; done
echo -n "slots 4 chunk 227328 copy 2:  "; FQ_SLOTS=4 FQ_DH_CHUNK_ROWS=227328 FQ_COPY_THREADS=2 timeout 120 python /tmp/pc3.py 2>&1 | tail -1
echo -n "slots 5 chunk 227328 copy 4:  "; FQ_SLOTS=5 FQ_DH_CHUNK_ROWS=227328 FQ_COPY_THREADS=4 timeout 120 python /tmp/pc3.py 2>&1 | tail -1
echo -n "slots 4 chunk 227328 copy 3 noramp:  "; FQ_PIPELINE_RAMP=0 FQ_SLOTS=4 FQ_DH_CHUNK_ROWS=227328 FQ_COPY_THREADS=3 timeout 120 python /tmp/pc3.py 2>&1 | tail -1
echo -n "slots 6 chunk 151552 copy 3 noramp:  "; FQ_PIPELINE_RAMP=0 FQ_SLOTS=6 FQ_DH_CHUNK_ROWS=151552 FQ_COPY_THREADS=3 timeout 120 python /tmp/pc3.py 2>&1 | tail -1
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python bench.py --steps 10 --warmup 3 --cpu-sample 0 --verify-rows 0 --no-configs > gpurun_out/r4_bench_endo.json 2> gpurun_out/r4_bench_endo.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r4_bench_endo.json")); r=d["roofline"]
print("%.2f Mrows/s  %.3f ms  ladder frac %.4f e2e %.2f  pageable %.2f (%.3f)  kernels %s" % (d["value"]/1e6, d["ms_per_step"], r["frac"], d["e2e"]["value"]/1e6, d["e2e_pageable"]["value"]/1e6, d["e2e_pageable"]["frac_of_e2e"], r["kernel_ms"]))
PY
