#!/bin/bash
# developer helper run under `gpurun --gpus 8`: the driver's launch contract at N = 8, 4, 2, 1 (short runs, no CPU sample)
mkdir -p gpurun_out
: > gpurun_out/scale.jsonl
for N in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2950$N bench.py --gpus $N --steps 10 --warmup 3 --cpu-sample 0 --verify-rows 0 >> gpurun_out/scale.jsonl 2> gpurun_out/scale_n$N.err
  echo "N=$N rc=$?"
done
python bench.py --steps 10 --warmup 3 --cpu-sample 0 --verify-rows 0 >> gpurun_out/scale.jsonl 2> gpurun_out/scale_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 >> gpurun_out/scale.jsonl 2>> gpurun_out/scale_n8.err
python - <<'PY'
import json
for l in open('gpurun_out/scale.jsonl'):
    d=json.loads(l); print(d.get('impl','ours'), d['n_gpus'], '%.2fM'%(d['value']/1e6), 'e2e %.2fM'%(d['e2e']['value']/1e6), d.get('clocks'))
PY
