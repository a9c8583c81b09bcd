#!/bin/bash
# developer helper run under `gpurun --gpus 8`: the driver's launch contract at N = 8, 4, 2, 1 (short runs, no CPU sample)
mkdir -p gpurun_out
for N in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2950$N bench.py --gpus $N --steps 10 --warmup 3 --cpu-sample 0 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  echo "N=$N rc=$?"; tail -c 600 gpurun_out/scale_n$N.err; python -c "
import json;d=json.load(open('gpurun_out/scale_n$N.json'));print(d['n_gpus'],d['value']/1e6,d['e2e']['value']/1e6,d['ms_per_step'])"
done
python bench.py --steps 10 --warmup 3 --cpu-sample 0 > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
python -c "
import json;d=json.load(open('gpurun_out/scale_n1.json'));print(d['n_gpus'],d['value']/1e6,d['e2e']['value']/1e6,d['ms_per_step'])"
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1; lscpu | head -20 > gpurun_out/lscpu.txt
