#!/bin/bash
# developer helper: e2e rows/s of fourq_b200.DH as a function of the pipeline chunk size
for c in 65536 131072 151552 189440 227328 262144 303104 524288; do
  FQ_DH_CHUNK_ROWS=$c python bench.py --steps 10 --warmup 3 --cpu-sample 0 --verify-rows 0 > gpurun_out/b_chunk.json
  python -c "
import json;d=json.load(open('gpurun_out/b_chunk.json'));print($c, 'value %.2fM e2e %.2fM' % (d['value']/1e6, d['e2e']['value']/1e6))"
done
