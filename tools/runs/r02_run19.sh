#!/bin/bash
# one B200: the GPU test suite under the non-default engine switches
mkdir -p gpurun_out
for env in "FQ_STRICT_SELECT=0" "FQ_SLOTS=2 FQ_COPY_THREADS=1 FQ_POPULATE=0" "FQ_SLOTS=6 FQ_COPY_THREADS=4 FQ_PIPELINE_RAMP=0 FQ_DH_CHUNK_ROWS=100000" "FQ_ADAPT=0 FQ_STEAL=0 FQ_CHUNK_QUANT=37888"; do
  echo "== $env"; env $env timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -1
done
