#!/bin/bash
# developer helper run under gpurun (one GPU): everything profiles/ quotes for one build.   usage: bash tools/final_run.sh TAG
TAG=${1:-cur}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_${TAG}.log 2>&1; echo "pytest rc=$?"
bash tools/profile_run.sh $TAG 2> gpurun_out/profile_run.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2>/dev/null
FQ_STRICT_SELECT=1 python bench.py --steps 10 --warmup 3 --cpu-sample 0 --verify-rows 0 > gpurun_out/bench_${TAG}_endo_strict.json 2>/dev/null
FQ_STRICT_SELECT=1 python bench.py --algorithm windowed --steps 10 --warmup 3 --cpu-sample 0 --verify-rows 0 > gpurun_out/bench_${TAG}_win_strict.json 2>/dev/null
python tests/checks/bench_configs.py > gpurun_out/configs_${TAG}.jsonl 2> gpurun_out/configs_${TAG}.err
FQ_STRICT_SELECT=1 python tests/checks/bench_configs.py > gpurun_out/configs_${TAG}_strict.jsonl 2>> gpurun_out/configs_${TAG}.err
python tools/ct_timing.py > gpurun_out/ct_timing_masked.jsonl 2>/dev/null
FQ_STRICT_SELECT=1 python tools/ct_timing.py > gpurun_out/ct_timing_strict.jsonl 2>/dev/null
python tests/checks/allkernels_check.py
