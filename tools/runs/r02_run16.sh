#!/bin/bash
# round 2, run 16: GFp25519 field ops (fq_fp25519_op) -- GPU parity, all-kernels pass, compare table, bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r16_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r16_pytest_gpu.log
python tests/checks/allkernels_check.py 2>&1 | tail -2
python tools/compare_ops.py --opmix profiles/r02_compare_opmix.json > gpurun_out/r16_compare.txt 2> gpurun_out/r16_compare.err; echo "compare rc=$?"; head -12 gpurun_out/r16_compare.txt
python bench.py --steps 20 --warmup 3 > gpurun_out/r16_bench_endo.json 2> gpurun_out/r16_bench_endo.err; echo "bench rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
