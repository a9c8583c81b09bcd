#!/bin/bash
# developer helper run under gpurun: GPU parity tests, both bench algorithms, occupancy / select experiments
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
bash tools/quickbench.sh 2>&1 | tail -3
./tools/kexp/occ > gpurun_out/occ.log 2>&1; cat gpurun_out/occ.log
./tools/kexp/mainloop > gpurun_out/mainloop.log 2>&1; cat gpurun_out/mainloop.log
./tools/kexp/mainloop_strict > gpurun_out/mainloop_strict.log 2>&1; cat gpurun_out/mainloop_strict.log
