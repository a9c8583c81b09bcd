#!/bin/bash
# one B200, final build: 2^24-row soak of every scalar-multiplication path, X25519 soak, engine soak, GPU tests with the wipe switch
mkdir -p gpurun_out
timeout 900 python tests/checks/soak.py 24 > gpurun_out/r20_soak.json 2> gpurun_out/r20_soak.err; echo "soak rc=$?"; cat gpurun_out/r20_soak.json
timeout 600 python tests/checks/x25519_soak.py > gpurun_out/r20_x25519_soak.json 2> gpurun_out/r20_x25519_soak.err; echo "x25519 soak rc=$?"; cat gpurun_out/r20_x25519_soak.json
timeout 300 python tests/checks/engine_soak.py 60 > gpurun_out/r20_engine_soak.json 2> gpurun_out/r20_engine_soak.err; echo "engine soak rc=$?"; cat gpurun_out/r20_engine_soak.json
FQ_WIPE_AFTER_CALL=1 timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -1
