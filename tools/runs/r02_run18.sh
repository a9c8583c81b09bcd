#!/bin/bash
mkdir -p gpurun_out
timeout 400 python tests/checks/engine_soak.py 90 > gpurun_out/r18_engine_soak.json 2> gpurun_out/r18_engine_soak.err; echo "soak rc=$?"; cat gpurun_out/r18_engine_soak.json; tail -3 gpurun_out/r18_engine_soak.err
