#!/bin/sh
# builds the test-only CPU libraries:
#   libfq_hostsim.so     CPU instruction-level simulation of the device code (fourq_b200/csrc/*.cuh with -DFQ_HOSTSIM)
#   libfq_mockengine.so  the host engine of fourq_b200/csrc/capi.cu compiled against a mock CUDA runtime whose kernels are that
#                        simulation (tests/test_host_engine.py); `sh build.sh tsan|asan` adds the sanitizer build of engine_stress
set -e
cd "$(dirname "$0")"
CXX="g++ -O2 -std=c++17 -fPIC -Wall -Wno-unused-function -pthread"
$CXX -shared -DFQ_HOSTSIM -o libfq_hostsim.so hostsim.cpp
ENGINE_SRC="-x c++ ../../fourq_b200/csrc/capi.cu -x none mock_cuda_runtime.cpp mock_kernels.cpp hostsim.cpp"
$CXX -shared -DFQ_HOSTSIM -DFQ_MOCK_CUDA -I. -o libfq_mockengine.so $ENGINE_SRC
for san in "$@"; do
  case "$san" in
    tsan) $CXX -g -O1 -fsanitize=thread -DFQ_HOSTSIM -DFQ_MOCK_CUDA -I. -o engine_stress_tsan engine_stress.cpp $ENGINE_SRC ;;
    asan) $CXX -g -O1 -fsanitize=address,undefined -fno-sanitize-recover=undefined -DFQ_HOSTSIM -DFQ_MOCK_CUDA -I. -o engine_stress_asan engine_stress.cpp $ENGINE_SRC ;;
  esac
done
