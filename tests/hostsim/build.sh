#!/bin/sh
# builds tests/hostsim/libfq_hostsim.so (test-only CPU simulation of the device code)
set -e
cd "$(dirname "$0")"
g++ -O2 -std=c++17 -fPIC -shared -DFQ_HOSTSIM -Wall -Wno-unused-function -o libfq_hostsim.so hostsim.cpp
