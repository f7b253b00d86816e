#!/bin/bash
# Build the bounds-instrumented library and run the GPU test-suite against it (run on a GPU box):
#   bash tools/debug_bounds.sh
set -e
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DAP_DEBUG_BOUNDS -shared -Xcompiler -fPIC \
  astropaint_b200/csrc/kernels.cu -o /tmp/libastropaint_b200_bounds.so
AP_LIB_PATH=/tmp/libastropaint_b200_bounds.so AP_REPORT_BOUNDS=1 python -m pytest tests -m gpu -q -x --timeout 900
