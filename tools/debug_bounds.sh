#!/bin/bash
# Build the bounds-instrumented library and run the GPU test-suite against it (run on a GPU box):
#   bash tools/debug_bounds.sh
set -e
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build_library('/tmp/libastropaint_b200_bounds.so', defines=['-DAP_DEBUG_BOUNDS'], obj_dir='/tmp/ap_bounds_obj', force=True)"
AP_LIB_PATH=/tmp/libastropaint_b200_bounds.so AP_REPORT_BOUNDS=1 python -m pytest tests -m gpu -q -x --timeout 900
