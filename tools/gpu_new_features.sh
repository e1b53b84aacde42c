#!/bin/bash
# Round-1 late additions (Canny, order-statistic filters, metrics, insert blending): GPU parity tests + a few timings.  Output under gpurun_out/.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r01e}
timeout 900 python -m pytest tests/test_gpu_metrics.py tests/test_gpu_filters.py tests/test_gpu_compose.py tests/test_gpu_warp.py -m gpu -q \
    -k "metrics or order_statistic or canny or insert or extract" 2>&1 | tail -60 > gpurun_out/${TAG}_new_tests.log
tail -5 gpurun_out/${TAG}_new_tests.log
if [ "$2" = "time" ]; then
    timeout 300 python tools/gpu_time_new.py > gpurun_out/${TAG}_new_timings.json 2> gpurun_out/${TAG}_new_timings.err
    cat gpurun_out/${TAG}_new_timings.json
fi
