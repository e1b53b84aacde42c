#!/bin/bash
# End-of-round check: the whole GPU suite, smoke(), and the timings of the 8(f) additions.  usage: tools/gpu_final_check.sh <tag> [bench]
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-rXX}
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -4 gpurun_out/${TAG}_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 300 python tools/gpu_time_new.py > gpurun_out/${TAG}_new_timings.json 2> gpurun_out/${TAG}_new_timings.err; cat gpurun_out/${TAG}_new_timings.json; tail -3 gpurun_out/${TAG}_new_timings.err
if [ "$2" = "bench" ]; then
    timeout 300 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; cut -c1-220 gpurun_out/${TAG}_bench.json
fi
