#!/bin/bash
# everything the driver runs at round end, on one GPU: the GPU test-suite, smoke(), the default bench line
# usage: gpurun --timeout 1200 -- bash scripts/gpu_full_check.sh [tag]
TAG=${1:-full}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/${TAG}_tests.log
tail -6 gpurun_out/${TAG}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"
cat gpurun_out/${TAG}_bench.json | cut -c1-1500
cat gpurun_out/${TAG}_bench_ref.json | cut -c1-600
