#!/bin/bash
# A/B of an environment switch on the 56 M single-GPU line + the GPU tests that cover the throughput kernels
# usage: gpurun --timeout 900 -- bash scripts/gpu_ab_check.sh <tag> "<ENV=val for A>" "<ENV=val for B>" [pytest -k expr]
TAG=$1; A=$2; B=$3; K=${4:-}
mkdir -p gpurun_out
if [ -n "$K" ]; then
  timeout 600 python -m pytest tests -m gpu -x -q -k "$K" > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/${TAG}_tests.log
else
  timeout 800 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/${TAG}_tests.log
fi
tail -4 gpurun_out/${TAG}_tests.log
C="--steps 20 --warmup 5 --no-cpu-baseline --no-verify --repeats 3"
env $A timeout 300 python bench.py $C > gpurun_out/${TAG}_A.json 2> gpurun_out/${TAG}_A.err; echo "A rc=$?"
env $B timeout 300 python bench.py $C > gpurun_out/${TAG}_B.json 2> gpurun_out/${TAG}_B.err; echo "B rc=$?"
for f in gpurun_out/${TAG}_A.json gpurun_out/${TAG}_B.json; do
  echo $f
  python - "$f" <<'PY' || tail -5 ${f%.json}.err
import json, sys
l = json.load(open(sys.argv[1]))
r = l["roofline"]
print(round(l["value"] / 1e9, 3), "G", round(l["ms_per_step"], 4), "ms/step", l.get("repeats", {}).get("ms_per_step_median"), r["kernel_avg_ms"])
PY
done
