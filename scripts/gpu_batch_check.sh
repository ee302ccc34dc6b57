#!/bin/bash
# GPU check of the batched ensemble: parity tests, then the 9 M-agent ensemble window batched (b = 8, 4) against one
# sample per replay (one lane / three lanes).  usage: gpurun --timeout 1500 -- bash scripts/gpu_batch_check.sh [tag]
TAG=${1:-b1}
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_batch.py -x -q -s > gpurun_out/${TAG}_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/${TAG}_tests.log
tail -30 gpurun_out/${TAG}_tests.log
C="--parallelism ensemble --agents 9000000 --window 30 --steps 30 --no-cpu-baseline --no-verify --repeats 2"
run() { name=$1; shift; timeout 240 python bench.py $C "$@" > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err; echo "$name rc=$?"; }
run batch8 --samples 8 --batch 8
run single --samples 6 --streams 1
run lanes3 --samples 6 --streams 3
for f in gpurun_out/${TAG}_*.json; do
  echo $f
  python - "$f" <<'PY' || tail -5 ${f%.json}.err
import json, sys
l = json.load(open(sys.argv[1]))
r = l["roofline"]
print(round(l["value"] / 1e9, 2), "G", round(l["ms_per_step"], 3), "ms/step", r["kernel_avg_ms"], r["kernel_frac"])
PY
done
