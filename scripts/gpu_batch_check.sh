#!/bin/bash
# GPU check of the batched ensemble: parity tests, then the 9 M-agent ensemble window batched against one sample per
# replay.  usage: gpurun --timeout 900 -- bash scripts/gpu_batch_check.sh [tag] [configs...]
TAG=${1:-b1}
shift
CONFIGS=${@:-batch8 batch4 single lanes3}
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_batch.py -x -q -s > gpurun_out/${TAG}_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/${TAG}_tests.log
tail -12 gpurun_out/${TAG}_tests.log
C="--parallelism ensemble --agents 9000000 --window 30 --steps 30 --no-cpu-baseline --no-verify --repeats 2"
run() { name=$1; shift; timeout 240 python bench.py $C "$@" > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err; echo "$name rc=$?"; }
for c in $CONFIGS; do
  case $c in
    batch8) run batch8 --samples 8 --batch 8 ;;
    batch8own) GJ_BATCH_OWN_NOISE=1 run batch8own --samples 8 --batch 8 ;;
    batch4) run batch4 --samples 8 --batch 4 ;;
    batch2) run batch2 --samples 8 --batch 2 ;;
    single) run single --samples 6 --streams 1 ;;
    lanes3) run lanes3 --samples 6 --streams 3 ;;
  esac
done
for f in gpurun_out/${TAG}_*.json; do
  echo $f
  python - "$f" <<'PY' || tail -5 ${f%.json}.err
import json, sys
l = json.load(open(sys.argv[1]))
r = l["roofline"]
print(round(l["value"] / 1e9, 2), "G", round(l["ms_per_step"], 3), "ms/step", r["kernel_avg_ms"], r["kernel_frac"])
PY
done
