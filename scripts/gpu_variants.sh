#!/bin/bash
# bench the 56 M line with each library under variants/ (built by scripts/variants.sh)
# usage: gpurun --timeout 900 -- bash scripts/gpu_variants.sh <tag> [bench args]
TAG=$1; shift
ARGS=${@:---steps 20 --warmup 5 --no-cpu-baseline --no-verify --repeats 3}
mkdir -p gpurun_out
for lib in variants/lib_*.so; do
  n=$(basename $lib .so); n=${n#lib_}
  GJ_LIB_PATH=$PWD/$lib timeout 300 python bench.py $ARGS > gpurun_out/${TAG}_$n.json 2> gpurun_out/${TAG}_$n.err; echo "$n rc=$?"
  python - gpurun_out/${TAG}_$n.json <<'PY' || tail -3 gpurun_out/${TAG}_$n.err
import json, sys
l = json.load(open(sys.argv[1])); r = l["roofline"]
print(round(l["value"] / 1e9, 3), "G", round(l["ms_per_step"], 4), "ms/step", l.get("repeats", {}).get("ms_per_step_median"), r["kernel_avg_ms"])
PY
done
