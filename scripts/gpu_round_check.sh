#!/bin/bash
# the GPU tests + the bench lines quoted in DESIGN.md / README.md at the current code
# usage: gpurun --timeout 1200 -- bash scripts/gpu_round_check.sh <tag>
TAG=${1:-rc}
mkdir -p gpurun_out
timeout 800 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/${TAG}_tests.log
tail -3 gpurun_out/${TAG}_tests.log
run() { name=$1; shift; timeout 400 python bench.py "$@" > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err; echo "$name rc=$?";
  python - gpurun_out/${TAG}_$name.json <<'PY' || tail -3 gpurun_out/${TAG}_$name.err
import json, sys
l = json.load(open(sys.argv[1])); r = l["roofline"]
print(round(l["value"] / 1e9, 3), "G", round(l["ms_per_step"], 4), "ms/step", l.get("repeats", {}).get("ms_per_step_median"), r["kernel_avg_ms"], "frac", round(r["frac"], 3), "step_frac", round(r["step_frac"], 3), l.get("parity"))
PY
}
run d20 --steps 20 --warmup 5
run d60 --steps 60 --warmup 5 --no-cpu-baseline --no-verify
run pol --steps 20 --warmup 5 --policies --no-cpu-baseline --no-verify
run shuf --steps 20 --warmup 5 --shuffle-agents --no-cpu-baseline --no-verify
run m9 --steps 60 --warmup 5 --agents 9000000 --no-cpu-baseline --no-verify
run ens8 --parallelism ensemble --agents 9000000 --window 30 --steps 30 --samples 8 --batch 8 --no-cpu-baseline --no-verify
