#!/bin/bash
# ncu evidence of one bench configuration (run on the GPU box through gpurun, one GPU):
#   1. the launch list (device time of every kernel of the library, cold-cache and serialised: compare SHARES);
#   2. one `--set full` capture of the throughput-mode agent kernels of ONE BPTT window.
# usage: scripts/ncu_profile.sh <tag> [bench args]          outputs -> gpurun_out/<tag>_*
set -u
tag=${1:-r1}
shift || true
args=${*:---steps 4 --window 4 --warmup 3 --no-cpu-baseline --no-verify --repeats 0 --no-graph}
mkdir -p gpurun_out
KERNELS='regex:k_lean|k_pipe|k_cell|k_dbeta|k_agent|k_tile|k_group'
python bench.py $args > gpurun_out/${tag}_plain.json 2> gpurun_out/${tag}_plain.err || { echo "plain run failed"; tail -5 gpurun_out/${tag}_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KERNELS" -c 400 --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py $args > gpurun_out/${tag}_ncu1.log 2>&1
# a window of 4 steps launches 28 throughput-mode agent / group kernels: 4 x [transmission (+ scatter), scatter
# finalize, forward] + 4 x [backward, group sums, fix, gather].  Skipped: the warm-up window (28) and the first forward
# step (3).  Captured (17, the report must stay below gpurun's 64 MiB): three forward steps, then the window's last
# step's backward (no incoming cotangents) and the backward of its third step — a TRUE middle step (all six state
# cotangents flow in and out).
ncu --set full --clock-control none --import-source on -k "regex:k_lean|k_pipe" -s 31 -c 17 -f \
    -o gpurun_out/${tag}_lean_full python bench.py $args > gpurun_out/${tag}_ncu2.log 2>&1
ls -la gpurun_out/${tag}_*
