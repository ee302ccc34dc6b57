#!/bin/bash
# DRAM bytes per launch of the big throughput kernels, batched (b = 8) against one sample per launch, on the 9 M-agent
# world: the evidence that a batched launch reads the world's index data and the profile once per b samples.
# usage: gpurun --timeout 900 -- bash scripts/ncu_batch_traffic.sh [tag]
TAG=${1:-nb}
mkdir -p gpurun_out
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct"
K='regex:k_pipe|k_lean_transmission|k_lean_group_sums|k_batch_noise'
C="--parallelism ensemble --agents 9000000 --window 3 --steps 3 --no-cpu-baseline --no-verify --repeats 0"
ncu --metrics $M --clock-control none -k "$K" -c 48 --csv --log-file gpurun_out/${TAG}_batch8.csv \
    python bench.py $C --samples 8 --batch 8 > gpurun_out/${TAG}_batch8.log 2>&1; echo "batch rc=$?"
ncu --metrics $M --clock-control none -k "$K" -c 48 --csv --log-file gpurun_out/${TAG}_single.csv \
    python bench.py $C --samples 2 --streams 1 > gpurun_out/${TAG}_single.log 2>&1; echo "single rc=$?"
ls -la gpurun_out/${TAG}_*
