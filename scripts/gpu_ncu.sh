#!/bin/bash
# ncu --set full captures of chosen kernels of one PGD iteration.  Usage: bash scripts/gpu_ncu.sh <tag> <kernel-regex> [count] [workload]
TAG=$1; K=$2; C=${3:-2}; WL=${4:-m128}
O=gpurun_out
mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:${K} -c ${C} \
    -f -o $O/${TAG}_full python scripts/one_step.py --workload $WL > $O/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i $O/${TAG}_full.ncu-rep --page raw --csv > $O/${TAG}_full_raw.csv 2>/dev/null
ls -la $O/${TAG}_full*
