#!/bin/bash
# Bench visit: default bench line (+ baselines), reference arm, other workloads.  Usage: bash scripts/gpu_bench.sh <tag> [extra workloads...]
TAG=$1; shift
O=gpurun_out
mkdir -p $O
timeout 900 python bench.py --steps 100 --warmup 5 --profile-out $O/${TAG}_bench_profile.json > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"
tail -c 1800 $O/${TAG}_bench.json
[ -n "$SKIP_REF" ] || { timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $O/${TAG}_bench_ref.json 2>> $O/${TAG}_bench.err; echo "ref rc=$?"; tail -c 900 $O/${TAG}_bench_ref.json; }
for wl in "$@"; do
  timeout 600 python bench.py --workload $wl --no-cpu-baseline --no-cuda-baseline --steps 50 > $O/${TAG}_bench_${wl}.json 2>> $O/${TAG}_bench.err; echo "bench $wl rc=$?"
done
tail -5 $O/${TAG}_bench.err
