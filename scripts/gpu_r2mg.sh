#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): weak m128 and the strong-scaling BASELINE configurations 4 and 5 on N ranks.
# Usage: bash scripts/gpu_r2mg.sh <tag> <N> [steps]
TAG=$1; N=$2; STEPS=${3:-30}; O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm --format=csv,noheader > $O/${TAG}_gpu.txt 2>&1
port=29510
run() {  # name, extra args
  port=$((port+1))
  if [ "$N" = "1" ]; then
    timeout 400 python bench.py --gpus 1 --steps $STEPS --warmup 3 --no-cpu-baseline --no-cuda-baseline $2 > $O/${TAG}_bench_$1_n$N.json 2>> $O/${TAG}_bench.err
  else
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps $STEPS --warmup 3 --no-cpu-baseline --no-cuda-baseline $2 > $O/${TAG}_bench_$1_n$N.json 2>> $O/${TAG}_bench.err
  fi
  echo "$1 N=$N rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("$O/${TAG}_bench_$1_n$N.json").read().strip().splitlines()[-1])
    print("  ", d["config"]["workload"][:50], "| it/s %.2f | ms %.3f | e2e %.2f | graph %s | exact %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("cuda_graph"), d["config"].get("exact_global")))
except Exception as e:
    print("  no line:", e)
PY
}
run m128 "--workload m128"
run c4 "--workload c4"
run c5 "--workload c5"
[ "$N" = "1" ] || run c5x "--workload c5 --exact-global"
[ "$N" = "1" ] || run c4x "--workload c4 --exact-global"
tail -5 $O/${TAG}_bench.err
