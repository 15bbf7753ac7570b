#!/bin/bash
# 2-GPU visit: NCCL exact-global tests (eager + CUDA-graph loop), full GPU suite, weak and strong scaling lines.
TAG=$1
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm --format=csv,noheader > $O/${TAG}_gpu.txt 2>&1
timeout 900 python -m pytest tests/test_sharding_nccl.py -m gpu -q -x > $O/${TAG}_pytest_nccl.log 2>&1; echo "nccl pytest rc=$?" | tee -a $O/${TAG}_pytest_nccl.log
tail -15 $O/${TAG}_pytest_nccl.log
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_sharding_nccl.py > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log
tail -5 $O/${TAG}_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu-baseline --no-cuda-baseline > $O/${TAG}_bench_m128_n2.json 2> $O/${TAG}_bench_n2.err; echo "bench m128 n2 rc=$?"
timeout 600 $TR bench.py --gpus 2 --workload c5 --steps 10 --warmup 3 --no-cpu-baseline --no-cuda-baseline > $O/${TAG}_bench_c5_n2.json 2>> $O/${TAG}_bench_n2.err; echo "bench c5 n2 rc=$?"
timeout 600 $TR bench.py --gpus 2 --workload c5 --exact-global --steps 10 --warmup 3 --no-cpu-baseline --no-cuda-baseline > $O/${TAG}_bench_c5x_n2.json 2>> $O/${TAG}_bench_n2.err; echo "bench c5 exact n2 rc=$?"
timeout 600 $TR bench.py --gpus 2 --workload c4 --exact-global --steps 10 --warmup 3 --no-cpu-baseline --no-cuda-baseline > $O/${TAG}_bench_c4x_n2.json 2>> $O/${TAG}_bench_n2.err; echo "bench c4 exact n2 rc=$?"
for f in m128_n2 c5_n2 c5x_n2 c4x_n2; do python - <<PY
import json
try:
    d = json.load(open("$O/${TAG}_bench_$f.json"))
    print("$f", {k: d.get(k) for k in ("value", "ms_per_step", "n_gpus", "scaling", "cuda_graph")}, d.get("e2e", {}).get("value"), d["config"].get("exact_global"))
except Exception as e:
    print("$f", "unreadable", e)
PY
done
tail -5 $O/${TAG}_bench_n2.err
