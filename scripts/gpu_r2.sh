#!/bin/bash
# One GPU-box visit (round 2): parity tests, smoke, bench, launch list.  Usage: bash scripts/gpu_r2.sh <tag> [pytest -k expr]
TAG=${1:-r02a}
KEXPR=${2:-}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/${TAG}_gpu.txt 2>&1
nproc >> $O/${TAG}_gpu.txt
if [ -n "$KEXPR" ]; then
  timeout 1500 python -m pytest tests -m gpu -q -x -k "$KEXPR" --durations=15 > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log
else
  timeout 1500 python -m pytest tests -m gpu -q --durations=15 > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log
fi
tail -30 $O/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/${TAG}_smoke.log
timeout 600 python bench.py --no-cpu-baseline --steps 100 --profile-out $O/${TAG}_bench_profile.json > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"
tail -c 2500 $O/${TAG}_bench.json
[ -n "$SKIP_NCU" ] || { timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $O/${TAG}_launches.csv python scripts/one_step.py > $O/${TAG}_ncu_list.log 2>&1; echo "ncu list rc=$?"; }
