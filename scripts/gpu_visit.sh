#!/bin/bash
# One GPU-box visit: parity tests (optionally a -k subset), smoke, launch lists of one PGD step under a list of
# environment variants, short bench.   Usage: bash scripts/gpu_visit.sh <tag> "<pytest -k expr or ''>" "<VAR=val;VAR=val ...>" [bench]
TAG=$1; KEXPR=$2; VARIANTS=$3; BENCH=$4
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > $O/${TAG}_gpu.txt 2>&1
if [ "$KEXPR" != "skip" ]; then
  if [ -n "$KEXPR" ]; then
    timeout 1500 python -m pytest tests -m gpu -q -x -k "$KEXPR" > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log
  else
    timeout 1500 python -m pytest tests -m gpu -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log
  fi
  tail -12 $O/${TAG}_pytest.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/${TAG}_smoke.log
fi
i=0
IFS=';' read -ra VS <<< "$VARIANTS"
for v in "${VS[@]}"; do
  [ "$v" = "base" ] && v=""
  env $v timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
      --log-file $O/${TAG}_launches_$i.csv python scripts/one_step.py > $O/${TAG}_ncu_list_$i.log 2>&1; echo "ncu list [$v] rc=$?"
  echo "== variant $i: [$v]"; python scripts/launch_table.py $O/${TAG}_launches_$i.csv 60 | grep -v "at::\|cudnn\|sm80_\|convolve"
  i=$((i+1))
done
if [ -n "$BENCH" ]; then
  timeout 600 python bench.py --no-cpu-baseline --no-cuda-baseline --steps 100 --profile-out $O/${TAG}_bench_profile.json > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"
  python -c "
import json; d=json.load(open('$O/${TAG}_bench.json')); print({k: d[k] for k in ('value','ms_per_step','advk_ms_per_step')}, d['e2e']['value']); print(d['kernel_ms_per_step'])"
fi
