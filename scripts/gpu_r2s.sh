#!/bin/bash
# Short visit: parity subset (or all), then an environment A/B on the bench.  Usage: bash scripts/gpu_r2s.sh <tag> "<pytest -k expr or ALL>" "<VAR=val;...>"
TAG=$1; KEXPR=$2; VARIANTS=$3
O=gpurun_out; mkdir -p $O
if [ "$KEXPR" = "ALL" ]; then timeout 1200 python -m pytest tests -m gpu -q -x > $O/${TAG}_pytest.log 2>&1; else
timeout 900 python -m pytest tests -m gpu -q -x -k "$KEXPR" > $O/${TAG}_pytest.log 2>&1; fi
echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log; tail -4 $O/${TAG}_pytest.log
[ -n "$VARIANTS" ] && bash scripts/gpu_ab.sh $TAG "$VARIANTS" 100
