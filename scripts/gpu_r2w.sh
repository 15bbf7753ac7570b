#!/bin/bash
# One GPU-box visit (round 2, late): parity suite on the early-verdict graph loop and the tile squaring-step
# kernels, per-kernel A/B of the field build under the tune masks, step-level A/B on the bench's timed regions.
# Usage (through gpurun): bash scripts/gpu_r2w.sh <tag>
TAG=${1:-r02w}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/${TAG}_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --durations=5 > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log
tail -5 $O/${TAG}_pytest.log
BENCH_MORPH_MASKS=${MASKS:-0,8,12,16,24} timeout 300 python scripts/bench_morph.py m128 1.0 4.0 > $O/${TAG}_bench_morph.log 2>&1; echo "bench_morph rc=$?"
grep -v "^vnorm" $O/${TAG}_bench_morph.log | tail -14
bash scripts/gpu_ab.sh $TAG "${VARIANTS:-base;ADVK_EARLY_VERDICT=0;ADVK_SSB_MODE=8;ADVK_SSB_MODE=16;ADVK_SSB_MODE=24}" 100 > $O/${TAG}_ab.log 2>&1; echo "ab rc=$?"
grep "^\[" $O/${TAG}_ab.log
