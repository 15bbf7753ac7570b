#!/bin/bash
# GPU-box visit: parity suite + multi-iteration loops with / without the speculative step count.
TAG=${1:-r02z}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_golden.py tests/test_gpu_final_pass.py tests/test_gpu_fullsize.py -m gpu -q --durations=5 > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log
tail -5 $O/${TAG}_pytest.log
grep -B2 -A25 "^____" $O/${TAG}_pytest.log | head -80
for w in "m128 5" "c3 5" "c3 10" "c5shard 10"; do
  for s in 1 0; do
    ADVK_SPECULATE=$s timeout 300 python scripts/exp_multi.py $w 2>&1 | tail -2 | tee -a $O/${TAG}_multi.log
  done
done
