#!/bin/bash
# Visit: parity suite under both PDL settings, A/B on the bench, 1-GPU runs of the strong-scaling workloads.
TAG=$1; O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/${TAG}_pytest.log 2>&1; echo "pytest (PDL on) rc=$?" | tee -a $O/${TAG}_pytest.log; tail -3 $O/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
bash scripts/gpu_ab.sh $TAG "base;ADVK_PDL=0;base;ADVK_PDL=0" 100
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "morph_field and 3" 2>&1 | tail -5
[ -n "$SKIP_STRONG" ] || bash scripts/gpu_r2mg.sh ${TAG} 1 20
