#!/bin/bash
# Short GPU-box visit: take the squaring-step adjoint apart (scripts/diag_ssb.py) + one ncu capture of the
# warp-box variant for the record.  Build the side libraries in the build container first
# (bash scripts/build_diag.sh -> scripts/_diag/, git-ignored, shipped by gpurun), then through gpurun:
#   bash scripts/gpu_exp_diag.sh <tag>
TAG=${1:-r01l}
O=gpurun_out
mkdir -p $O
for k in 0 1 2 3 4; do
  timeout 120 python scripts/diag_ssb.py $k m128 >> $O/${TAG}_diag_ssb.log 2>&1
done
cat $O/${TAG}_diag_ssb.log | grep "^diag"
ADVK_SSB_MODE=72 timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:ss_step_bwd_box -c 2 -f -o $O/${TAG}_full_ss_step_bwd_box python scripts/one_step.py > $O/${TAG}_ncu_full_box.log 2>&1; echo "ncu box rc=$?"
ADVK_SSB_MODE=24 timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:ss_step_bwd_box -c 2 -f -o $O/${TAG}_full_ss_step_bwd_box1 python scripts/one_step.py > $O/${TAG}_ncu_full_box1.log 2>&1; echo "ncu box1 rc=$?"
ls -la $O | tail -8
