#!/bin/bash
# A/B of environment variants on the bench's own timed regions (no profiler).  Usage: bash scripts/gpu_ab.sh <tag> "<VAR=val;...>" [steps]
TAG=$1; VARIANTS=$2; STEPS=${3:-100}
O=gpurun_out
mkdir -p $O
i=0
IFS=';' read -ra VS <<< "$VARIANTS"
for v in "${VS[@]}"; do
  [ "$v" = "base" ] && v=""
  env $v timeout 300 python bench.py --no-cpu-baseline --no-cuda-baseline --steps $STEPS > $O/${TAG}_ab_$i.json 2> $O/${TAG}_ab_$i.err
  python - <<PY
import json
d = json.load(open("$O/${TAG}_ab_$i.json"))
k = d["kernel_ms_per_step"]
print("   ", {n: k.get(n) for n in ("chain_pk_bwd", "chain_img_bwd", "chain_pk_fwd", "chain_img_fwd", "smooth_fwd", "smooth_bwd", "loss_contour", "loss_contour_adj", "loss_softmax", "loss_grad")})
print("[%s]" % "$v", "value %.1f it/s  %.4f ms | e2e %.1f | advk %.4f | regions %s | e2e regions %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["advk_ms_per_step"], d["ms_per_step_regions"], d["e2e"]["ms_per_step_regions"]))
PY
  i=$((i+1))
done
