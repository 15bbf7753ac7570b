#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, ncu launch list + full captures of the top kernels.
# Usage (through gpurun): bash scripts/gpu_round.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/${TAG}_build.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py --profile-out gpurun_out/${TAG}_bench_profile.json > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/${TAG}_bench.json
# FAST=1: skip the reference arm and the other workloads (short visits at the end of a GPU budget)
[ -n "$FAST" ] || { timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err; echo "ref rc=$?"; }
for wl in $([ -n "$FAST" ] || echo c2 c3 c4shard c5shard); do
  timeout 600 python bench.py --workload $wl --no-cpu-baseline --steps 50 > gpurun_out/${TAG}_bench_${wl}.json 2>> gpurun_out/${TAG}_bench.err; echo "bench $wl rc=$?"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/${TAG}_launches.csv python scripts/one_step.py > gpurun_out/${TAG}_ncu_list.log 2>&1; echo "ncu list rc=$?"
# one full capture per hot kernel family (2 launches each)
# (gpurun brings back at most 64 MiB: the big chain kernels are captured without their source listing)
for k in ss_step_bwd_lean_kernel ss_step_lean_kernel chain_bwd_kernel chain_fwd_kernel smooth3d_xy smooth3d_z loss_contour; do
  SRC="--import-source on"
  case $k in chain_*) SRC="";; esac
  timeout 600 ncu --set full --clock-control none $SRC --profile-from-start off -k regex:${k} -c 2 \
      -f -o gpurun_out/${TAG}_full_${k} python scripts/one_step.py > gpurun_out/${TAG}_ncu_full_${k}.log 2>&1; echo "ncu full ${k} rc=$?"
  case $k in chain_*)      # 30 MB reports (the SASS of the multi-stage kernels): keep the raw metric page only
    ncu -i gpurun_out/${TAG}_full_${k}.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_${k}.csv 2>/dev/null
    rm -f gpurun_out/${TAG}_full_${k}.ncu-rep;;
  esac
done
du -sh gpurun_out
ls -la gpurun_out | tail -30
