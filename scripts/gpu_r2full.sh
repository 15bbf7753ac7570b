#!/bin/bash
# One GPU-box visit (round 2, full evidence): parity suite, smoke, bench (all legs), launch list, ncu --set full of
# the hot kernel families.  Usage (through gpurun): bash scripts/gpu_r2full.sh <tag>
TAG=${1:-r02q}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/${TAG}_gpu.txt 2>&1
nproc >> $O/${TAG}_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q --durations=10 > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log
tail -5 $O/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/${TAG}_smoke.log
timeout 900 python bench.py --profile-out $O/${TAG}_bench_profile.json > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"
tail -c 1500 $O/${TAG}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $O/${TAG}_launches.csv python scripts/one_step.py > $O/${TAG}_ncu_list.log 2>&1; echo "ncu list rc=$?"
for w in c2 c3; do
  timeout 300 python bench.py --workload $w --steps 100 --no-cpu-baseline --no-cuda-baseline > $O/${TAG}_bench_$w.json 2> $O/${TAG}_bench_$w.err; echo "bench $w rc=$?"
done
for spec in "ss_step_bwd_tile_kernel:2" "ss_step_lean_kernel:2" "lean_warp:8" "smooth3d_tma:4" "loss_contour_fused:1" "lean_intensity:2"; do
  k=${spec%%:*}; c=${spec##*:}
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:${k} -c ${c} \
      -f -o $O/${TAG}_full_${k} python scripts/one_step.py > $O/${TAG}_ncu_full_${k}.log 2>&1; echo "ncu full ${k} rc=$?"
  ncu -i $O/${TAG}_full_${k}.ncu-rep --page raw --csv > $O/${TAG}_full_${k}.csv 2>/dev/null
  case $k in ss_step_bwd_tile_kernel) ;; *) rm -f $O/${TAG}_full_${k}.ncu-rep;; esac
done
du -sh $O
