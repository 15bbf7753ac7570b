#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, ncu launch list + full capture of the top kernels.
# Usage (through gpurun): bash scripts/gpu_round.sh <tag> [kernel-regex]
TAG=${1:-r01}
KRE=${2:-ss_step}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/${TAG}_build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py --profile-out gpurun_out/${TAG}_bench_profile.json > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/${TAG}_bench.json
timeout 300 python bench.py --workload c2 --no-cpu-baseline > gpurun_out/${TAG}_bench_c2.json 2>> gpurun_out/${TAG}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/${TAG}_launches.csv python scripts/one_step.py > gpurun_out/${TAG}_ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:${KRE} -c 6 \
    -f -o gpurun_out/${TAG}_prof python scripts/one_step.py > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | tail -20
