#!/bin/bash
# One short GPU-box visit for the squaring-step kernel variants (advk_morph_tune masks): parity of every variant, kernel A/B,
# whole-step A/B (bench.py with the default and with the fastest variant), one ncu capture of the winner.
# Usage (through gpurun): bash scripts/gpu_exp_box.sh <tag>
TAG=${1:-r01k}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > $O/${TAG}_gpu.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k "test_morph_field" > $O/${TAG}_pytest_morph.log 2>&1; echo "pytest morph rc=$?" | tee -a $O/${TAG}_pytest_morph.log
tail -3 $O/${TAG}_pytest_morph.log
BENCH_MORPH_OUT=$O/${TAG}_morph_m128.json timeout 150 python scripts/bench_morph.py m128 1.0 4.0 > $O/${TAG}_morph_m128.log 2>&1; echo "bench_morph m128 rc=$?"
tail -2 $O/${TAG}_morph_m128.log
BENCH_MORPH_OUT=$O/${TAG}_morph_c2.json timeout 120 python scripts/bench_morph.py c2 1.0 4.0 > $O/${TAG}_morph_c2.log 2>&1; echo "bench_morph c2 rc=$?"
tail -1 $O/${TAG}_morph_c2.log
BEST=$(python -c "import json; print(json.load(open('$O/${TAG}_morph_m128.json'))['best'])" 2>/dev/null || echo 9)
echo "best mask (m128): $BEST"
timeout 240 python bench.py --no-cpu-baseline --steps 100 > $O/${TAG}_bench_default.json 2> $O/${TAG}_bench.err; echo "bench default rc=$?"
ADVK_SSB_MODE=$BEST timeout 240 python bench.py --no-cpu-baseline --steps 100 > $O/${TAG}_bench_best.json 2>> $O/${TAG}_bench.err; echo "bench best rc=$?"
python - <<PY
import json
for n in ("default", "best"):
    try:
        j = json.loads(open("$O/${TAG}_bench_%s.json" % n).read().strip().splitlines()[-1])
        print(n, "value", j["value"], "ms/step", j["ms_per_step"], "e2e", j["e2e"]["value"], "roofline", j["roofline"])
    except Exception as e:
        print(n, "unreadable:", e)
PY
[ -n "$SKIP_NCU" ] || ADVK_SSB_MODE=$BEST timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:ss_step_bwd -c 2 -f -o $O/${TAG}_full_ss_step_bwd python scripts/one_step.py > $O/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
[ -n "$SKIP_NCU" ] || ADVK_SSB_MODE=$BEST timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $O/${TAG}_launches.csv python scripts/one_step.py > $O/${TAG}_ncu_list.log 2>&1; echo "ncu list rc=$?"
ADVK_SSB_MODE=$BEST timeout 400 python -m pytest tests -m gpu -q -x > $O/${TAG}_pytest_best.log 2>&1; echo "pytest -m gpu (best mask as the default) rc=$?" | tee -a $O/${TAG}_pytest_best.log
tail -3 $O/${TAG}_pytest_best.log
ADVK_SSB_MODE=$BEST timeout 120 python bench.py --workload c2 --no-cpu-baseline --steps 50 > $O/${TAG}_bench_c2_best.json 2>> $O/${TAG}_bench.err; echo "bench c2 best rc=$?"
timeout 120 python bench.py --workload c2 --no-cpu-baseline --steps 50 > $O/${TAG}_bench_c2_default.json 2>> $O/${TAG}_bench.err; echo "bench c2 default rc=$?"
ls -la $O | tail -25
