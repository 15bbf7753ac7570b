mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r01r_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r01r_pytest.log
timeout 300 python scripts/bench_morph.py m128 1.0 > gpurun_out/r01r_morph.log 2>&1; head -2 gpurun_out/r01r_morph.log
timeout 300 python scripts/bench_morph.py c2 1.0 > gpurun_out/r01r_morph_c2.log 2>&1; head -2 gpurun_out/r01r_morph_c2.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r01r_bench.json 2> gpurun_out/r01r_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
l=json.loads(open('gpurun_out/r01r_bench.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ['value','ms_per_step','ms_per_step_eager','e2e','gpu_launches','kernel_ms_per_step','advk_ms_per_step']})
PY
