mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "bicubic" > gpurun_out/r01y_pytest.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/r01y_pytest.log
timeout 900 python bench.py --steps 200 --no-cpu-baseline > gpurun_out/r01y_bench.json 2> gpurun_out/r01y_bench.err; python - <<'PY'
import json
l=json.loads(open('gpurun_out/r01y_bench.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ['value','ms_per_step','e2e','clocks']})
PY
