mkdir -p gpurun_out
timeout 900 python bench.py --steps 100 > gpurun_out/r01u_bench.json 2> gpurun_out/r01u_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r01u_bench.err; python - <<'PY'
import json
l=json.loads(open('gpurun_out/r01u_bench.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ['value','ms_per_step','ms_per_step_eager','e2e','gpu_launches','clocks','roofline','cpu_baseline']})
PY
