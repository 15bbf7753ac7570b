mkdir -p gpurun_out
timeout 600 python bench.py --steps 100 --no-cpu-baseline 2> gpurun_out/r02g_bench.err | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k:l[k] for k in ['value','ms_per_step','ms_per_step_eager','e2e','advk_ms_per_step']}); print(l['kernel_ms_per_step'])"
