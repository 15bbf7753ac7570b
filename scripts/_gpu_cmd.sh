mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_golden.py -m gpu -q -k "morph or fixture" 2>&1 | grep -v "^WARNING" | tail -3
timeout 600 python bench.py --steps 100 --no-cpu-baseline 2> gpurun_out/r02f_bench.err | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k:l[k] for k in ['value','ms_per_step','e2e']}); print({k:v for k,v in l['kernel_ms_per_step'].items() if k in ('init_phi0','unorm2')})"
