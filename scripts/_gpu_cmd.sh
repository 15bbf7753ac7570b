mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -v "^WARNING" | tail -6
timeout 600 python bench.py --steps 200 --no-cpu-baseline 2> gpurun_out/r02d_bench.err | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k:l[k] for k in ['value','ms_per_step','e2e','cuda_graph']})"
tail -2 gpurun_out/r02d_bench.err
