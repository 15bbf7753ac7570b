mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_golden.py -m gpu -q -k "redoes" 2>&1 | grep -v "^WARNING" | tail -25
