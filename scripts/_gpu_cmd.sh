mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -v "^WARNING" | tail -12
