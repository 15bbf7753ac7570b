mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_golden.py -m gpu -q > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r02b_pytest.log
