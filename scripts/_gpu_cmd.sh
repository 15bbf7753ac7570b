mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_golden.py -m gpu -q -k "loss" > gpurun_out/r01z_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r01z_pytest.log
