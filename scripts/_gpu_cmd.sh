mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_fullsize.py -m gpu -q --durations=8 > gpurun_out/r01t_pytest.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/r01t_pytest.log
