python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu4.log
tail -15 gpurun_out/r2_pytest_gpu4.log
