timeout -k 5 600 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1
grep -v "^$" gpurun_out/pytest_gpu.log | tail -60
