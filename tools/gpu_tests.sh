timeout -k 5 900 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1
grep -v "^$" gpurun_out/pytest_gpu.log | grep -v "bpp [0-9]" | tail -40
