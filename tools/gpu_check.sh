python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
python bench.py --workload cfg1 --steps 5 --warmup 3 > gpurun_out/bench_cfg1.log 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_kodak.log 2>&1
tail -5 gpurun_out/smoke.log; tail -60 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/bench_cfg1.log; tail -3 gpurun_out/bench_kodak.log
