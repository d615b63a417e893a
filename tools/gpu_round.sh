# full GPU check of the round: tests, smoke, default bench, cfg 3 training-step timing
mkdir -p gpurun_out
timeout -k 5 1200 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1
tail -n 3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_default.log 2>&1; tail -n1 gpurun_out/bench_default.log | cut -c1-600
timeout 300 python tools/train_time.py > gpurun_out/train_time.log 2>&1; tail -n1 gpurun_out/train_time.log
