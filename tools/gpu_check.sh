python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout -k 5 600 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1
for m in exact fast; do
timeout -k 5 300 python bench.py --steps 5 --warmup 3 --mode $m --no-cpu-baseline > gpurun_out/bench_kodak_$m.log 2>&1
done
tail -3 gpurun_out/smoke.log; grep -v "^$" gpurun_out/pytest_gpu.log | tail -50; tail -n1 gpurun_out/bench_kodak_*.log
