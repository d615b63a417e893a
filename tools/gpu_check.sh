python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout -k 5 600 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1
for m in exact fast; do
timeout -k 5 300 python bench.py --steps 5 --warmup 3 --mode $m --no-cpu-baseline --with-decode > gpurun_out/bench_kodak_$m.log 2>&1
done
tail -3 gpurun_out/smoke.log; grep -v "^$" gpurun_out/pytest_gpu.log | grep -v "^\.\|bpp [0-9]" | tail -30; for m in exact fast; do tail -n1 gpurun_out/bench_kodak_$m.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['mode'], round(d['value'],1), 'MPix/s', round(d['ms_per_step'],2),'ms; e2e', round(d['e2e']['value'],1), 'decode', d['decode'], {k:round(v,2) for k,v in d['kernel_ms_per_step'].items()}, 'roof', round(d['roofline']['frac'],3), 'parity', d['parity']['symbol_mismatches'], d['parity']['max_abs_dbpp'])"; done
