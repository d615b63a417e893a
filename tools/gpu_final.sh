python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout -k 5 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
for w in kodak24 b64_512 cfg4 cfg1; do
timeout -k 5 400 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --with-decode > gpurun_out/final_$w.log 2>&1
tail -n1 gpurun_out/final_$w.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'], round(d['value'],1), 'MPix/s', round(d['ms_per_step'],2),'ms; e2e', round(d['e2e']['value'],1), 'decode', round(d['decode']['MPix_per_s'],1), {k:round(v,2) for k,v in d['kernel_ms_per_step'].items()}, 'roof', round(d['roofline']['frac'],3), 'parity', d['parity']['symbol_mismatches'],'/',d['parity']['symbols'], d['parity']['max_abs_dbpp'], d['clocks'])"
done
timeout -k 5 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --mode fast > gpurun_out/final_fast.log 2>&1; tail -n1 gpurun_out/final_fast.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('fast', round(d['value'],1), round(d['ms_per_step'],2), round(d['roofline']['frac'],3), d['parity']['symbol_mismatches'])"
