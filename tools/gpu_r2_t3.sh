# round 2, call T3: final round check -- full suite, smoke, default bench line
mkdir -p gpurun_out
timeout -k 5 1200 python -m pytest tests -m gpu -q -x --durations=3 > gpurun_out/r2t3_pytest.log 2>&1; tail -n 6 gpurun_out/r2t3_pytest.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
/usr/bin/time -v timeout 900 python bench.py > gpurun_out/r2t3_bench.log 2> gpurun_out/r2t3_bench_time.txt
grep "Elapsed (wall clock)" gpurun_out/r2t3_bench_time.txt
tail -n1 gpurun_out/r2t3_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'launch_ms', d['roofline']['avg_launch_ms'])
print(d['kernel_ms_per_step']); print('clocks', d['clocks'])
print('parity', {k: (d['parity'][k]['symbol_mismatches'], d['parity'][k]['max_abs_dbpp']) for k in ('exact','fp32')})
h=d['headline']; print('headline', h['value'], h['ms_per_step'], h['roofline']['frac'], h['kernel_ms_per_step'])
print('train', d['train_step']['ms_per_step'], d['train_step']['roofline']['frac'], 'real_bpp', d['real_bpp']['compress_ms_per_image'], d['real_bpp']['tables_ms_per_image'], d['real_bpp']['decompress_ms_per_image'])
print('cpu', d['cpu_baseline'])
"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -n 1 | cut -c1-300
