# round 2, call E: context model with B-concatenated MMAs, u32 codec pipeline: full GPU suite + bench
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/r2e_pytest.log 2>&1; tail -n 16 gpurun_out/r2e_pytest.log | cut -c1-220
timeout 500 python bench.py --steps 10 --warmup 3 > gpurun_out/r2e_bench.log 2>&1
tail -n1 gpurun_out/r2e_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'launch_ms', d['roofline']['avg_launch_ms'])
print(d['kernel_ms_per_step']); print('clocks', d['clocks'])
print('parity', d['parity']['exact'])
h=d['headline']; print('headline', h['value'], h['ms_per_step'], h['roofline']['frac'], h['kernel_ms_per_step'])
print('train', d['train_step']['ms_per_step'], 'real_bpp', d['real_bpp'])
"
