# round 2, call T16: last full check of the committed state
mkdir -p gpurun_out
timeout -k 5 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2t16_pytest.log 2>&1; tail -n 3 gpurun_out/r2t16_pytest.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 1
timeout 900 python bench.py > gpurun_out/r2t16_bench.log 2>&1
tail -n1 gpurun_out/r2t16_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['kernel_ms_per_step'], d['clocks'])
print('headline', d['headline']['value'], 'train', d['train_step']['ms_per_step'], 'real_bpp', d['real_bpp']['compress_ms_per_image'])"
