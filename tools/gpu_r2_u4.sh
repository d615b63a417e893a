# round 2, call U4: lean epilogue of the depth walk's plane-output layers: full GPU suite, smoke, issuer counters, default bench
mkdir -p gpurun_out
timeout -k 5 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2u4_pytest.log 2>&1; tail -n 3 gpurun_out/r2u4_pytest.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 1
IC_TC_DBG=2 timeout 300 python tools/hbm_kernels_once.py 24 2> gpurun_out/r2u4_dbg.txt | tail -n 1
grep "IC_TC_DBG" gpurun_out/r2u4_dbg.txt | grep "pair=0" | tail -n 3
timeout 900 python bench.py > gpurun_out/r2u4_bench.log 2>&1
tail -n1 gpurun_out/r2u4_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['kernel_ms_per_step'], d['clocks'])
print('headline', d['headline']['value'], 'train', d['train_step']['ms_per_step'], 'real_bpp', d['real_bpp']['compress_ms_per_image'], 'parity', d['parity']['exact']['symbol_mismatches'], d['parity']['fp32']['symbol_mismatches'])"
