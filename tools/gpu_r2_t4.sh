# round 2, call T4: trunk weights packed once per step; training tests, step time, default bench line (T3's bench did not run)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training_step.py tests/test_gpu_full_size.py tests/test_gpu_hotpath.py -m gpu -q -x -k "train or fused or graph or exact or boundary" 2>&1 | tail -n 3
timeout 300 python tools/train_time.py --graph --cpu-batch 0 --steps 20 2>&1 | tail -n 1 | cut -c1-200 | tee gpurun_out/r2t4_train_time.txt
SECONDS=0
timeout 900 python bench.py > gpurun_out/r2t4_bench.log 2>&1
echo "bench.py default run: $SECONDS s"
tail -n1 gpurun_out/r2t4_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'launch_ms', d['roofline']['avg_launch_ms'])
print(d['kernel_ms_per_step']); print('clocks', d['clocks'])
print('parity', {k: (d['parity'][k]['symbol_mismatches'], d['parity'][k]['max_abs_dbpp']) for k in ('exact','fp32')})
h=d['headline']; print('headline', h['value'], h['ms_per_step'], h['roofline']['frac'], h['kernel_ms_per_step'])
print('train', d['train_step']['ms_per_step'], d['train_step']['roofline']['frac'], 'real_bpp', d['real_bpp']['compress_ms_per_image'], d['real_bpp']['tables_ms_per_image'], d['real_bpp']['decompress_ms_per_image'])
print('cpu', d['cpu_baseline'])
"
