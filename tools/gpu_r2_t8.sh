# round 2, call T8: final round check -- full suite, smoke, default bench, training launch list + step time
mkdir -p gpurun_out
timeout -k 5 1200 python -m pytest tests -m gpu -q -x --durations=3 > gpurun_out/r2t8_pytest.log 2>&1; tail -n 6 gpurun_out/r2t8_pytest.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
SECONDS=0
timeout 900 python bench.py > gpurun_out/r2t8_bench.log 2>&1
echo "bench.py default run: $SECONDS s"
tail -n1 gpurun_out/r2t8_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'launch_ms', d['roofline']['avg_launch_ms'])
print(d['kernel_ms_per_step']); print('clocks', d['clocks'])
print('parity', {k: (d['parity'][k]['symbol_mismatches'], d['parity'][k]['max_abs_dbpp']) for k in ('exact','fp32')})
h=d['headline']; print('headline', h['value'], h['ms_per_step'], h['roofline']['frac'])
print('train', d['train_step']['ms_per_step'], d['train_step']['roofline']['frac'], d['train_step']['gpu_launches_per_step'])
"
timeout 300 python tools/train_time.py --graph --cpu-batch 0 --steps 20 2>&1 | tail -n 1 | cut -c1-200 | tee gpurun_out/r2t8_train_time.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2t8_launches_train_step.csv python tools/train_time.py --cpu-batch 0 --steps 1 > gpurun_out/r2t8_ncu_train.log 2>&1; tail -n1 gpurun_out/r2t8_ncu_train.log | cut -c1-80
