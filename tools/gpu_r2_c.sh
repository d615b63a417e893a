# round 2, call C: warp-uniform MMA issue loops -- full GPU test-suite, bench (single + pair), ncu tensor-pipe check
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2c_pytest.log 2>&1; tail -n 4 gpurun_out/r2c_pytest.log
timeout 500 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c_bench.log 2>&1
tail -n1 gpurun_out/r2c_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'launch_ms', d['roofline']['avg_launch_ms'])
print(d['kernel_ms_per_step']); print('clocks', d['clocks'])
print('parity', {k: d['parity'][k] for k in ('exact','fp32')})
h=d['headline']; print('headline', h['value'], h['ms_per_step'], h['roofline']['frac'], h['kernel_ms_per_step'])
print('train', d['train_step']['ms_per_step'], 'real_bpp', d['real_bpp'])
"
IC_CONV_PAIR=1 timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-parity --no-extras > gpurun_out/r2c_bench_pair.log 2>&1
tail -n1 gpurun_out/r2c_bench_pair.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('pair', d['ms_per_step'], d['kernel_ms_per_step'], d['roofline']['avg_launch_ms'])"
export IC_BENCH_ALLOW_SHORT=1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:conv_tc_kernel -s 12 -c 2 -f -o gpurun_out/r2c_single python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-extras > gpurun_out/ncu_single.log 2>&1
ncu -i gpurun_out/r2c_single.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); hdr=rows[0]
for r in rows[2:]:
    d=dict(zip(hdr,r)); print(d['gpu__time_duration.sum'], d['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'], d['l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed'])
"
