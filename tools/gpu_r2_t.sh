# round 2, call T: the round check -- full GPU suite, smoke(), default bench line, ncu (launch list + full capture of the
# CTA-pair 3x3 conv), training-step timing after the col_partial change
mkdir -p gpurun_out
timeout -k 5 1200 python -m pytest tests -m gpu -q -x --durations=3 > gpurun_out/r2t_pytest.log 2>&1; tail -n 6 gpurun_out/r2t_pytest.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
timeout 900 python bench.py > gpurun_out/r2t_bench.log 2>&1
tail -n1 gpurun_out/r2t_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'launch_ms', d['roofline']['avg_launch_ms'])
print(d['kernel_ms_per_step']); print('clocks', d['clocks'])
print('parity', {k: (d['parity'][k]['symbol_mismatches'], d['parity'][k]['max_abs_dbpp']) for k in ('exact','fp32')})
h=d['headline']; print('headline', h['value'], h['ms_per_step'], h['roofline']['frac'], h['kernel_ms_per_step'])
print('train', d['train_step']['ms_per_step'], 'real_bpp', d['real_bpp'])
print('cpu_baseline', d['cpu_baseline'])
"
timeout 300 python tools/train_time.py --graph --cpu-batch 0 --steps 10 2>&1 | tail -n 1 | cut -c1-200 | tee gpurun_out/r2t_train_time.txt
export IC_BENCH_ALLOW_SHORT=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2t_launches_kodak24_exact.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-extras > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_tc_kernel -s 45 -c 6 -f -o gpurun_out/r2t_conv3x3_pair python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-extras > gpurun_out/ncu_conv.log 2>&1
ncu -i gpurun_out/r2t_conv3x3_pair.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); hdr=rows[0]
for r in rows[2:]:
    d=dict(zip(hdr,r)); print(d['Kernel Name'][:60], d['gpu__time_duration.sum'], d['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'], d['l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed'], d['dram__bytes_read.sum'], d['dram__bytes_write.sum'])
"
