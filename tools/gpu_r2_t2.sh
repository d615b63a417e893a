# round 2, call T2: round check after the context-model / training work -- full suite, smoke, default bench, launch lists, pc ncu
mkdir -p gpurun_out
timeout -k 5 1200 python -m pytest tests -m gpu -q -x --durations=3 > gpurun_out/r2t2_pytest.log 2>&1; tail -n 6 gpurun_out/r2t2_pytest.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
timeout 900 python bench.py > gpurun_out/r2t2_bench.log 2>&1
tail -n1 gpurun_out/r2t2_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'launch_ms', d['roofline']['avg_launch_ms'])
print(d['kernel_ms_per_step']); print('clocks', d['clocks'])
print('parity', {k: (d['parity'][k]['symbol_mismatches'], d['parity'][k]['max_abs_dbpp']) for k in ('exact','fp32')})
h=d['headline']; print('headline', h['value'], h['ms_per_step'], h['roofline']['frac'], h['kernel_ms_per_step'])
print('train', d['train_step']['ms_per_step'], 'real_bpp', d['real_bpp']['compress_ms_per_image'], d['real_bpp']['tables_ms_per_image'])
"
export IC_BENCH_ALLOW_SHORT=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2t2_launches_kodak24_exact.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-extras > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2t2_launches_train_step.csv python tools/train_time.py --cpu-batch 0 --steps 1 > gpurun_out/r2t2_ncu_train.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:conv_tc_kernel -s 35 -c 3 -f -o gpurun_out/r2t2_pc python tools/hbm_kernels_once.py > gpurun_out/ncu_pc.log 2>&1
ncu -i gpurun_out/r2t2_pc.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); hdr=rows[0]
for r in rows[2:]:
    d=dict(zip(hdr,r))
    print('%-70s %9s us  tensor %s  tc-smem %s  dram r/w %s / %s  sm%% %s' % (d['Kernel Name'][:70], d['gpu__time_duration.sum'], d['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'], d['l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed'], d['dram__bytes_read.sum'], d['dram__bytes_write.sum'], d['sm__throughput.avg.pct_of_peak_sustained_elapsed']))
"
