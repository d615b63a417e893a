# round 2, call Y: context-model kernels with 6 activation-tile slots (one CTA per SM): tests, bench, ncu
mkdir -p gpurun_out
export IC_BENCH_ALLOW_SHORT=1
timeout -k 5 900 python -m pytest tests/test_gpu_hotpath.py tests/test_gpu_decoder.py tests/test_gpu_full_size.py tests/test_gpu_train_ops.py tests/test_gpu_training_step.py -m gpu -q -x > gpurun_out/r2y_pytest.log 2>&1; tail -n 4 gpurun_out/r2y_pytest.log | cut -c1-200
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2y_bench.log 2>&1
tail -n1 gpurun_out/r2y_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms', d['ms_per_step'], 'value', d['value'], d['kernel_ms_per_step'], 'parity', {k: (d['parity'][k]['symbol_mismatches'], d['parity'][k]['max_abs_dbpp']) for k in ('exact','fp32')})"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:conv_tc_kernel -s 35 -c 3 -f -o gpurun_out/r2y_pc python tools/hbm_kernels_once.py > gpurun_out/ncu_pc.log 2>&1; tail -1 gpurun_out/ncu_pc.log
ncu -i gpurun_out/r2y_pc.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); hdr=rows[0]
for r in rows[2:]:
    d=dict(zip(hdr,r))
    print('%-70s %9s us  tensor %s  tc-smem %s  dram r/w %s / %s  sm%% %s lts %s' % (d['Kernel Name'][:70], d['gpu__time_duration.sum'], d['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'], d['l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed'], d['dram__bytes_read.sum'], d['dram__bytes_write.sum'], d['sm__throughput.avg.pct_of_peak_sustained_elapsed'], d['lts__throughput.avg.pct_of_peak_sustained_elapsed']))
"
timeout 300 python tools/train_time.py --graph --cpu-batch 0 --steps 20 2>&1 | tail -n 1 | cut -c1-200 | tee gpurun_out/r2y_train_time.txt
