# round 2, call X: final profiles -- launch lists (kodak24 step, training step), full capture of the context-model kernels
mkdir -p gpurun_out
export IC_BENCH_ALLOW_SHORT=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2x_launches_kodak24_exact.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-extras > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2x_launches_train_step.csv python tools/train_time.py --cpu-batch 0 --steps 1 > gpurun_out/r2x_ncu_train.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:conv_tc_kernel -s 35 -c 3 -f -o gpurun_out/r2x_pc python tools/hbm_kernels_once.py > gpurun_out/ncu_pc.log 2>&1; tail -1 gpurun_out/ncu_pc.log
ncu -i gpurun_out/r2x_pc.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); hdr=rows[0]
for r in rows[2:]:
    d=dict(zip(hdr,r))
    print('%-70s %9s us  tensor %s  tc-smem %s  dram r/w %s / %s  sm%% %s issue %s' % (d['Kernel Name'][:70], d['gpu__time_duration.sum'], d['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'], d['l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed'], d['dram__bytes_read.sum'], d['dram__bytes_write.sum'], d['sm__throughput.avg.pct_of_peak_sustained_elapsed'], d['sm__issue_active.avg.pct_of_peak_sustained_elapsed']))
"
timeout 300 python tools/train_time.py --graph --cpu-batch 0 --steps 20 2>&1 | tail -n 1 | cut -c1-200 | tee gpurun_out/r2x_train_time.txt
