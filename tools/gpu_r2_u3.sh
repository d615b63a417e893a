# round 2, call U3: ncu --set full (with source counters) of the context-model kernels with the depth walk (layers 1-2, head), Kodak batch
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_tc_kernel -s 35 -c 3 -f -o gpurun_out/r2u3_pc python tools/hbm_kernels_once.py 24 > gpurun_out/r2u3_ncu.log 2>&1; tail -1 gpurun_out/r2u3_ncu.log
ncu -i gpurun_out/r2u3_pc.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); hdr=rows[0]
for r in rows[2:]:
    d=dict(zip(hdr,r))
    print('%-70s %9s us  tensor %s  tc-smem %s  dram r/w %s / %s  sm%% %s issue %s' % (d['Kernel Name'][:70], d['gpu__time_duration.sum'], d['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'], d['l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed'], d['dram__bytes_read.sum'], d['dram__bytes_write.sum'], d['sm__throughput.avg.pct_of_peak_sustained_elapsed'], d['sm__issue_active.avg.pct_of_peak_sustained_elapsed']))
"
