# round 2, call U5: compute-sanitizer memcheck of the inference path with the depth walk; ncu --set full of the final context-model kernels
mkdir -p gpurun_out
IC_SANITIZE_INFER_ONLY=1 timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_small.py > gpurun_out/r2u5_sanitize_memcheck.log 2>&1; tail -n 6 gpurun_out/r2u5_sanitize_memcheck.log | cut -c1-200
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_tc_kernel -s 35 -c 3 -f -o gpurun_out/r2u5_pc python tools/hbm_kernels_once.py 24 > gpurun_out/r2u5_ncu.log 2>&1; tail -1 gpurun_out/r2u5_ncu.log
ncu -i gpurun_out/r2u5_pc.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); hdr=rows[0]
for r in rows[2:]:
    d=dict(zip(hdr,r))
    print('%-70s %9s us  tensor %s  tc-smem %s  dram r/w %s / %s  sm%% %s issue %s inst %s' % (d['Kernel Name'][:70], d['gpu__time_duration.sum'], d['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'], d['l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed'], d['dram__bytes_read.sum'], d['dram__bytes_write.sum'], d['sm__throughput.avg.pct_of_peak_sustained_elapsed'], d['smsp__issue_active.avg.pct_of_peak_sustained_active'], d['smsp__inst_executed.sum']))
"
