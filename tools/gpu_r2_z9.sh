# round 2, call Z9: ncu --set full of the backward kernels of the training step (filter gradient, backward statistics)
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'wgrad_tc_kernel|col_partial|col_finalize|wgrad_reduce' -s 150 -c 8 -f -o gpurun_out/r2z9_train_bwd python tools/train_time.py --cpu-batch 0 --steps 1 > gpurun_out/ncu_wgrad.log 2>&1
ncu -i gpurun_out/r2z9_train_bwd.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); hdr=rows[0]
for r in rows[2:]:
    d=dict(zip(hdr,r))
    print('%-50s %8s us tensor %6s tc-smem %6s lts %6s dram %6s l1 %6s sm %6s r/w %s/%s' % (d['Kernel Name'][:50], d['gpu__time_duration.sum'], d['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'][:6], d['l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed'][:6], d['lts__throughput.avg.pct_of_peak_sustained_elapsed'][:6], d['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'][:6], d['l1tex__throughput.avg.pct_of_peak_sustained_elapsed'][:6], d['sm__throughput.avg.pct_of_peak_sustained_elapsed'][:6], d['dram__bytes_read.sum'][:8], d['dram__bytes_write.sum'][:8]))
"
