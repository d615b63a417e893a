# round 2, call V: ncu --set full of the context-model kernels (layers 1-2, head) at the Kodak batch + training timing with pair convs
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:conv_tc_kernel -s 35 -c 3 -f -o gpurun_out/r2v_pc python tools/hbm_kernels_once.py > gpurun_out/ncu_pc.log 2>&1; tail -1 gpurun_out/ncu_pc.log
ncu -i gpurun_out/r2v_pc.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); hdr=rows[0]
for r in rows[2:]:
    d=dict(zip(hdr,r))
    print('%-70s %9s us  tensor %s  tc-smem %s  dram r/w %s / %s  sm%% %s issue %s' % (d['Kernel Name'][:70], d['gpu__time_duration.sum'], d['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'], d['l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed'], d['dram__bytes_read.sum'], d['dram__bytes_write.sum'], d['sm__throughput.avg.pct_of_peak_sustained_elapsed'], d['sm__issue_active.avg.pct_of_peak_sustained_elapsed']))
"
timeout 600 python -m pytest tests/test_gpu_training_step.py -m gpu -q -x -k "conv3x3 or fused or exact" 2>&1 | tail -n 3
for v in 0 1; do
IC_CONV_PAIR=$v timeout 300 python tools/train_time.py --graph --cpu-batch 0 --steps 10 2>&1 | tail -n 1 | cut -c1-200
done | tee gpurun_out/r2v_train_time.txt
