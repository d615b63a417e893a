# round 2, call G: ncu of the remaining HBM-bound kernels (input prep, heatmap + quantizer, context-model layer 0) + new tests
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_train_ops.py tests/test_gpu_training_step.py -m gpu -q -x > gpurun_out/r2g_pytest.log 2>&1; tail -n 3 gpurun_out/r2g_pytest.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'heatmap_quantize|prep_input_s2d|pc_conv0' -s 3 -c 3 -f -o gpurun_out/r2g_hbm python tools/hbm_kernels_once.py > gpurun_out/ncu_hbm2.log 2>&1; tail -2 gpurun_out/ncu_hbm2.log
ncu -i gpurun_out/r2g_hbm.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); hdr=rows[0]
for r in rows[2:]:
    d=dict(zip(hdr,r))
    print('%-60s %9s us  read %10s write %10s  dram%% %s  regs %s' % (d['Kernel Name'][:60], d['gpu__time_duration.sum'], d['dram__bytes_read.sum'], d['dram__bytes_write.sum'], d.get('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','?'), d.get('launch__registers_per_thread','?')))
"
