# round 2, call F: ncu of the HBM-bound kernels + launch list of the training step + launch list of the bench step
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_hotpath.py -m gpu -q -x > gpurun_out/r2f_pytest.log 2>&1; tail -n 3 gpurun_out/r2f_pytest.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'heatmap_quantize|prep_input_s2d|pc_conv0|ssim_level_kernel|downsample_kernel' -s 7 -c 14 -f -o gpurun_out/r2f_hbm python tools/hbm_kernels_once.py > gpurun_out/ncu_hbm.log 2>&1; tail -2 gpurun_out/ncu_hbm.log
ncu -i gpurun_out/r2f_hbm.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); hdr=rows[0]
for r in rows[2:]:
    d=dict(zip(hdr,r))
    t=float(d['gpu__time_duration.sum']); rd=float(d['dram__bytes_read.sum']); wr=float(d['dram__bytes_write.sum'])
    print('%-60s %9.1f us  read %8.1f %s write %8.1f %s  dram%% %s' % (d['Kernel Name'][:60], t, rd, '', wr, '', d.get('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','?')))
" 
export IC_BENCH_ALLOW_SHORT=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_train_step.csv python tools/train_time.py --steps 1 --cpu-batch 0 > gpurun_out/ncu_train.log 2>&1; tail -1 gpurun_out/ncu_train.log | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 --csv --log-file gpurun_out/r2_launches_kodak24_exact.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity --no-extras > gpurun_out/ncu_bench.log 2>&1; tail -1 gpurun_out/ncu_bench.log | cut -c1-200
