# round 2, call T12: ncu --set full of the tile-transposing planes <-> NHWC kernels of the training step
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'bn_apply_planes_tt|bn_bwd_apply_planes_tt|merge_tt' -s 300 -c 8 -f -o gpurun_out/r2t12_tt python tools/train_time.py --cpu-batch 0 --steps 1 > gpurun_out/ncu_tt.log 2>&1
ncu -i gpurun_out/r2t12_tt.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); hdr=rows[0]
for r in rows[2:]:
    d=dict(zip(hdr,r))
    print('%-46s %7s us lts %5s dram %5s l1 %5s r/w %s/%s MB' % (d['Kernel Name'][:46], d['gpu__time_duration.sum'][:7], d['lts__throughput.avg.pct_of_peak_sustained_elapsed'][:5], d['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'][:5], d['l1tex__throughput.avg.pct_of_peak_sustained_elapsed'][:5], d['dram__bytes_read.sum'][:6], d['dram__bytes_write.sum'][:6]))
"
