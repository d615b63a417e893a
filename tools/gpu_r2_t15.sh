# round 2, call T15: persistent + prefetching backward plane writer (merge kernel left per-group: the persistent variant measured slower, 12.5-13.2 -> 14.2 us)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training_step.py -m gpu -q -x -k "fused or oracle or graph or exact" 2>&1 | tail -n 2
for i in 1 2; do timeout 300 python tools/train_time.py --graph --cpu-batch 0 --steps 20 2>&1 | tail -n 1 | cut -c1-200; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'bn_bwd_apply_planes_tt' -s 40 -c 10 --csv --log-file gpurun_out/r2t15_tt.csv python tools/train_time.py --cpu-batch 0 --steps 1 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2t15_tt.csv')) if len(r)>5]
hdr=[r for r in rows if 'Kernel Name' in r][0]
ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[rows.index(hdr)+1:]: print(r[ki][15:60], r[vi])
PY
