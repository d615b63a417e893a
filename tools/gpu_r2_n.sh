# round 2, call N: row-streaming MS-SSIM level kernel: tests, timing old/new, per-launch durations of both
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_hotpath.py -m gpu -q -x -k "msssim or golden" > gpurun_out/r2n_pytest.log 2>&1; tail -n 4 gpurun_out/r2n_pytest.log
timeout 300 python tools/msssim_time.py 2>&1 | tee gpurun_out/r2n_msssim_time.txt
for t in 1 0; do
IC_MSSSIM_TILED=$t timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'ssim_level|downsample' -s 18 -c 18 --csv --log-file gpurun_out/r2n_msssim_launches_tiled$t.csv python tools/hbm_kernels_once.py > /dev/null 2>&1
python - gpurun_out/r2n_msssim_launches_tiled$t.csv <<'PY'
import csv,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>5]
hdr=[r for r in rows if 'Kernel Name' in r][0]
ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); gi=hdr.index('Grid Size')
for r in rows[rows.index(hdr)+1:]:
    print('%-70s %-16s %10s ns'%(r[ki][15:85], r[gi], r[vi]))
PY
done
