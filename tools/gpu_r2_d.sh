# round 2, call D: B-concatenated kernel (single CTA, then CTA pair): parity tests, bench, ncu
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_hotpath.py tests/test_gpu_full_size.py -m gpu -q -x --durations=12 > gpurun_out/r2d_pytest.log 2>&1; tail -n 22 gpurun_out/r2d_pytest.log | cut -c1-200
grep -E "vs float64|mismatches vs" gpurun_out/r2d_pytest.log | head -20
for m in 1 2 0; do
IC_CONV_CAT=$m timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2d_bench_cat$m.log 2>&1
tail -n1 gpurun_out/r2d_bench_cat$m.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('CAT=$m ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'launch_ms', d['roofline']['avg_launch_ms'], d['kernel_ms_per_step'], d['clocks'])
print('   parity', d['parity']['exact'])
"
done
export IC_BENCH_ALLOW_SHORT=1
for m in 1 2; do
IC_CONV_CAT=$m timeout 300 ncu --set full --import-source on --clock-control none -k regex:conv_cat_kernel -s 12 -c 2 -f -o gpurun_out/r2d_cat$m python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-extras > gpurun_out/ncu_cat$m.log 2>&1
ncu -i gpurun_out/r2d_cat$m.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); hdr=rows[0]
for r in rows[2:]:
    d=dict(zip(hdr,r)); print('ncu CAT=$m', d['gpu__time_duration.sum'], d['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'], d['l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed'], d['dram__bytes_read.sum'], d['dram__bytes_write.sum'])
"
done
