# round 2, call L: CTA pairs with the leader announcing both CTAs' bytes (no remote arrive per stage)
mkdir -p gpurun_out
export IC_BENCH_ALLOW_SHORT=1
IC_CONV_PAIR=1 timeout 600 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_hotpath.py -m gpu -q -x > gpurun_out/r2l_pytest_pair.log 2>&1; tail -n 3 gpurun_out/r2l_pytest_pair.log | cut -c1-200
IC_CONV_PAIR=1 IC_TC_DBG=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras --no-parity 2> gpurun_out/r2l_dbg_pair.txt | cut -c1-100
sort gpurun_out/r2l_dbg_pair.txt | uniq -c | sort -rn | head -4
for v in "IC_CONV_PAIR=1" "IC_CONV_CAT=2" "IC_CONV_PAIR=0"; do
env $v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2l_bench.log 2>&1
tail -n1 gpurun_out/r2l_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$v ms', d['ms_per_step'], 'value', d['value'], 'frac', d['roofline']['frac'], 'launch', d['roofline']['avg_launch_ms'], d['kernel_ms_per_step'], d['clocks'])
print('   ', d['parity']['exact'])"
done
