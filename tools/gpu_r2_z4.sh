# round 2, call Z4: compile-time tap schedule for the streamed-weight kernels (3x3 convs, h1, transposed convs): full suite, A/B
mkdir -p gpurun_out
export IC_BENCH_ALLOW_SHORT=1
timeout -k 5 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2z4_pytest.log 2>&1; tail -n 3 gpurun_out/r2z4_pytest.log | cut -c1-200
for v in 0 1 0 1; do
IC_CONV_STATIC=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-parity > gpurun_out/r2z4_bench_$v.log 2>&1
tail -n1 gpurun_out/r2z4_bench_$v.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('static=$v ms', d['ms_per_step'], 'value', d['value'], d['kernel_ms_per_step'], 'launch', d['roofline']['avg_launch_ms'], d['clocks']['sm_mhz'])"
done
for v in 0 1; do
IC_CONV_STATIC=$v timeout 300 python tools/train_time.py --graph --cpu-batch 0 --steps 20 2>&1 | tail -n 1 | cut -c1-200
done
