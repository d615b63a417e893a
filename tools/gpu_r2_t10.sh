# round 2, call T10: context-model planes with 3 stored chunks (the zero padding chunk is neither written nor read): tests + bench
mkdir -p gpurun_out
export IC_BENCH_ALLOW_SHORT=1
timeout -k 5 900 python -m pytest tests/test_gpu_hotpath.py tests/test_gpu_decoder.py tests/test_gpu_full_size.py tests/test_gpu_codec.py tests/test_gpu_conv_tc.py -m gpu -q -x > gpurun_out/r2t10_pytest.log 2>&1; tail -n 3 gpurun_out/r2t10_pytest.log | cut -c1-200
for i in 1 2; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2t10_bench.log 2>&1
tail -n1 gpurun_out/r2t10_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
k=d['kernel_ms_per_step']
print('ms', d['ms_per_step'], 'value', d['value'], k, 'pc/conv3x3', k['probclass']/k['conv3x3'], 'parity', {m: (d['parity'][m]['symbol_mismatches'], d['parity'][m]['max_abs_dbpp']) for m in ('exact','fp32')})"
done
