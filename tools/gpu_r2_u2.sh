# round 2, call U2: residual prefetch in the depth walk's epilogue; MMA micro-benchmark with the A operand read out of a halo tile
mkdir -p gpurun_out
export IC_BENCH_ALLOW_SHORT=1
timeout 120 tools/ubench/mma_shapes 20000 -1 20 > gpurun_out/r2u2_mma_shapes.txt 2>&1; cat gpurun_out/r2u2_mma_shapes.txt | cut -c1-150
timeout -k 5 600 python -m pytest tests/test_gpu_hotpath.py tests/test_gpu_decoder.py tests/test_gpu_full_size.py tests/test_gpu_codec.py -m gpu -q -x > gpurun_out/r2u2_pytest.log 2>&1; tail -n 3 gpurun_out/r2u2_pytest.log | cut -c1-200
IC_TC_DBG=2 timeout 300 python tools/hbm_kernels_once.py 24 2> gpurun_out/r2u2_dbg.txt | tail -n 1
grep "IC_TC_DBG" gpurun_out/r2u2_dbg.txt | grep "pair=0" | tail -n 3
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-parity > gpurun_out/r2u2_bench.log 2>&1
tail -n1 gpurun_out/r2u2_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms', d['ms_per_step'], 'value', d['value'], d['kernel_ms_per_step'])"
