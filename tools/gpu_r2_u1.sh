# round 2, call U1: depth walk of the context model (ConvTcParams::walk): A/B test + existing parity tests, issuer counters, bench A/B
mkdir -p gpurun_out
export IC_BENCH_ALLOW_SHORT=1
timeout -k 5 240 python -m pytest tests/test_gpu_hotpath.py -m gpu -q -x -s -k "depth_walk or probclass_batched" > gpurun_out/r2u1_walk.log 2>&1
rc=$?; grep "depth walk\|passed\|failed\|Error\|error" gpurun_out/r2u1_walk.log | head -n 20 | cut -c1-220
if [ $rc -ne 0 ]; then echo "walk tests rc=$rc"; tail -n 30 gpurun_out/r2u1_walk.log | cut -c1-220; exit 0; fi
timeout -k 5 600 python -m pytest tests/test_gpu_hotpath.py tests/test_gpu_decoder.py tests/test_gpu_full_size.py tests/test_gpu_codec.py -m gpu -q -x > gpurun_out/r2u1_pytest.log 2>&1; tail -n 3 gpurun_out/r2u1_pytest.log | cut -c1-200
IC_TC_DBG=2 timeout 300 python tools/hbm_kernels_once.py 24 2> gpurun_out/r2u1_dbg.txt | tail -n 1
grep "IC_TC_DBG" gpurun_out/r2u1_dbg.txt | grep "pair=0" | tail -n 3
for wk in 1 0; do
IC_PC_WALK=$wk timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-parity > gpurun_out/r2u1_bench_walk$wk.log 2>&1
tail -n1 gpurun_out/r2u1_bench_walk$wk.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('walk=$wk ms', d['ms_per_step'], 'value', d['value'], d['kernel_ms_per_step'])"
done
