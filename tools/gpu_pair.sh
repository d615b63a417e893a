export IC_CONV_PAIR=1
for cfg in "1 16 16 exact" "2 40 40 exact res" "3 48 24 fast res" "1 192 128 exact res"; do
  echo "== $cfg"; timeout -k 5 60 python tools/tc_debug.py $cfg 2>&1 | tail -8; echo "rc=$?"
done
timeout -k 5 300 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_hotpath.py -m gpu -q -x 2>&1 | tail -5
for m in exact fast; do
timeout -k 5 200 python bench.py --steps 5 --warmup 3 --mode $m --no-cpu-baseline > gpurun_out/bench_pair_$m.log 2>&1; tail -n1 gpurun_out/bench_pair_$m.log | cut -c1-1500
done
