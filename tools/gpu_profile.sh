export IC_BENCH_ALLOW_SHORT=1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 39 -c 3 -f -o gpurun_out/prof_conv3x3_exact python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log; ls -la gpurun_out
python bench.py --workload b64_512 --steps 5 --warmup 3 > gpurun_out/bench_b64_exact.log 2>&1; tail -n1 gpurun_out/bench_b64_exact.log
