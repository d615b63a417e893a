export IC_BENCH_ALLOW_SHORT=1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 43 -c 3 -f -o gpurun_out/prof_conv3x3_exact python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 74 -c 3 -f -o gpurun_out/prof_pc python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/ncu_pc.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 43 -c 2 -f -o gpurun_out/prof_conv3x3_fast python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --mode fast > gpurun_out/ncu_fast.log 2>&1
ls -la gpurun_out | tail -12
