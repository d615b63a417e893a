export IC_BENCH_ALLOW_SHORT=1
export IC_CONV_PAIR=1
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 40 -c 2 -f -o gpurun_out/prof_pair python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/ncu_pair.log 2>&1
tail -2 gpurun_out/ncu_pair.log | cut -c1-200
