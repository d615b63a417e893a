export IC_BENCH_ALLOW_SHORT=1
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 71 -c 3 -f -o gpurun_out/prof_pc python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/ncu_pc.log 2>&1
tail -2 gpurun_out/ncu_pc.log | cut -c1-300
