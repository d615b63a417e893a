# round 2, call B: unrolled MMA-shape ubench under sustained load + ncu of the pair kernel vs the single-CTA kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap --format=csv,noheader -lms 100 > gpurun_out/ub_clocks.csv & SMI=$!
timeout 200 tools/ubench/mma_shapes 64000 -1 60 > gpurun_out/r2_mma_shapes.txt 2>&1; kill $SMI; cat gpurun_out/r2_mma_shapes.txt
awk -F, '{print int($1/50)*50}' gpurun_out/ub_clocks.csv | sort -n | uniq -c | tr '\n' ';'; echo
export IC_BENCH_ALLOW_SHORT=1
IC_CONV_PAIR=1 timeout 300 ncu --set full --import-source on --clock-control none -k regex:conv_tc_kernel -s 12 -c 2 -f -o gpurun_out/r2_pair python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-extras > gpurun_out/ncu_pair.log 2>&1; tail -3 gpurun_out/ncu_pair.log | cut -c1-300
timeout 300 ncu --set full --import-source on --clock-control none -k regex:conv_tc_kernel -s 12 -c 1 -f -o gpurun_out/r2_single python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-extras > gpurun_out/ncu_single.log 2>&1; tail -3 gpurun_out/ncu_single.log | cut -c1-300
ls -la gpurun_out/*.ncu-rep
