# round-1 (second session) profile refresh: launch list of the default bench command, full capture of six
# consecutive residual 3x3 convs (traffic), launch list of one cfg-3 training step
export IC_BENCH_ALLOW_SHORT=1
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1b_launches_kodak24_exact.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 43 -c 6 -f -o gpurun_out/r1b_conv3x3 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/ncu_conv.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 1100 -c 420 --csv --log-file gpurun_out/r1b_launches_train_step.csv python tools/train_time.py --steps 1 --cpu-batch 0 > gpurun_out/ncu_train.log 2>&1
tail -n 2 gpurun_out/ncu_train.log | cut -c1-300
ls -la gpurun_out
