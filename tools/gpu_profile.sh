export IC_BENCH_ALLOW_SHORT=1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 74 -c 3 -f -o gpurun_out/prof_pc python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/ncu_pc.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for i in 1 2; do python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -n1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), {k:round(v,2) for k,v in d['kernel_ms_per_step'].items()}, d['clocks'])"; done
