for w in cfg1 cfg4 b64_512; do
timeout -k 5 400 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --with-decode > gpurun_out/bench_$w.log 2>&1; tail -n1 gpurun_out/bench_$w.log | cut -c1-2200
done
timeout -k 5 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --with-decode > gpurun_out/bench_kodak_dec.log 2>&1; tail -n1 gpurun_out/bench_kodak_dec.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('kodak24', d['value'], d['decode'], d['kernel_ms_per_step'])"
