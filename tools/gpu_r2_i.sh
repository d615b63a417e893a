# round 2, call I: pipelined h1 kernel, forward cache in the tensor-core training leg: full GPU suite + bench
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/r2i_pytest.log 2>&1; tail -n 12 gpurun_out/r2i_pytest.log | cut -c1-200
for g in 0 1; do
IC_H1_GENERIC=$g timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-parity > gpurun_out/r2i_bench_g$g.log 2>&1
tail -n1 gpurun_out/r2i_bench_g$g.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('H1_GENERIC=$g ms', d['ms_per_step'], 'value', d['value'], 'frac', d['roofline']['frac'], d['kernel_ms_per_step'], d['gpu_launches'])
"
done
timeout 300 python tools/train_time.py --steps 5 --cpu-batch 0 --graph 2>&1 | tail -1 | cut -c1-400
