# round 2, call H: dedicated h1 kernel + vectorised pc_conv0: full GPU suite + bench (new / generic h1)
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/r2h_pytest.log 2>&1; tail -n 12 gpurun_out/r2h_pytest.log | cut -c1-200
for g in 0 1; do
IC_H1_GENERIC=$g timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2h_bench_g$g.log 2>&1
tail -n1 gpurun_out/r2h_bench_g$g.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('H1_GENERIC=$g ms', d['ms_per_step'], 'value', d['value'], 'frac', d['roofline']['frac'], d['kernel_ms_per_step'], d['gpu_launches'])
print('   parity', d['parity']['exact'])
"
done
