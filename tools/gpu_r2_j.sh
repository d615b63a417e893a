# round 2, call J: 8-warp h1 builder; robust cfg3 gate: full GPU suite + bench (new / generic h1)
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/r2j_pytest.log 2>&1; tail -n 9 gpurun_out/r2j_pytest.log | cut -c1-200
timeout 300 python -m pytest tests/test_gpu_full_size.py -m gpu -q -s -k cfg3 2>&1 | grep -E "cfg3|gradient error|passed|failed"
for g in 0 1; do
IC_H1_GENERIC=$g timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-parity > gpurun_out/r2j_bench_g$g.log 2>&1
tail -n1 gpurun_out/r2j_bench_g$g.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('H1_GENERIC=$g ms', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['kernel_ms_per_step'], d['gpu_launches'])
"
done
