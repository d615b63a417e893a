# round 2, call T6: dedicated h1 kernel vs prep pass + grouped-tap kernel, A/B on one box
mkdir -p gpurun_out
for v in 1 0 1 0; do
IC_H1_GENERIC=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-parity > gpurun_out/r2t6_bench_$v.log 2>&1
tail -n1 gpurun_out/r2t6_bench_$v.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('h1_generic=$v ms', d['ms_per_step'], 'value', d['value'], d['kernel_ms_per_step'])"
done
