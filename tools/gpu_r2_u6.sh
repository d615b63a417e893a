# round 2, call U6: depth walk with one CTA per SM and deeper activation prefetch (IC_PC_ASLOTS=4 / 6 -> shared memory forces one CTA per SM) against the default (2 slots, two CTAs per SM)
mkdir -p gpurun_out
export IC_BENCH_ALLOW_SHORT=1
for sl in 2 4 6; do
IC_PC_ASLOTS=$sl timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-parity > gpurun_out/r2u6_bench_aslots$sl.log 2>&1
tail -n1 gpurun_out/r2u6_bench_aslots$sl.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('aslots=$sl ms', d['ms_per_step'], 'value', d['value'], d['kernel_ms_per_step'])"
done
