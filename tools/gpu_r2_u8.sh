# round 2, call U8: depth walk with / without the paired third chunks (IC_PC_PAIR_CHUNKS): 14 k-steps per slice with 4 of them on overlapping-core-matrix descriptors, or 21 plain ones
mkdir -p gpurun_out
export IC_BENCH_ALLOW_SHORT=1
for pc in 1 0; do
IC_PC_PAIR_CHUNKS=$pc timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-parity > gpurun_out/r2u8_bench_pair$pc.log 2>&1
tail -n1 gpurun_out/r2u8_bench_pair$pc.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('pair_chunks=$pc ms', d['ms_per_step'], 'value', d['value'], d['kernel_ms_per_step'])"
done
