# round 2, call W: context-model chunk pairing (LBO) + batch-norm backward writing the conv's planes: full suite, bench A/B, train A/B
mkdir -p gpurun_out
export IC_BENCH_ALLOW_SHORT=1
timeout -k 5 1200 python -m pytest tests -m gpu -q -x --durations=3 > gpurun_out/r2w_pytest.log 2>&1; tail -n 6 gpurun_out/r2w_pytest.log | cut -c1-200
for v in 0 1; do
IC_PC_PAIR_CHUNKS=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2w_bench_$v.log 2>&1
tail -n1 gpurun_out/r2w_bench_$v.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('pair_chunks=$v ms', d['ms_per_step'], 'value', d['value'], d['kernel_ms_per_step'], 'parity', {k: (d['parity'][k]['symbol_mismatches'], d['parity'][k]['max_abs_dbpp']) for k in ('exact','fp32')})"
done
for v in 0 1; do
IC_TRAIN_FUSED_BWD=$v timeout 300 python tools/train_time.py --graph --cpu-batch 0 --steps 10 2>&1 | tail -n 1 | cut -c1-200
done | tee gpurun_out/r2w_train_time.txt
