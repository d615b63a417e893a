# round 2, call Z2: compile-time issue schedule for the context-model kernels: tests, issuer counters, bench A/B
mkdir -p gpurun_out
export IC_BENCH_ALLOW_SHORT=1
timeout -k 5 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2z2_pytest.log 2>&1; tail -n 4 gpurun_out/r2z2_pytest.log | cut -c1-200
for v in 0 1; do
IC_PC_STATIC=$v IC_TC_DBG=2 timeout 300 python tools/hbm_kernels_once.py 24 2> gpurun_out/r2z2_dbg_$v.txt | tail -n 1
echo "== static $v"; grep "IC_TC_DBG" gpurun_out/r2z2_dbg_$v.txt | grep "pair=0" | tail -n 3
IC_PC_STATIC=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2z2_bench_$v.log 2>&1
tail -n1 gpurun_out/r2z2_bench_$v.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('static=$v ms', d['ms_per_step'], 'value', d['value'], d['kernel_ms_per_step'], 'parity', {k: (d['parity'][k]['symbol_mismatches'], d['parity'][k]['max_abs_dbpp']) for k in ('exact','fp32')})"
done
IC_PC_ASLOTS=4 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-parity > gpurun_out/r2z2_bench_a4.log 2>&1
tail -n1 gpurun_out/r2z2_bench_a4.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('aslots=4 ms', d['ms_per_step'], 'value', d['value'], d['kernel_ms_per_step'])"
timeout 300 python tools/train_time.py --graph --cpu-batch 0 --steps 20 2>&1 | tail -n 1 | cut -c1-200 | tee gpurun_out/r2z2_train_time.txt
