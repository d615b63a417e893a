# round 2, call M: CTA pairs by default + 8 epilogue warps: full GPU suite, issuer waits, bench pair / single
mkdir -p gpurun_out
export IC_BENCH_ALLOW_SHORT=1
timeout -k 5 900 python -m pytest tests -m gpu -q -x --durations=3 > gpurun_out/r2m_pytest.log 2>&1; tail -n 7 gpurun_out/r2m_pytest.log | cut -c1-200
for v in "IC_CONV_PAIR=1" "IC_CONV_PAIR=0"; do
env $v IC_TC_DBG=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras --no-parity 2> gpurun_out/r2m_dbg.txt | cut -c1-60
sort gpurun_out/r2m_dbg.txt | uniq -c | sort -rn | sed -n 2,4p
env $v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2m_bench.log 2>&1
tail -n1 gpurun_out/r2m_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$v ms', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'launch', d['roofline']['avg_launch_ms'], d['kernel_ms_per_step'], d['clocks'])"
done
