# round 2, call K: where does the MMA issuer wait?  single CTA vs CTA pair (IC_TC_DBG=1: prints per-launch wait shares)
mkdir -p gpurun_out
export IC_BENCH_ALLOW_SHORT=1
IC_TC_DBG=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras --no-parity 2> gpurun_out/r2k_dbg_single.txt | cut -c1-100
sort gpurun_out/r2k_dbg_single.txt | uniq -c | sort -rn | head -8
IC_CONV_PAIR=1 IC_TC_DBG=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras --no-parity 2> gpurun_out/r2k_dbg_pair.txt | cut -c1-100
sort gpurun_out/r2k_dbg_pair.txt | uniq -c | sort -rn | head -8
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-parity | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('default ms', d['ms_per_step'], 'value', d['value'], 'frac', d['roofline']['frac'], d['kernel_ms_per_step'])"
