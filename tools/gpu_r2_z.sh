# round 2, call Z: where the context-model kernels' MMA issuer waits (IC_TC_DBG=2) with 2 and 6 activation-tile slots
mkdir -p gpurun_out
export IC_BENCH_ALLOW_SHORT=1
timeout 600 python -m pytest tests/test_gpu_training_step.py -m gpu -q -x -k "graph" 2>&1 | tail -n 3
for v in 2 6; do
IC_PC_ASLOTS=$v IC_TC_DBG=2 timeout 300 python tools/hbm_kernels_once.py 1 2> gpurun_out/r2z_dbg_$v.txt | tail -n 1
echo "== aslots $v (1 image)"; grep "pair=0" gpurun_out/r2z_dbg_$v.txt | grep -v "issuers=148 \|issuers=74 " | sort | uniq -c | sort -rn | head -8
IC_PC_ASLOTS=$v IC_TC_DBG=2 timeout 300 python tools/hbm_kernels_once.py 24 2> gpurun_out/r2z_dbg24_$v.txt | tail -n 1
echo "== aslots $v (24 images)"; grep "IC_TC_DBG" gpurun_out/r2z_dbg24_$v.txt | tail -n 12
IC_PC_ASLOTS=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-parity > gpurun_out/r2z_bench_$v.log 2>&1
tail -n1 gpurun_out/r2z_bench_$v.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('aslots=$v ms', d['ms_per_step'], 'value', d['value'], d['kernel_ms_per_step'])"
done
