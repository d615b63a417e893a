# round 2, call U9 (last seconds of the budget): default bench line with the guarded sub-records, without the CPU legs
mkdir -p gpurun_out
IC_BENCH_ALLOW_SHORT=1 timeout 40 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2u9_bench.log 2> gpurun_out/r2u9_bench.err
tail -n1 gpurun_out/r2u9_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value', d['value'], 'headline', d.get('headline',{}).get('value', d.get('headline')), 'train', d.get('train_step',{}).get('ms_per_step', d.get('train_step')), 'real_bpp', d.get('real_bpp',{}).get('compress_ms_per_image', d.get('real_bpp')))"
tail -n 3 gpurun_out/r2u9_bench.err | cut -c1-200
