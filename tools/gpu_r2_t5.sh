# round 2, call T5: set_params launch dropped; training tests + time; bench with decode
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training_step.py tests/test_gpu_full_size.py -m gpu -q -x -k "train or fused or graph or exact" 2>&1 | tail -n 2
timeout 300 python tools/train_time.py --graph --cpu-batch 0 --steps 20 2>&1 | tail -n 1 | cut -c1-200 | tee gpurun_out/r2t5_train_time.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-parity --with-decode > gpurun_out/r2t5_bench_decode.log 2>&1
tail -n1 gpurun_out/r2t5_bench_decode.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms', d['ms_per_step'], 'value', d['value'], 'decode', d['decode'])"
