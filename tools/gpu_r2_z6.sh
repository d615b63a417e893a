# round 2, call Z6: residual-gradient accumulation fused into the data-gradient conv's merge pass: tests + step time
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training_step.py tests/test_gpu_full_size.py -m gpu -q -x -k "train or fused or graph or exact" 2>&1 | tail -n 3
timeout 300 python tools/train_time.py --graph --cpu-batch 0 --steps 20 2>&1 | tail -n 1 | cut -c1-200 | tee gpurun_out/r2z6_train_time.txt
