# round 2, call T9: to_bn forward on the tensor-core plan (48 / 80 columns, float32 output): tests + step time
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_ops.py tests/test_gpu_training_step.py tests/test_gpu_full_size.py -m gpu -q -x 2>&1 | tail -n 3
timeout 300 python tools/train_time.py --graph --cpu-batch 0 --steps 20 2>&1 | tail -n 1 | cut -c1-200 | tee gpurun_out/r2t9_train_time.txt
