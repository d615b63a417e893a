# round 2, call T7: backward batch-norm statistics accumulated in the data-gradient conv's merge pass: tests + A/B step time
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training_step.py tests/test_gpu_full_size.py tests/test_gpu_train_ops.py -m gpu -q -x 2>&1 | tail -n 3
for v in 0 1; do
IC_TRAIN_FUSED_BWD_STATS=$v timeout 300 python tools/train_time.py --graph --cpu-batch 0 --steps 20 2>&1 | tail -n 1 | cut -c1-200
done | tee gpurun_out/r2t7_train_time.txt
