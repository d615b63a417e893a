# round 2, call S: fused trunk layers (BN -> planes, conv -> BN statistics, one weight-scale launch)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training_step.py tests/test_gpu_full_size.py tests/test_gpu_train_ops.py -m gpu -q -x -k "train or fused or bn or distortion" -s 2>&1 | grep -i "fused vs\|passed\|failed\|error\|assert" | tail -n 12 | tee gpurun_out/r2s_tests.log
for v in 0 1; do
IC_TRAIN_FUSED=$v timeout 300 python tools/train_time.py --graph --cpu-batch 0 --steps 10 2>&1 | tail -n 1 | cut -c1-330
done | tee gpurun_out/r2s_train_time.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2s_launches_train_step.csv python tools/train_time.py --cpu-batch 0 --steps 1 > gpurun_out/r2s_ncu_train.log 2>&1; tail -n 1 gpurun_out/r2s_ncu_train.log | cut -c1-100
