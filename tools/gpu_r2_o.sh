# round 2, call O: planned tensor-core convs in the training step (h2 / h12 / h13 / context-model layers 1-3, fwd + dgrad)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_ops.py -m gpu -q -x -k "tc_plan" -s 2>&1 | grep -v "^$" | tail -n 40 > gpurun_out/r2o_plan_tests.log; tail -n 25 gpurun_out/r2o_plan_tests.log
timeout 900 python -m pytest tests/test_gpu_training_step.py tests/test_gpu_full_size.py -m gpu -q -x -k "train" 2>&1 | tail -n 8 | tee gpurun_out/r2o_train_tests.log
for v in 0 1; do
IC_TRAIN_TC_EXTRA=$v timeout 300 python tools/train_time.py --graph --cpu-batch 0 --steps 10 2>&1 | tail -n 1 | cut -c1-420
done | tee gpurun_out/r2o_train_time.txt
