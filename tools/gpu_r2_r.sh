# round 2, call R: tcgen05 filter gradients of h2 / h12 / context-model layers; launch list of the training step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_ops.py -m gpu -q -x -k "tc_plan" -s 2>&1 | grep -v "^$" | tail -n 40 > gpurun_out/r2r_plan_tests.log; grep -i "wgrad\|passed\|failed\|error" gpurun_out/r2r_plan_tests.log | tail -n 20
timeout 900 python -m pytest tests/test_gpu_training_step.py tests/test_gpu_full_size.py -m gpu -q -x -k "train or distortion" 2>&1 | tail -n 5 | tee gpurun_out/r2r_train_tests.log
timeout 300 python tools/train_time.py --graph --cpu-batch 0 --steps 10 2>&1 | tail -n 1 | cut -c1-300 | tee gpurun_out/r2r_train_time.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2r_launches_train_step.csv python tools/train_time.py --cpu-batch 0 --steps 1 > gpurun_out/r2r_ncu_train.log 2>&1; tail -n 1 gpurun_out/r2r_ncu_train.log | cut -c1-100
