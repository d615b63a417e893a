# round 2, call Q: vectorised BN kernels + mse/psnr seed sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_ops.py tests/test_gpu_training_step.py tests/test_gpu_full_size.py -m gpu -q -x -k "not mse_and_psnr and (train or bn or tc_plan or conv)" 2>&1 | tail -n 5 | tee gpurun_out/r2q_tests.log
timeout 300 python tools/train_time.py --graph --cpu-batch 0 --steps 10 2>&1 | tail -n 1 | cut -c1-300 | tee gpurun_out/r2q_train_time.txt
timeout 900 python tools/dist_seed_sweep.py 2>&1 | grep -v Warning | tee gpurun_out/r2q_dist_sweep.txt
