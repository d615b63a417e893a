# round 2, call P: full GPU suite + launch list of the training step after the planned tensor-core convs
mkdir -p gpurun_out
export IC_BENCH_ALLOW_SHORT=1
timeout -k 5 1200 python -m pytest tests -m gpu -q -x --durations=3 > gpurun_out/r2p_pytest.log 2>&1; tail -n 7 gpurun_out/r2p_pytest.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2p_launches_train_step.csv python tools/train_time.py --cpu-batch 0 --steps 1 > gpurun_out/r2p_ncu_train.log 2>&1; tail -n 2 gpurun_out/r2p_ncu_train.log | cut -c1-200
