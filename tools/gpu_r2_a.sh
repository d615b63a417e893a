# round 2, call A: MMA-shape ubench, baseline bench (single / pair), the new full-size parity tests + boundary tests
mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > gpurun_out/ub_clocks.csv & SMI=$!
timeout 150 tools/ubench/mma_shapes 200000 > gpurun_out/r2_mma_shapes.txt 2>&1; kill $SMI; cat gpurun_out/r2_mma_shapes.txt
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench0.log 2>&1; tail -n1 gpurun_out/r2_bench0.log | cut -c1-300
IC_CONV_PAIR=1 timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-parity --no-extras > gpurun_out/r2_bench0_pair.log 2>&1
tail -n1 gpurun_out/r2_bench0_pair.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['kernel_ms_per_step'], d['roofline']['avg_launch_ms'])"
timeout 600 python -m pytest tests/test_gpu_full_size.py tests/test_gpu_hotpath.py::test_training_mode_boundary_drops_in tests/test_gpu_training_step.py::test_cuda_graph_survives_growth_of_the_shared_workspace -m gpu -q -s > gpurun_out/r2_fullsize0.log 2>&1
grep -E "mismatch|bpp|ms-ssim|grad err|median|passed|failed|Error|assert" gpurun_out/r2_fullsize0.log | head -80
