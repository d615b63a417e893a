# round 2, call U: compute-sanitizer memcheck over inference + training (new tensor-core plans, fused trunk, MS-SSIM stream)
mkdir -p gpurun_out
IC_SANITIZE_GRAPH=0 timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_small.py > gpurun_out/r2u_sanitize_memcheck.log 2>&1; tail -n 12 gpurun_out/r2u_sanitize_memcheck.log
timeout 600 python -m pytest tests/test_gpu_conv_tc.py -m gpu -q -x --durations=2 2>&1 | tail -n 5
