# round 2: 2-GPU sanity of the driver's scaling command (torchrun, NCCL) + gloo CPU tests are run in the container
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_2gpu.log 2>&1
tail -n1 gpurun_out/r2_bench_2gpu.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('n_gpus', d['n_gpus'], 'ms', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'scaling', d['scaling'], d['config'].get('parallelism'))"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>&1 | tail -n 1 | cut -c1-300
