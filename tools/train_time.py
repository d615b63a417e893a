"""cfg 3 (BASELINE.json configs[2]): time of one training step (forward + MS-SSIM / rate loss + backward + Adam) at
B=32, 160x160, cvpr/med + res_shallow on one B200, with the live share of each kernel class, next to the CPU oracle
(torch-CPU float32 autograd, all host cores) on a reduced batch.

    python tools/train_time.py [--batch 32] [--size 160] [--steps 5] [--cpu-batch 4]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from imgcomp_cvpr_b200 import _lib, config, trainer, weights


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=32)
    ap.add_argument('--size', type=int, default=160)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--cpu-batch', type=int, default=4)
    ap.add_argument('--ae', default='cvpr/med')
    ap.add_argument('--mode', default='exact', choices=['fp32', 'exact'])
    ap.add_argument('--graph', action='store_true', help='replay the step as one CUDA graph')
    args = ap.parse_args()
    a, p = config.ae_config(args.ae), config.pc_config('cvpr/res_shallow')
    W = weights.synthetic_weights(a.num_chan_bn, a.num_centers, p.arch_param__k, a.arch_param_B)
    x = torch.from_numpy(weights.synthetic_images(args.batch, args.size, args.size, seed=77)).cuda()
    tr = trainer.Trainer(a, p, W, num_itr_per_epoch=1000, mode=args.mode)
    if args.graph:
        tr.enable_cuda_graph(x)
    L = _lib.lib()
    for _ in range(2):
        out = tr.step(x)
    torch.cuda.synchronize()
    n0 = L.ic_launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = tr.step(x)
    ev[1].record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / args.steps * 1e3
    ms = ev[0].elapsed_time(ev[1]) / args.steps
    pix = args.batch * args.size * args.size
    res = {'workload': 'cfg3 training step', 'ae': args.ae, 'batch': args.batch, 'H': args.size, 'W': args.size,
           'ms_per_step': ms, 'cuda_graph': bool(args.graph), 'wall_ms_per_step': wall, 'images_per_s': args.batch / (ms * 1e-3), 'MPix_per_s': pix / (ms * 1e-3) / 1e6,
           'dtype': {'fp32': 'f32 (FFMA kernels)', 'exact': 'f32 + f16x3 tensor-core convs (fwd, dgrad, wgrad)'}[args.mode], 'loss': out['total_loss'], 'bpp': out['bpp'], 'ms_ssim': out['ms_ssim'],
           'counted_launches_per_step': (L.ic_launch_count() - n0) / args.steps,
           # forward FLOPs of SURVEY.md 8(d) x 3 (forward + data gradient + filter gradient)
           'algorithmic_tflop_per_step': 3 * 2 * (310562 + 309488 + 10470) * pix / 1e12}
    res['tflops'] = res['algorithmic_tflop_per_step'] / (ms * 1e-3)
    if args.cpu_batch > 0:
        from oracle import train_oracle as T
        torch.set_num_threads(os.cpu_count())
        xs = weights.synthetic_images(args.cpu_batch, args.size, args.size, seed=77)
        T.training_step(xs, W, a, p, dtype=torch.float32)
        t0 = time.perf_counter()
        T.training_step(xs, W, a, p, dtype=torch.float32)
        dt = time.perf_counter() - t0
        res['cpu_baseline'] = {'kind': 'port', 'cores': os.cpu_count(), 'sample': '%d of %d images, forward + backward (no Adam)' % (args.cpu_batch, args.batch),
                               'images_per_s': args.cpu_batch / dt}
    print(json.dumps(res))


if __name__ == '__main__':
    main()
