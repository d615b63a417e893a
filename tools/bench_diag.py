"""Where does bench.py's `value` loop lose time against the kernel sum?  Times the same step with profiling on/off,
L2 flush on/off, per-step event pairs vs one pair around K steps."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from imgcomp_cvpr_b200 import _lib, weights

L = _lib.lib()
ae_name, N, H, Wd = bench.WORKLOADS['kodak24']
a, p, W, ae, pc = bench.make_models(ae_name, 'exact')
x = torch.from_numpy(weights.synthetic_images(N, H, Wd, seed=1234)).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')


def step():
    enc = ae.encode(x, is_training=False)
    pc.bitcost(enc.qbar, enc.symbols, is_training=False, pad_value=pc.auto_pad_value(ae))


for _ in range(3):
    step()
torch.cuda.synchronize()
K = 10
for prof in (0, 1):
    for fl in (0, 1):
        L.ic_profile_reset(); L.ic_profile_enable(prof)
        evs = []
        t0 = time.perf_counter()
        for _ in range(K):
            if fl:
                flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); step(); e1.record(); evs.append((e0, e1))
        t_host = (time.perf_counter() - t0) / K * 1e3
        torch.cuda.synchronize()
        t_wall = (time.perf_counter() - t0) / K * 1e3
        per = [a_.elapsed_time(b_) for a_, b_ in evs]
        tot = evs[0][0].elapsed_time(evs[-1][1]) / K
        L.ic_profile_enable(0)
        print('prof %d flush %d: host enqueue %.2f ms/step, wall %.2f, per-step events mean %.2f (min %.2f max %.2f), first->last/K %.2f'
              % (prof, fl, t_host, t_wall, sum(per) / K, min(per), max(per), tot), flush=True)
# host cost of the two calls alone
torch.cuda.synchronize()
t0 = time.perf_counter()
enc = ae.encode(x, is_training=False)
t1 = time.perf_counter()
pc.bitcost(enc.qbar, enc.symbols, is_training=False, pad_value=pc.auto_pad_value(ae))
t2 = time.perf_counter()
torch.cuda.synchronize()
print('host: encode call %.2f ms, bitcost call %.2f ms' % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
