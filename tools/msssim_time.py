"""Times both MS-SSIM variants on the Kodak batch (24 x 768 x 512) with CUDA events; IC_MSSSIM_TILED=1 selects the old
16x16-tile level kernel."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
from imgcomp_cvpr_b200 import ms_ssim, ms_ssim_np, weights

n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
x = torch.from_numpy(weights.synthetic_images(n, 768, 512, seed=3)).cuda()
y = (x.int() + torch.randint(-9, 10, x.shape, device='cuda')).clamp(0, 255).to(torch.uint8)
xf, yf = x.float(), y.float()


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for tiled in ('1', '0'):
    os.environ['IC_MSSSIM_TILED'] = tiled
    t_np = timed(lambda: ms_ssim_np.MultiScaleSSIM_batch(x, y, data_format='NCHW'))
    t_tf = timed(lambda: ms_ssim.MultiScaleSSIM(xf, yf, data_format='NCHW'))
    v = ms_ssim_np.MultiScaleSSIM_batch(x, y, data_format='NCHW').mean().item()
    t = ms_ssim.MultiScaleSSIM(xf, yf, data_format='NCHW').item()
    print('tiled=%s  np(double,u8) %.3f ms  tf(float) %.3f ms  values %.12f %.8f' % (tiled, t_np, t_tf, v, t))
