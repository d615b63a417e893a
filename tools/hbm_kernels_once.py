"""One pass over every HBM-bound kernel of the path at the Kodak batch shape (24 x 768 x 512, cvpr/low), for ncu:
input prep, heatmap + quantizer, context-model layer 0, both MS-SSIMs (level + downsample kernels)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
from imgcomp_cvpr_b200 import autoencoder, config, ms_ssim, ms_ssim_np, probclass, weights

n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
a, p = config.ae_config('cvpr/low'), config.pc_config('cvpr/res_shallow')
W = weights.synthetic_weights()
ae = autoencoder.get_network_cls(a)(a, weights=W, mode='exact')
pc = probclass.get_network_cls(p)(p, num_centers=6, weights=W)
x = torch.from_numpy(weights.synthetic_images(n, 768, 512, seed=3)).cuda()
for _ in range(2):
    enc = ae.encode(x, False)
    pc.bitcost(enc.qbar, enc.symbols, False, pad_value=pc.auto_pad_value(ae))
    x_out = ae.decode(enc.qhard, False)
    v = ms_ssim_np.MultiScaleSSIM_batch(x, ae.extra['x_out_u8'], data_format='NCHW')
    t = ms_ssim.MultiScaleSSIM(x.float(), x_out, data_format='NCHW')
torch.cuda.synchronize()
print('ok', float(v.mean()), float(t))
