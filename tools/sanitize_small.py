"""Small end-to-end run for compute-sanitizer (memcheck): inference (exact mode) + codec tables + one training step
(exact mode: tcgen05 forward / data / filter gradients) + CUDA-graph replay, on tiny shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from imgcomp_cvpr_b200 import autoencoder, config, probclass, trainer, weights

a, p = config.ae_config('cvpr/low'), config.pc_config('cvpr/res_shallow')
W = weights.synthetic_weights(a.num_chan_bn, a.num_centers, p.arch_param__k, a.arch_param_B)
x = torch.from_numpy(weights.synthetic_images(2, 48, 72, seed=3)).cuda()
ae = autoencoder.get_network_cls(a)(a, weights=W)
pc = probclass.get_network_cls(p)(p, num_centers=a.num_centers, weights=W)
enc = ae.encode(x, is_training=False)
bc = pc.bitcost(enc.qbar, enc.symbols, is_training=False, pad_value=pc.auto_pad_value(ae))
xo = ae.decode(enc.qhard, is_training=False)
f, _ = pc.freqs(enc.symbols, ae.get_centers_variable())
torch.cuda.synchronize()
print('inference ok', float(bc.sum()), float(xo.mean()), int(f.sum() > 0))
if os.environ.get('IC_SANITIZE_INFER_ONLY') == '1':      # the context model's depth walk on a second, ragged batch (segments of 1..3 outputs)
    x2 = torch.from_numpy(weights.synthetic_images(3, 40, 104, seed=5)).cuda()
    e2 = ae.encode(x2, is_training=False)
    bc2 = pc.bitcost(e2.qbar, e2.symbols, is_training=False, pad_value=pc.auto_pad_value(ae))
    torch.cuda.synchronize()
    print('second batch ok', float(bc2.sum()))
    sys.exit(0)
tr = trainer.Trainer(a, p, W, mode='exact')
xt = torch.from_numpy(weights.synthetic_images(2, 64, 64, seed=4)).cuda()
out = tr.step(xt)
torch.cuda.synchronize()
print('training step ok', out['total_loss'])
if os.environ.get('IC_SANITIZE_GRAPH', '1') == '1':
    tr.enable_cuda_graph(xt)
    out = tr.step(xt)
    torch.cuda.synchronize()
    print('graph step ok', out['total_loss'])
