"""--real_bpp timing on a Kodak-shaped image (configs[4]): one batched context-model pass + host range coder
(encode + teacher-forced decode + equality check), vs the reference README's ~350 s encode + ~200 s decode."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np, torch
from imgcomp_cvpr_b200 import autoencoder, bit_counter, config, probclass, weights

a, p = config.ae_config('cvpr/low'), config.pc_config('cvpr/res_shallow')
W = weights.synthetic_weights()
ae = autoencoder.get_network_cls(a)(a, weights=W, mode='exact')
pc = probclass.get_network_cls(p)(p, num_centers=6, weights=W)
x = torch.from_numpy(weights.synthetic_images(1, 512, 768, seed=3)).cuda()
enc = ae.encode(x, False)
pc.bitcost(enc.qbar, enc.symbols, False, pad_value=pc.auto_pad_value(ae))
pred = probclass.PredictionNetwork(pc, p, ae.get_centers_variable(), None)
sym = enc.symbols[0].cpu().numpy()
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    f, theory = pred.get_all_freqs(sym)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    nbits = bit_counter.encode_decode_to_file_ctx(sym, pred, syms_format='CHW')
    t2 = time.perf_counter()
print('symbols %d  coded bits %d  theoretical %.1f  bpp %.5f' % (sym.size, nbits, theory, nbits / (512 * 768)))
print('batched freqs pass (GPU + D2H of %d int64 tables): %.1f ms; full encode+decode+verify: %.1f ms' % (
    sym.size, (t1 - t0) * 1e3, (t2 - t1) * 1e3))
