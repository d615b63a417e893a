"""Codec timing on Kodak-shaped images (configs[4], --real_bpp): compress = autoencoder + codec tables
(batched) + host range coder (one thread per image); decompress = sequential on-device decode
(one CTA per image) + decoder network.  Reference README.md:65: ~350 s encode + ~200 s decode per image."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np, torch
from imgcomp_cvpr_b200 import autoencoder, codec, config, probclass, weights

n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
a, p = config.ae_config('cvpr/low'), config.pc_config('cvpr/res_shallow')
W = weights.synthetic_weights()
ae = autoencoder.get_network_cls(a)(a, weights=W, mode='exact')
pc = probclass.get_network_cls(p)(p, num_centers=6, weights=W)
imgs = [np.transpose(x, (1, 2, 0)) for x in weights.synthetic_images(n, 512, 768, seed=3)]
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    blobs = codec.compress(imgs, ae, pc, batch_size=n, threads=min(n, os.cpu_count() or 1))
    torch.cuda.synchronize(); t1 = time.perf_counter()
    rec = codec.decompress(blobs, ae, pc, batch_size=n)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    # the sequential kernel alone
    items = [codec.unpack(b) for b in blobs]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    pc.decode_streams([it['stream'] for it in items], [it['first_sym'] for it in items], (32, 64, 96), ae.centers_tensor())
    e1.record(); torch.cuda.synchronize()
bpp = np.mean([8 * len(b) / (512 * 768) for b in blobs])
print('%d images 768x512: %.4f bpp | compress %.1f ms (%.1f ms/image) | decompress %.1f ms (%.1f ms/image), '
      'of which sequential context-model decode %.1f ms for all images' % (
          n, bpp, (t1 - t0) * 1e3, (t1 - t0) * 1e3 / n, (t2 - t1) * 1e3, (t2 - t1) * 1e3 / n, e0.elapsed_time(e1)))
