"""val.py-style driver over in-memory images (code/val.py:40-212 without the
checkpoint / TensorBoard / CSV plumbing): per image bpp, MS-SSIM (float64 numpy
variant), PSNR; optional --real_bpp; images sharded over ranks with ONE
all-reduce of the metric sums at the end.

    python -m torch.distributed.run --nproc-per-node 8 -m imgcomp_cvpr_b200.val --synthetic 64 --height 512 --width 768
"""
import numpy as np

METRICS = ('bpp', 'ms-ssim', 'psnr')


def add_padding(im_hwc, pad):
    """images_iterator.CachedImageLoader.add_padding (code/images_iterator.py:39-59): centred zero
    padding of H and W to a multiple of `pad`; alpha channel dropped.  -> (padded, undo)."""
    h, w, chan = im_hwc.shape
    if chan == 4:
        return add_padding(im_hwc[:, :, :3], pad)
    if h % pad == 0 and w % pad == 0:
        return im_hwc, (lambda x: x)
    hp, wp = (pad - h % pad) % pad, (pad - w % pad) % pad
    ht, wl = hp // 2, wp // 2
    hb, wr = hp - ht, wp - wl
    out = np.pad(im_hwc, [[ht, hb], [wl, wr], [0, 0]], mode='constant')
    return out, (lambda x: x[ht:(-hb or None), wl:(-wr or None), :])


def shard_range(n, rank, world):
    """contiguous shard [lo, hi) of n items for `rank`; sizes differ by at most 1."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def reduce_metric_sums(local_rows, dist=None):
    """local_rows: list of dicts with METRICS -> averages over ALL ranks' images (val.py:240-252,
    ValuesAggregator) using one all-reduce(sum) of [sum_bpp, sum_ms_ssim, sum_psnr, count]."""
    import torch
    v = torch.zeros(len(METRICS) + 1, dtype=torch.float64)
    for row in local_rows:
        for i, k in enumerate(METRICS):
            assert not np.isnan(row[k]), 'nan encountered in {}'.format(row)       # val.py:248
            v[i] += float(row[k])
        v[-1] += 1
    if dist is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dev = 'cuda' if dist.get_backend() == 'nccl' else 'cpu'
        v = v.to(dev)
        dist.all_reduce(v)
        v = v.cpu()
    n = max(float(v[-1]), 1.0)
    return {k: float(v[i]) / n for i, k in enumerate(METRICS)}, int(v[-1])


def psnr_u8(a, b):
    """val.psnr_np -> skimage compare_psnr on uint8 (code/val.py:227-237), per image, on device."""
    import torch
    mse = ((a.double() - b.double()) ** 2).flatten(1).mean(1)
    return 10 * torch.log10(255.0 ** 2 / mse)


def measure_batch(x_u8, ae, pc, real_bpp=False, keep_images=False):
    """The val.py fetch for a batch of same-sized uint8 images (code/val.py:81-94,161-175).
    -> list of per-image dicts (bpp, ms-ssim, psnr[, bpp_real, bpp_theory][, img_out: the padded uint8 CHW output])."""
    import torch
    from . import bpp_helpers, ms_ssim_np, probclass
    enc = ae.encode(x_u8, is_training=False)
    x_out = ae.decode(enc.qhard, is_training=False)
    pc.bitcost(enc.qbar, enc.symbols, is_training=False, pad_value=pc.auto_pad_value(ae))
    num_pixels = x_u8.shape[2] * x_u8.shape[3]
    bpp = (pc.last_bits_per_image / num_pixels).float().cpu().numpy()         # bits.bitcost_to_bpp per image
    x_out_u8 = ae.extra['x_out_u8']
    ms = ms_ssim_np.MultiScaleSSIM_batch(x_u8, x_out_u8, data_format='NCHW').float().cpu().numpy()
    ps = psnr_u8(x_u8, x_out_u8).float().cpu().numpy()
    rows = [{'bpp': float(bpp[i]), 'ms-ssim': float(ms[i]), 'psnr': float(ps[i])} for i in range(x_u8.shape[0])]
    if keep_images:                                                     # fetch_dict['img_out'] (val.py:108-109)
        out_host = x_out_u8.cpu().numpy()
        for i, row in enumerate(rows):
            row['img_out'] = out_host[i]
    if real_bpp:
        pred = probclass.PredictionNetwork(pc, pc.config, ae.get_centers_variable(), None)
        checker = probclass.ProbclassNetworkTesting(pc, ae, None)
        fetcher = bpp_helpers.BppFetcher(pred, checker)
        sym = enc.symbols.cpu().numpy()
        for i, row in enumerate(rows):
            bpp_real, bpp_theory = fetcher.get_bpp(sym[i:i + 1], num_pixels)
            assert abs(bpp_theory - row['bpp']) < 1e-3, 'Expected bpp_theory to match loss! Got {} and {}'.format(
                bpp_theory, row['bpp'])                                              # val.py:174
            row['bpp_real'], row['bpp_theory'] = float(bpp_real), float(bpp_theory)
    return rows


def validate(images_u8, ae, pc, real_bpp=False, batch_size=8, dist=None, keep_images=False):
    """images_u8: list of HWC or CHW uint8 arrays (any sizes).  Pads like the reference's
    ImagesIterator, shards over ranks, batches same-sized images.  -> (averages, n_images, local rows)."""
    import torch
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist is not None and dist.is_initialized() else (0, 1)
    lo, hi = shard_range(len(images_u8), rank, world)
    f = ae.get_subsampling_factor()
    padded = []
    for im in images_u8[lo:hi]:
        im = np.asarray(im)
        hwc = im if im.shape[-1] in (3, 4) else np.transpose(im, (1, 2, 0))
        p, _ = add_padding(hwc, f)
        padded.append(np.ascontiguousarray(np.transpose(p, (2, 0, 1))))
    rows = [None] * len(padded)
    by_shape = {}
    for i, p in enumerate(padded):
        by_shape.setdefault(p.shape, []).append(i)
    for shape, idxs in by_shape.items():
        for s in range(0, len(idxs), batch_size):
            chunk = idxs[s:s + batch_size]
            x = torch.from_numpy(np.stack([padded[i] for i in chunk])).cuda()
            for i, row in zip(chunk, measure_batch(x, ae, pc, real_bpp, keep_images)):
                rows[i] = row
    avgs, n = reduce_metric_sums(rows, dist)
    return avgs, n, rows


class MeasuresWriter(object):
    """val_files.MeasuresWriter (code/val_files.py:62-77): out_dir/measures.csv, one row per image."""

    def __init__(self, out_dir):
        import os
        os.makedirs(out_dir, exist_ok=True)
        self.fout = open(os.path.join(out_dir, 'measures.csv'), 'w')
        self.fout.write('img_name,bpp,ms-ssim,psnr\n')

    def append(self, img_name, otp):
        self.fout.write('{},{},{},{}\n'.format(img_name, otp['bpp'], otp['ms-ssim'], otp['psnr']))

    def close(self):
        self.fout.close()


def save_img(img_name, img_out_chw, out_dir):
    """val.save_img (code/val.py:215-225): <out_dir>/imgs/<name>.png, the (padded) output image as the graph produced it"""
    import os
    from PIL import Image
    img_dir = os.path.join(out_dir, 'imgs')
    os.makedirs(img_dir, exist_ok=True)
    p = os.path.join(img_dir, img_name if img_name.endswith('.png') else img_name + '.png')
    Image.fromarray(np.ascontiguousarray(np.transpose(img_out_chw, (1, 2, 0)))).save(p)
    return p


def load_images(pattern):
    """images_iterator.ImagesIterator's file side (code/images_iterator.py:20-37): sorted glob, RGB uint8 HWC,
    names without extension."""
    import glob
    import os
    from PIL import Image
    paths = sorted(glob.glob(pattern))
    assert paths, 'no images match {}'.format(pattern)
    return [np.asarray(Image.open(p).convert('RGB')) for p in paths], [os.path.splitext(os.path.basename(p))[0] for p in paths]


def main():
    import argparse
    import os
    import torch
    import torch.distributed as dist
    from . import autoencoder, config, probclass, tf_checkpoint, weights
    ap = argparse.ArgumentParser(description='val.py-style run on synthetic images / weights')
    ap.add_argument('--ae_config', default='cvpr/low')
    ap.add_argument('--pc_config', default='cvpr/res_shallow')
    ap.add_argument('--synthetic', type=int, default=24, help='number of synthetic images')
    ap.add_argument('--height', type=int, default=512)
    ap.add_argument('--width', type=int, default=768)
    ap.add_argument('--mode', default='exact', choices=['fp32', 'exact', 'fast'])
    ap.add_argument('--real_bpp', action='store_true')
    ap.add_argument('--images', default=None, help='glob of image files instead of synthetic images')
    ap.add_argument('--weights', default=None, help='.npz of TF variable name -> array, or a TensorFlow checkpoint prefix / ckpts directory of the '
                    'reference (read without TensorFlow: tf_checkpoint.py); default: seeded synthetic weights')
    ap.add_argument('--out_dir', default=None, help='write measures.csv here (rank 0 writes its own shard only)')
    ap.add_argument('--save_ours', '-o', action='store_true', help='store the output images in OUT_DIR/imgs (code/val.py:268-269)')
    ap.add_argument('--how_many', type=int, default=None, help='number of images to use (code/val.py:270)')
    args = ap.parse_args()
    world = int(os.environ.get('WORLD_SIZE', '1'))
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
    if world > 1:
        dist.init_process_group('nccl')
    a, p = config.ae_config(args.ae_config), config.pc_config(args.pc_config)
    W = tf_checkpoint.load_weights(args.weights) if args.weights else weights.synthetic_weights(
        a.num_chan_bn, a.num_centers, p.arch_param__k, a.arch_param_B)
    ae = autoencoder.get_network_cls(a)(a, weights=W, mode=args.mode)
    pc = probclass.get_network_cls(p)(p, num_centers=a.num_centers, weights=W)
    if args.images:
        imgs, names = load_images(args.images)
    else:
        imgs = list(weights.synthetic_images(args.synthetic, args.height, args.width))
        names = ['synthetic_%04d' % i for i in range(len(imgs))]
    if args.how_many is not None:
        imgs, names = imgs[:args.how_many], names[:args.how_many]
    assert not args.save_ours or args.out_dir, '--save_ours needs --out_dir'
    avgs, n, rows = validate(imgs, ae, pc, args.real_bpp, dist=dist if world > 1 else None, keep_images=args.save_ours)
    if args.out_dir:
        rank = int(os.environ.get('RANK', '0'))
        lo, hi = shard_range(len(imgs), rank, world)
        wr_dir = args.out_dir if world == 1 else os.path.join(args.out_dir, 'rank%d' % rank)
        wr = MeasuresWriter(wr_dir)
        for name, row in zip(names[lo:hi], rows):
            wr.append(name, row)
            if args.save_ours:
                save_img(name, row['img_out'], wr_dir)
        wr.close()
    if int(os.environ.get('RANK', '0')) == 0:
        print('%d images | Mean: %s' % (n, ', '.join('{}: {:.4f}'.format(k, avgs[k]) for k in METRICS)))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
