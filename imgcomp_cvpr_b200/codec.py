"""Image <-> bitstream: the pieces of --real_bpp (code/val.py:161-175,
code/bit_counter.py:13-74) arranged as a codec.

The reference never leaves the process: `_encode` writes the coded symbols to a
temporary file, `_decode` gets the shape and the first symbol handed over in
memory (code/bit_counter.py:49,64) and the round trip is only asserted.  Here
the same stream gets a small header so that it stands alone:

    compress(images)  : autoencoder.encode -> symbols -> context-model tables
                        (one batched pass, codec arithmetic) -> host range coder
                        (one thread per image) -> container bytes
    decompress(blobs) : container -> sequential decode on the device
                        (probclass.decode_streams: only the stream and the first
                        symbol go in) -> centers[symbols] -> autoencoder.decode

Container (little endian, 24 bytes + stream):
    4s magic 'ICB2' | B version | B L (number of centers) | H C (latent channels)
    I H | I W (image size before padding) | B first symbol | 3x pad | I stream bytes
Padding to a multiple of 8 is the reference's centred zero padding
(code/images_iterator.py:39-59) and is undone after decoding.

    python -m imgcomp_cvpr_b200.codec compress  in.png out.icb
    python -m imgcomp_cvpr_b200.codec decompress out.icb back.png
"""
import struct
from concurrent.futures import ThreadPoolExecutor

import numpy as np

MAGIC = b'ICB2'
MAX_SIDE = 1 << 16
VERSION = 1
_HEADER = struct.Struct('<4sBBHIIB3xI')


def pack(stream, first_sym, C, L, H, W):
    """-> container bytes"""
    return _HEADER.pack(MAGIC, VERSION, L, C, H, W, first_sym, len(stream)) + bytes(stream)


def unpack(blob):
    """container bytes -> dict(stream, first_sym, C, L, H, W); ValueError on a foreign or truncated blob"""
    blob = bytes(blob)
    if len(blob) < _HEADER.size:
        raise ValueError('not an ICB2 container: %d bytes' % len(blob))
    magic, version, L, C, H, W, first_sym, n = _HEADER.unpack_from(blob)
    if magic != MAGIC:
        raise ValueError('not an ICB2 container: magic %r' % magic)
    if version != VERSION:
        raise ValueError('ICB2 container version %d, this build reads %d' % (version, VERSION))
    if len(blob) != _HEADER.size + n:
        raise ValueError('ICB2 container truncated: header says %d stream bytes, %d present' % (n, len(blob) - _HEADER.size))
    if not (0 <= first_sym < L) or C == 0 or H == 0 or W == 0:
        raise ValueError('ICB2 container: inconsistent header')
    # sanity bounds BEFORE anything is allocated from header fields: image side <= 65536, and a range-coded stream of
    # C*h*w symbols from L <= 8 centres cannot exceed ~4 bytes per symbol
    if H > MAX_SIDE or W > MAX_SIDE or C > 1024:
        raise ValueError('ICB2 container: implausible size %dx%d, C=%d' % (H, W, C))
    hp, wp = padded_size(H, W)
    if n > 4 * C * (hp // 8) * (wp // 8) + 64:
        raise ValueError('ICB2 container: %d stream bytes for %d symbols' % (n, C * (hp // 8) * (wp // 8)))
    return {'stream': blob[_HEADER.size:], 'first_sym': first_sym, 'C': C, 'L': L, 'H': H, 'W': W}


def padded_size(H, W, f=8):
    return (H + f - 1) // f * f, (W + f - 1) // f * f


def _as_hwc(im):
    im = np.asarray(im)
    assert im.ndim == 3 and im.dtype == np.uint8, 'uint8 HWC or CHW image expected'
    hwc = im if im.shape[-1] in (3, 4) else np.transpose(im, (1, 2, 0))
    return hwc[:, :, :3]


def _encode_one(freqs, symbols):
    """one image: uint32 tables (C*h*w, L) + uint8 symbols (C*h*w,) (views into pinned staging memory) -> stream bytes"""
    from . import arithmetic_coding as ac
    enc = ac.ArithmeticEncoder()
    L = freqs.shape[-1]
    enc.write_u32(freqs.reshape(-1, L)[1:], symbols.reshape(-1)[1:])      # first symbol is side information
    return enc.finish()[0]


class _Staging(object):
    """One of the two pinned host buffers of the compress pipeline (tables as uint32 + symbols as uint8) and the event
    that says its device-to-host copies have landed."""

    def __init__(self):
        self.freqs = self.sym = self.event = None

    def fit(self, n_freq, n_sym):
        import torch
        if self.freqs is None or self.freqs.numel() < n_freq:
            self.freqs = torch.empty(n_freq, dtype=torch.int32).pin_memory()
        if self.sym is None or self.sym.numel() < n_sym:
            self.sym = torch.empty(n_sym, dtype=torch.uint8).pin_memory()
        if self.event is None:
            self.event = torch.cuda.Event()


def compress(images, ae, pc, batch_size=8, threads=8):
    """images: list of uint8 HWC (or CHW) arrays of any sizes -> list of container bytes.

    Pipeline: for every batch the GPU side -- autoencoder.encode, the codec tables of the whole latent
    in one pass (uint32), asynchronous copies of tables + uint8 symbols into one of TWO pinned staging buffers -- is
    enqueued without a single host synchronisation; the host range coder (one thread per image, the GIL is released
    inside the C call) then works on batch i while the device already computes batch i+1."""
    import torch
    from .val import add_padding
    f = ae.get_subsampling_factor()
    hwc = [_as_hwc(im) for im in images]
    padded = [np.ascontiguousarray(np.transpose(add_padding(im, f)[0], (2, 0, 1))) for im in hwc]
    by_shape = {}
    for i, p in enumerate(padded):
        by_shape.setdefault(p.shape, []).append(i)
    chunks = [idxs[s:s + batch_size] for idxs in by_shape.values() for s in range(0, len(idxs), batch_size)]
    out = [None] * len(images)
    centers_host = ae.centers_host()
    stage = [_Staging(), _Staging()]
    L = pc.L

    def enqueue(k, chunk):
        x = torch.from_numpy(np.stack([padded[i] for i in chunk])).pin_memory().cuda(non_blocking=True)
        enc = ae.encode(x, is_training=False)
        sym8 = ae.extra['symbols_u8']
        tables = pc.codec_freqs_u32(enc.symbols, centers_host)
        st = stage[k & 1]
        st.fit(tables.numel(), sym8.numel())
        st.freqs[:tables.numel()].copy_(tables.view(-1), non_blocking=True)
        st.sym[:sym8.numel()].copy_(sym8.view(-1), non_blocking=True)
        st.event.record()
        return st, tuple(enc.symbols.shape)

    def finish(pool, chunk, st, shape):
        st.event.synchronize()
        N, C, h, w = shape
        fr = st.freqs[:N * C * h * w * L].numpy().view(np.uint32).reshape(N, C * h * w, L)
        sy = st.sym[:N * C * h * w].numpy().reshape(N, C * h * w)
        streams = list(pool.map(_encode_one, fr, sy))
        for k, i in enumerate(chunk):
            out[i] = pack(streams[k], int(sy[k, 0]), C, L, hwc[i].shape[0], hwc[i].shape[1])

    with ThreadPoolExecutor(max_workers=max(1, threads)) as pool:
        pending = None
        for k, chunk in enumerate(chunks):
            cur = (chunk,) + enqueue(k, chunk)        # device work of batch k is in flight ...
            if pending is not None:
                finish(pool, *pending)                # ... while the host codes batch k-1
            pending = cur
        if pending is not None:
            finish(pool, *pending)
    return out


def decompress(blobs, ae, pc, batch_size=8):
    """list of container bytes -> list of uint8 HWC reconstructions (padding removed)."""
    import torch
    f = ae.get_subsampling_factor()
    items = [unpack(b) for b in blobs]
    centers = ae.centers_tensor()
    by_shape = {}
    for i, it in enumerate(items):
        if it['C'] != ae.config.num_chan_bn or it['L'] != centers.numel():
            raise ValueError('container was written for C=%d, L=%d; this model has C=%d, L=%d' % (
                it['C'], it['L'], ae.config.num_chan_bn, centers.numel()))
        Hp, Wp = padded_size(it['H'], it['W'], f)
        by_shape.setdefault((it['C'], Hp // f, Wp // f), []).append(i)
    out = [None] * len(items)
    for shape, idxs in by_shape.items():
        for s in range(0, len(idxs), batch_size):
            chunk = idxs[s:s + batch_size]
            sym = pc.decode_streams([items[i]['stream'] for i in chunk], [items[i]['first_sym'] for i in chunk],
                                    shape, centers)
            ae.decode(centers[sym.long()], is_training=False)
            x_u8 = ae.extra['x_out_u8'].cpu().numpy()                       # uint8 cast of val.py:91
            for k, i in enumerate(chunk):
                H, W = items[i]['H'], items[i]['W']
                Hp, Wp = x_u8.shape[2], x_u8.shape[3]
                t, l = (Hp - H) // 2, (Wp - W) // 2                          # add_padding puts the smaller half first
                out[i] = np.ascontiguousarray(np.transpose(x_u8[k, :, t:t + H, l:l + W], (1, 2, 0)))
    return out


def _models(args):
    import numpy as np
    from . import autoencoder, config, probclass, weights
    a, p = config.ae_config(args.ae_config), config.pc_config(args.pc_config)
    if args.weights:
        from . import tf_checkpoint
        W = tf_checkpoint.load_weights(args.weights)          # .npz or a TensorFlow checkpoint prefix / directory
    else:
        W = weights.synthetic_weights(a.num_chan_bn, a.num_centers, p.arch_param__k, a.arch_param_B)
    ae = autoencoder.get_network_cls(a)(a, weights=W, mode=args.mode)
    pc = probclass.get_network_cls(p)(p, num_centers=a.num_centers, weights=W)
    return ae, pc


def main():
    import argparse
    ap = argparse.ArgumentParser(description='compress / decompress one image with the CVPR autoencoder + context model')
    ap.add_argument('action', choices=['compress', 'decompress'])
    ap.add_argument('src')
    ap.add_argument('dst')
    ap.add_argument('--ae_config', default='cvpr/low')
    ap.add_argument('--pc_config', default='cvpr/res_shallow')
    ap.add_argument('--weights', default=None, help='.npz of TF variable name -> array (default: seeded synthetic weights)')
    ap.add_argument('--mode', default='exact', choices=['fp32', 'exact', 'fast'])
    args = ap.parse_args()
    from PIL import Image
    ae, pc = _models(args)
    if args.action == 'compress':
        im = np.asarray(Image.open(args.src).convert('RGB'))
        blob = compress([im], ae, pc)[0]
        with open(args.dst, 'wb') as fo:
            fo.write(blob)
        print('%s: %dx%d -> %d bytes (%.4f bpp)' % (args.src, im.shape[1], im.shape[0], len(blob),
                                                    8.0 * len(blob) / (im.shape[0] * im.shape[1])))
    else:
        with open(args.src, 'rb') as fi:
            im = decompress([fi.read()], ae, pc)[0]
        Image.fromarray(im).save(args.dst)
        print('%s -> %s: %dx%d' % (args.src, args.dst, im.shape[1], im.shape[0]))


if __name__ == '__main__':
    main()
