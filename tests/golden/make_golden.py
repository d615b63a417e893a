"""Generate golden vectors by running the UNMODIFIED reference modules
(/root/reference/code/*.py) on the numpy/torch TF1 shim (tests/tf1_shim).

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Writes tests/golden/*.npz.  The graphs built below are the ones code/val.py
builds (val.py:78-94, 131-134) and the training-loss MS-SSIM (train.py:431).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.normpath(os.path.join(HERE, '..', '..'))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import tf1_shim                                              # noqa: E402
from imgcomp_cvpr_b200 import config as cfgmod               # noqa: E402
from imgcomp_cvpr_b200 import weights as wm                  # noqa: E402

REF = '/root/reference/code'
tf = tf1_shim.install(REF)
import autoencoder, probclass, bits, ms_ssim, ms_ssim_np, bit_counter, bpp_helpers   # noqa: E402,E401
import arithmetic_coding                                     # noqa: E402,F401


def ref_configs(ae_name, pc_name):
    """Parse the reference's own config files with our parser and check the
    built-in effective values against them."""
    ae_cfg, _ = cfgmod.parse(os.path.join(REF, 'ae_configs', ae_name))
    pc_cfg, _ = cfgmod.parse(os.path.join(REF, 'pc_configs', pc_name))
    for built, parsed in ((cfgmod.ae_config(ae_name), ae_cfg), (cfgmod.pc_config(pc_name), pc_cfg)):
        for k, v in built.__dict__.items():
            assert getattr(parsed, k) == v, (k, getattr(parsed, k), v)
    return ae_cfg, pc_cfg


def run_val_graph(x_u8, W, ae_cfg, pc_cfg, real_bpp=False):
    tf1_shim.WEIGHTS.clear()
    tf1_shim.WEIGHTS.update(W)
    ae = autoencoder.get_network_cls(ae_cfg)(ae_cfg)
    pc = probclass.get_network_cls(pc_cfg)(pc_cfg, num_centers=ae_cfg.num_centers)
    out = {}
    N = x_u8.shape[0]
    per = {k: [] for k in ('bpp', 'ms_ssim_np')}
    keys = ('qbar', 'qhard', 'symbols', 'z', 'heatmap', 'bitcost', 'x_out', 'x_out_u8')
    acc = {k: [] for k in keys}
    for i in range(N):                                     # reference val.py runs batch = 1
        x_val_uint8 = tf.constant(x_u8[i:i + 1])
        x_val = tf.to_float(x_val_uint8)
        enc = ae.encode(x_val, is_training=False)
        x_out = ae.decode(enc.qhard, is_training=False)
        bc = pc.bitcost(enc.qbar, enc.symbols, is_training=False, pad_value=pc.auto_pad_value(ae))
        bpp = bits.bitcost_to_bpp(bc, x_val)
        x_out_u8 = tf.cast(x_out, tf.uint8)
        msssim = ms_ssim_np.tf_msssim_np(x_val_uint8, x_out_u8, data_format='NCHW')
        per['bpp'].append(np.float32(bpp.value))
        per['ms_ssim_np'].append(np.float32(msssim.value))
        for k, t in zip(keys, (enc.qbar, enc.qhard, enc.symbols, enc.z, enc.heatmap, bc, x_out, x_out_u8)):
            acc[k].append(t.value[0])
    for k in keys:
        out[k] = np.stack(acc[k])
    out['symbols'] = out['symbols'].astype(np.int8)
    out['bpp'] = np.array(per['bpp'])
    out['ms_ssim_np'] = np.array(per['ms_ssim_np'])
    # training-loss MS-SSIM: one scalar for the batch (train.py:431, ms_ssim.py:115)
    try:
        v = ms_ssim.MultiScaleSSIM(tf.constant(x_u8.astype(np.float32)), tf.constant(out['x_out']),
                                   data_format='NCHW')
        out['ms_ssim_tf'] = np.float32(v.value)
    except RuntimeError as e:       # the reference's blur pads by W only (ms_ssim.py:24-29): H < taps raises
        print('reference ms_ssim.MultiScaleSSIM raises on', x_u8.shape, ':', str(e)[:80])
        out['ms_ssim_tf'] = np.float32(np.nan)
        out['ms_ssim_tf_raises'] = np.int32(1)
    if real_bpp:
        sess = tf.Session()
        pred = probclass.PredictionNetwork(pc, pc_cfg, ae.get_centers_variable(), sess)
        checker = probclass.ProbclassNetworkTesting(pc, ae, sess)
        syms = out['symbols'][0].astype(np.int64)          # CHW
        # capture the bitstream the reference writes before it deletes the file
        captured = {}
        real_remove = os.remove

        def grab(p):
            with open(p, 'rb') as f:
                captured['bytes'] = f.read()
            real_remove(p)
        bit_counter.os.remove = grab
        try:
            nbits = bit_counter.encode_decode_to_file_ctx(syms, pred, syms_format='CHW')
        finally:
            bit_counter.os.remove = real_remove
        out['real_bits'] = np.int64(nbits)
        out['bitstream'] = np.frombuffer(captured['bytes'], np.uint8)
        out['theory_bits'] = np.float32(checker.get_total_bit_cost(syms))
        # every context's frequencies, in the reference's raster order (probclass.py:382-387)
        padded = pred.pad_symbols_volume(syms)
        freqs = [pred.get_freqs(ctx) for ctx in probclass.iter_over_blocks(padded, pred.input_ctx_shape)]
        out['freqs'] = np.array(freqs, np.int64).reshape(syms.shape + (ae_cfg.num_centers,))
    return out


def main():
    cases = [
        # name, ae cfg, N, H, W, real_bpp
        ('tiny_low_1x64x64', 'cvpr/low', 1, 64, 64, True),
        ('ragged_low_2x48x72', 'cvpr/low', 2, 48, 72, False),
        ('cfg1_low_1x128x128', 'cvpr/low', 1, 128, 128, False),
        ('tiny_hi_1x40x24', 'cvpr/hi', 1, 40, 24, False),
    ]
    for name, ae_name, N, H, Wd, real in cases:
        ae_cfg, pc_cfg = ref_configs(ae_name, 'cvpr/res_shallow')
        W = wm.synthetic_weights(ae_cfg.num_chan_bn, ae_cfg.num_centers, pc_cfg.arch_param__k, ae_cfg.arch_param_B)
        x = wm.synthetic_images(N, H, Wd, seed=1234)
        out = run_val_graph(x, W, ae_cfg, pc_cfg, real_bpp=real)
        out['x_u8'] = x
        if name.startswith('cfg1'):                        # keep the fixture small
            for k in ('x_out', 'heatmap', 'qhard'):
                out.pop(k)
        p = os.path.join(HERE, name + '.npz')
        np.savez_compressed(p, **out)
        print(name, 'bpp', out['bpp'], 'ms-ssim(np)', out['ms_ssim_np'], 'ms-ssim(tf)', out['ms_ssim_tf'],
              'real_bits', out.get('real_bits'), 'theory', out.get('theory_bits'),
              '%.1f KB' % (os.path.getsize(p) / 1024))

    # MS-SSIM goldens on image pairs that are actually similar (decoder output of
    # random weights is not), incl. odd sizes after downsampling
    rng = np.random.RandomState(5)
    ms = {}
    for tag, (N, H, Wd) in (('a', (2, 160, 160)), ('b', (1, 176, 264)), ('c', (1, 128, 128))):
        a = wm.synthetic_images(N, H, Wd, seed=99)
        b = np.clip(a.astype(np.int32) + rng.randint(-12, 13, a.shape), 0, 255).astype(np.uint8)
        ms['x_' + tag], ms['y_' + tag] = a, b
        ms['np_' + tag] = np.array([ms_ssim_np.MultiScaleSSIM(a[i:i + 1].transpose(0, 2, 3, 1),
                                                              b[i:i + 1].transpose(0, 2, 3, 1)) for i in range(N)])
        ms['tf_' + tag] = np.float32(ms_ssim.MultiScaleSSIM(tf.constant(a.astype(np.float32)),
                                                            tf.constant(b.astype(np.float32)),
                                                            data_format='NCHW').value)
        print('ms-ssim', tag, ms['np_' + tag], ms['tf_' + tag])
    np.savez_compressed(os.path.join(HERE, 'msssim_pairs.npz'), **ms)


if __name__ == '__main__':
    main()
