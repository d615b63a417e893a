"""compress -> bytes -> decompress (codec.py): what comes back is exactly what val.py would have
reconstructed from the encoder's own symbols, and the container has the size the context model predicts."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_codec_round_trip_mixed_sizes(gpu_models):
    from imgcomp_cvpr_b200 import codec, weights as wm
    from imgcomp_cvpr_b200.val import add_padding
    ae, pc, W = gpu_models('cvpr/low', 'exact')
    imgs = [np.transpose(x, (1, 2, 0)) for x in wm.synthetic_images(3, 64, 96, seed=5)]
    imgs += [np.transpose(x, (1, 2, 0))[:61, :75] for x in wm.synthetic_images(2, 64, 80, seed=6)]     # needs padding
    blobs = codec.compress(imgs, ae, pc, batch_size=2)
    assert all(isinstance(b, bytes) for b in blobs)
    recon = codec.decompress(blobs, ae, pc, batch_size=4)
    for im, blob, rec in zip(imgs, blobs, recon):
        assert rec.shape == im.shape and rec.dtype == np.uint8
        p, undo = add_padding(np.ascontiguousarray(im), 8)
        x = torch.from_numpy(np.ascontiguousarray(np.transpose(p, (2, 0, 1)))[None]).cuda()
        enc = ae.encode(x, False)
        ae.decode(enc.qhard, False)
        want = undo(np.transpose(ae.extra['x_out_u8'][0].cpu().numpy(), (1, 2, 0)))
        assert np.array_equal(rec, want)
        # container size vs the cross-entropy estimate (val.py:174 / bit_counter.py:51)
        bits = pc.bitcost(enc.qbar, enc.symbols, False, pad_value=pc.auto_pad_value(ae)).sum().item()
        assert abs(8 * (len(blob) - 24) - bits) < 64


def test_codec_rejects_foreign_model(gpu_models):
    from imgcomp_cvpr_b200 import codec
    ae, pc, W = gpu_models('cvpr/low', 'exact')
    blob = codec.pack(b'\0' * 8, first_sym=0, C=64, L=6, H=64, W=64)
    with pytest.raises(ValueError):
        codec.decompress([blob], ae, pc)
