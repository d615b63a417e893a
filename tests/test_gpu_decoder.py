"""Decoder side of --real_bpp: sequential context-model decode on the device
(ic_pc_decode_fwd) against the batched encoder pass (ic_pc_codec_freqs_fwd).

The reference's property (code/bit_counter.py:59-68): the symbols decoded from
the file equal the encoded ones.  Here additionally: every table the decoder
derives from its own output is bit-identical to the encoder's table."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _streams(pc, centers, sym):
    """sym N,C,h,w int64 cuda -> ([stream bytes], [first symbols], freqs numpy N,C,h,w,L)"""
    from imgcomp_cvpr_b200 import arithmetic_coding as ac
    f, _ = pc.freqs(sym, centers, codec=True)
    f = f.cpu().numpy()
    s = sym.cpu().numpy()
    streams, firsts = [], []
    for n in range(s.shape[0]):
        enc = ac.ArithmeticEncoder()
        enc.write(f[n].reshape(-1, f.shape[-1])[1:], s[n].reshape(-1)[1:])
        b, _ = enc.finish()
        streams.append(bytes(b))
        firsts.append(int(s[n].reshape(-1)[0]))
    return streams, firsts, f


def _rand_symbols(shape, seed, L=6):
    """spatially correlated symbols: cheap to code, exercises every table entry"""
    rng = np.random.RandomState(seed)
    base = rng.randint(0, L, shape)
    smooth = np.repeat(np.repeat(rng.randint(0, L, (shape[0], shape[1], (shape[2] + 3) // 4, (shape[3] + 3) // 4)), 4, 2), 4, 3)
    smooth = smooth[:, :, :shape[2], :shape[3]]
    pick = rng.rand(*shape) < 0.7
    return torch.from_numpy(np.where(pick, smooth, base).astype(np.int64)).cuda()


@pytest.mark.parametrize('shape', [(1, 8, 5, 7), (2, 32, 8, 8), (1, 5, 16, 3), (3, 4, 1, 1), (1, 32, 12, 20)])
def test_decoder_tables_match_encoder(gpu_models, shape):
    """teacher forcing: the tables the sequential kernel builds from cached activations are the
    encoder's, bit for bit, at every position (edges, tiny and ragged volumes included)."""
    ae, pc, W = gpu_models('cvpr/low')
    centers = torch.from_numpy(W['autoencoder/encoder/centers']).cuda()
    sym = _rand_symbols(shape, 1)
    streams, firsts, f = _streams(pc, centers, sym)
    out, seen = pc.decode_streams(streams, firsts, shape[1:], centers, force_symbols=sym, return_freqs=True)
    assert np.array_equal(seen.cpu().numpy(), f)
    assert np.array_equal(out.cpu().numpy(), sym.cpu().numpy())


@pytest.mark.parametrize('shape', [(1, 8, 5, 7), (2, 32, 8, 8), (3, 4, 1, 1), (1, 32, 12, 20)])
def test_decode_round_trip(gpu_models, shape):
    """only the bitstream and the first symbol go in; the symbols come back"""
    ae, pc, W = gpu_models('cvpr/low')
    centers = torch.from_numpy(W['autoencoder/encoder/centers']).cuda()
    sym = _rand_symbols(shape, 2)
    streams, firsts, f = _streams(pc, centers, sym)
    out, seen = pc.decode_streams(streams, firsts, shape[1:], centers, return_freqs=True)
    out = out.cpu().numpy()
    ref = sym.cpu().numpy()
    if not np.array_equal(out, ref):
        bad = np.argwhere(out != ref)[0]
        raise AssertionError('first mismatch at (n,c,y,x) = %s; tables equal up to there: %s' % (
            bad, np.array_equal(seen.cpu().numpy()[tuple(bad)], f[tuple(bad)])))
    assert np.array_equal(seen.cpu().numpy(), f)


def test_decode_golden_image_symbols(gpu_models):
    """the latent of the golden image: encode -> stream -> decode -> same symbols -> same reconstruction;
    the stream has the size the reference's coder produced for these symbols (+- table rounding)"""
    ae, pc, W = gpu_models('cvpr/low')
    g = load_golden('tiny_low_1x64x64')
    x = torch.from_numpy(g['x_u8']).cuda()
    enc = ae.encode(x, False)
    centers = ae.get_centers_variable()
    streams, firsts, _ = _streams(pc, centers, enc.symbols)
    if np.array_equal(enc.symbols.cpu().numpy(), g['symbols']):
        assert abs(len(streams[0]) * 8 - int(g['real_bits'])) <= 16
    out = pc.decode_streams(streams, firsts, enc.symbols.shape[1:], centers)
    assert torch.equal(out.long(), enc.symbols)
    x1 = ae.decode(centers[out.long()], False)
    x0 = ae.decode(enc.qhard, False)
    assert torch.equal(x0, x1)


def test_codec_tables_and_stream_against_reference_run_golden(gpu_models):
    """The tables a real stream is coded with (codec=True: the float32 chain the decoder re-derives) against the tables the
    reference's own per-symbol loop produced for the same symbols (tests/golden/make_golden.py: PredictionNetwork.get_freqs
    through code/bit_counter.py:103-134): within 256 of 1e9 counts (the numpy oracle is within 128: float32 softmax
    rounding), the stream within 2 bytes of the reference's, byte-identical whenever every table is."""
    from imgcomp_cvpr_b200 import arithmetic_coding as ac
    ae, pc, W = gpu_models('cvpr/low')
    g = load_golden('tiny_low_1x64x64')
    centers = torch.from_numpy(W['autoencoder/encoder/centers']).cuda()
    sym = torch.from_numpy(g['symbols'].astype(np.int64)).cuda()
    f, bits = pc.freqs(sym, centers, codec=True)
    f = f[0].cpu().numpy()
    d = np.abs(f - g['freqs'])
    print('codec tables vs reference-run golden: max |df| %d, tables differing %d / %d' % (d.max(), (d.max(-1) > 0).sum(), d[..., 0].size))
    assert d.max() <= 256
    assert abs(bits.item() - float(g['theory_bits'])) < 0.05
    enc = ac.ArithmeticEncoder()
    s_flat = g['symbols'].reshape(-1).astype(np.int64)
    enc.write(f.reshape(-1, 6)[1:], s_flat[1:])
    stream = np.frombuffer(enc.finish()[0], np.uint8)
    assert abs(len(stream) - len(g['bitstream'])) <= 2
    if d.max() == 0:
        assert np.array_equal(stream, g['bitstream'])
    # the reference's stream itself (coded with the reference's tables) decodes on the device up to the first position
    # whose table differs; with identical tables it decodes completely
    if d.max() == 0:
        out = pc.decode_streams([g['bitstream'].tobytes()], [int(s_flat[0])], (32, 8, 8), centers)
        assert np.array_equal(out[0].cpu().numpy().astype(np.int64), g['symbols'][0])
    # default (tensor-core) tables: same distribution, hi/lo arithmetic instead of the float32 chain
    fb, _ = pc.freqs(sym, centers)
    assert np.abs(fb[0].cpu().numpy() - g['freqs']).max() <= 3e4


def test_codec_tables_close_to_default_tables(gpu_models):
    """codec tables (float32 chain) vs the default tensor-core tables: same distribution"""
    ae, pc, W = gpu_models('cvpr/low')
    centers = torch.from_numpy(W['autoencoder/encoder/centers']).cuda()
    sym = _rand_symbols((1, 32, 8, 8), 3)
    fa, ba = pc.freqs(sym, centers, codec=True)
    fb, bb = pc.freqs(sym, centers)
    assert (fa - fb).abs().max().item() <= 3e4
    assert abs(ba.item() - bb.item()) < 0.05


def test_decode_kodak_sized_volume(gpu_models):
    """full Kodak latent (32 x 96 x 64), two images at once; reports the decode time"""
    ae, pc, W = gpu_models('cvpr/low')
    centers = torch.from_numpy(W['autoencoder/encoder/centers']).cuda()
    sym = _rand_symbols((2, 32, 96, 64), 4)
    streams, firsts, _ = _streams(pc, centers, sym)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    out = pc.decode_streams(streams, firsts, sym.shape[1:], centers)
    t1.record()
    torch.cuda.synchronize()
    print('decode of 2 x %d symbols: %.1f ms' % (sym[0].numel(), t0.elapsed_time(t1)))
    assert torch.equal(out.long(), sym)


def test_decode_rejects_unbuilt_configs(gpu_models):
    ae, pc, W = gpu_models('cvpr/low')
    centers = torch.from_numpy(W['autoencoder/encoder/centers']).cuda()
    with pytest.raises(RuntimeError):
        pc.decode_streams([b'\0' * 8], [0], (4, 4, 600), centers)      # latent width beyond one CTA's shared memory
