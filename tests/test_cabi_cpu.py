"""CPU-only checks of the C-ABI library: it loads without a GPU, exports every
symbol include/imgcomp_b200.h declares, and its host range coder reproduces the
reference's arithmetic coder bit for bit (golden bitstream written by
code/bit_counter.py running on the TF1 shim)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden
from imgcomp_cvpr_b200 import _lib, arithmetic_coding as ac, config as cfgmod


def _declared_symbols():
    with open(os.path.join(ROOT, 'include', 'imgcomp_b200.h')) as f:
        src = f.read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(ic_[a-z0-9_]+)\s*\(', src)))


def test_library_loads_and_exports_every_declared_symbol():
    names = _declared_symbols()
    assert len(names) >= 30
    L = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(L, n), n
    assert sorted(_lib.SIGNATURES) == names
    assert _lib.lib().ic_abi_version() == 1


def test_tensor_schema_matches_synthetic_weights(synth):
    L = _lib.lib()
    for ae_name in ('cvpr/low', 'cvpr/hi'):
        a, p, W = synth(ae_name)
        c = _lib.AeConfig(a.num_chan_bn, a.arch_param_B, a.num_centers, 1, 1)
        n = L.ic_ae_num_tensors(c)
        pcfg = _lib.PcConfig(p.kernel_size, p.arch_param__k, a.num_centers)
        names = [L.ic_ae_tensor_name(c, i).decode() for i in range(n)]
        names += [L.ic_pc_tensor_name(pcfg, i).decode() for i in range(L.ic_pc_num_tensors(pcfg))]
        assert sorted(names) == sorted(W)
        for i in range(n):
            assert W[names[i]].size == L.ic_ae_tensor_numel(c, i)
        for i in range(8):
            assert W[L.ic_pc_tensor_name(pcfg, i).decode()].size == L.ic_pc_tensor_numel(pcfg, i)
    assert L.ic_pc_num_tensors(_lib.PcConfig(5, 24, 6)) < 0        # kernel_size 5 is not on the hot path


def test_coder_reproduces_reference_bitstream():
    g = load_golden('tiny_low_1x64x64')
    syms = g['symbols'][0].astype(np.int64).reshape(-1)
    freqs = g['freqs'].reshape(-1, 6)
    enc = ac.ArithmeticEncoder()
    enc.write(freqs[1:], syms[1:])                 # first symbol is side info (bit_counter.py:118-121)
    stream, nbits = enc.finish()
    assert np.array_equal(np.frombuffer(stream, np.uint8), g['bitstream'])
    assert len(stream) * 8 == int(g['real_bits'])
    dec = ac.ArithmeticDecoder(stream)
    assert np.array_equal(dec.read(freqs[1:]), syms[1:])


@pytest.mark.parametrize('seed', [0, 1, 2])
def test_coder_round_trip_random_tables(seed):
    rng = np.random.RandomState(seed)
    n, L = 5000, 6
    p = rng.dirichlet(np.ones(L) * (0.05 if seed else 1.0), size=n).astype(np.float32)
    freqs = np.maximum((p * np.float32(1e9)).astype(np.int64), 1)
    syms = np.array([rng.choice(L, p=q / q.sum()) for q in p.astype(np.float64)])
    enc = ac.ArithmeticEncoder()
    for lo in range(0, n, 777):                    # chunked writes == one write
        enc.write(freqs[lo:lo + 777], syms[lo:lo + 777])
    stream, nbits = enc.finish()
    ideal = -np.log2(freqs[np.arange(n), syms] / freqs.sum(1)).sum()
    assert abs(nbits - ideal) < 50
    assert np.array_equal(ac.ArithmeticDecoder(stream).read(freqs), syms)


def test_coder_empty_and_errors():
    enc = ac.ArithmeticEncoder()
    stream, nbits = enc.finish()
    assert nbits == 1 and stream == b'\x80'        # finish() writes a single 1 bit (arithmetic_coding.py:146-147)
    enc = ac.ArithmeticEncoder()
    with pytest.raises(ValueError):                # total > MAX_TOTAL = 2^30 + 2 (arithmetic_coding.py:96-97)
        enc.write(np.full((1, 6), 2 ** 29, np.int64), [0])
    with pytest.raises(ValueError):                # zero frequency
        enc.write(np.array([[0, 5, 5, 5, 5, 5]], np.int64), [0])


def test_coder_matches_reference_module_when_available():
    ref = '/root/reference/code'
    if not os.path.isdir(ref):
        pytest.skip('reference tree not present on this box')
    import importlib.util
    import io
    spec = importlib.util.spec_from_file_location('ref_arithmetic_coding', os.path.join(ref, 'arithmetic_coding.py'))
    rac = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(rac)
    rng = np.random.RandomState(7)
    n, L = 3000, 6
    p = rng.dirichlet(np.ones(L) * 0.3, size=n).astype(np.float32)
    freqs = np.maximum((p * np.float32(1e9)).astype(np.int64), 1)
    syms = rng.randint(0, L, n)
    buf = io.BytesIO()
    buf.close = lambda: None
    bo = rac.BitOutputStream(buf)
    e = rac.ArithmeticEncoder(bo)
    for f, s in zip(freqs, syms):
        e.write(rac.SimpleFrequencyTable([int(v) for v in f]), int(s))
    e.finish()
    bo.close()
    enc = ac.ArithmeticEncoder()
    enc.write(freqs, syms)
    stream, _ = enc.finish()
    assert stream == buf.getvalue()


def test_config_parser_and_builtin_values(tmp_path):
    (tmp_path / 'ae_configs' / 'cvpr').mkdir(parents=True)
    (tmp_path / 'ae_configs' / 'base').write_text(
        'num_chan_bn = 3*8*8\nconstrain normalization :: OFF, FIXED\nnormalization = FIXED\nheatmap = True\n'
        "arch = 'TwoLayerNet'\ncrop_size = (128, 128)  # comment\nH_target = None\n")
    (tmp_path / 'ae_configs' / 'cvpr' / 'low').write_text("use ../base\n\nnum_chan_bn = 32\nH_target = 2*0.2\narch = 'CVPR'\n")
    c, rel = cfgmod.parse(str(tmp_path / 'ae_configs' / 'cvpr' / 'low'))
    assert (c.num_chan_bn, c.normalization, c.heatmap, c.arch, c.crop_size, c.H_target) == \
        (32, 'FIXED', True, 'CVPR', (128, 128), 0.4)
    assert rel == os.path.join('cvpr', 'low')
    (tmp_path / 'bad').write_text('constrain normalization :: OFF, FIXED\nnormalization = SOMETHING\n')
    with pytest.raises(Exception):
        cfgmod.parse(str(tmp_path / 'bad'))
    assert cfgmod.ae_config('cvpr/hi').num_chan_bn == 64 and cfgmod.ae_config('cvpr/high').H_target == 1.0
    assert cfgmod.pc_config('cvpr/res_shallow_64').arch_param__k == 64
    ref = '/root/reference/code'
    if os.path.isdir(ref):
        for name in ('cvpr/low', 'cvpr/med', 'cvpr/hi'):
            parsed, _ = cfgmod.parse(os.path.join(ref, 'ae_configs', name))
            for k, v in cfgmod.ae_config(name).__dict__.items():
                assert getattr(parsed, k) == v, (name, k)
        for name in ('cvpr/res_shallow', 'cvpr/res_shallow_64'):
            parsed, _ = cfgmod.parse(os.path.join(ref, 'pc_configs', name))
            for k, v in cfgmod.pc_config(name).__dict__.items():
                assert getattr(parsed, k) == v, (name, k)


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU box')
    from imgcomp_cvpr_b200 import autoencoder
    a = cfgmod.ae_config('cvpr/low')
    with pytest.raises(_lib.IcError):
        autoencoder.get_network_cls(a)(a, weights={})


def test_library_sass_has_tcgen05_and_tma_and_no_legacy_mma():
    """cuobjdump -sass of the built library (no GPU needed): the conv and filter-gradient kernels issue tcgen05 MMAs
    (UTCHMMA), read accumulators with tcgen05.ld (LDTM) and are fed by TMA (UTMALDG); nothing falls back to the legacy
    mma.sync path (HMMA).  B200_PROFILING.md lists these mnemonics as the evidence of a Blackwell-native kernel."""
    import re
    import shutil
    import subprocess
    from imgcomp_cvpr_b200 import _lib
    if shutil.which('cuobjdump') is None:
        pytest.skip('cuobjdump not available')
    sass = subprocess.run(['cuobjdump', '-sass', _lib.LIB_PATH], capture_output=True, text=True, timeout=300).stdout
    per_kernel, cur = {}, None
    for line in sass.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = m.group(1)
            per_kernel[cur] = set()
        elif cur is not None:
            for mn in ('UTCHMMA', 'LDTM', 'UTMALDG', 'HMMA'):
                if re.search(r'\b' + mn + r'[\.\s]', line):
                    per_kernel[cur].add(mn)
    conv = [k for k in per_kernel if 'conv_tc_kernel' in k]
    wgrad = [k for k in per_kernel if 'wgrad_tc_kernel' in k]
    assert len(conv) >= 10 and len(wgrad) == 1
    for k in conv + wgrad:
        assert {'UTCHMMA', 'LDTM', 'UTMALDG'} <= per_kernel[k], (k, per_kernel[k])
    assert not [k for k, v in per_kernel.items() if 'HMMA' in v]


@pytest.mark.parametrize('case', ['rare', 'half', 'jitter', 'maxtotal', 'single', 'collapse', 'underflow'])
def test_coder_stress_goldens(case):
    """Streams the reference's own coder (code/arithmetic_coding.py, run by tests/golden/make_coder_stress.py) wrote for
    tables that reach the corners of ArithmeticCoderBase.update (:80-115) -- up to 30 bits per symbol, intervals collapsing
    to one value (32 shared leading bits), underflow runs of hundreds of bits, total = 2^30 + 2, a one-symbol alphabet --
    against the closed-form / branch-free update of csrc/coder.cpp: both encoder entry points byte for byte, chunked
    writes, and the decoder."""
    g = load_golden('coder_stress')
    f, s, want = g[case + '_freqs'], g[case + '_symbols'], g[case + '_stream'].tobytes()
    e = ac.ArithmeticEncoder()
    for lo in range(0, len(s), 333):
        e.write(np.ascontiguousarray(f[lo:lo + 333]), np.ascontiguousarray(s[lo:lo + 333]))
    got, nbits = e.finish()
    assert bytes(got) == want and (nbits + 7) // 8 == len(want)
    e = ac.ArithmeticEncoder()
    e.write_u32(np.ascontiguousarray(f.astype(np.uint32)), np.ascontiguousarray(s.astype(np.uint8)))
    got32, nbits32 = e.finish()
    assert bytes(got32) == want and nbits32 == nbits
    assert np.array_equal(ac.ArithmeticDecoder(want).read(np.ascontiguousarray(f)), s)


def test_coder_u32_tables_give_the_same_stream():
    """ic_ac_enc_write_u32 (uint32 tables + uint8 symbols: what the compress pipeline stages in pinned memory) codes
    exactly what ic_ac_enc_write codes from int64 tables, including the golden bitstream of the reference's coder"""
    from imgcomp_cvpr_b200 import arithmetic_coding as ac
    g = load_golden('tiny_low_1x64x64')
    syms = g['symbols'].reshape(-1).astype(np.int64)
    freqs = g['freqs'].reshape(-1, 6)
    enc = ac.ArithmeticEncoder()
    enc.write_u32(np.ascontiguousarray(freqs[1:].astype(np.uint32)), np.ascontiguousarray(syms[1:].astype(np.uint8)))
    stream, nbits = enc.finish()
    assert np.array_equal(np.frombuffer(stream, np.uint8), g['bitstream']) and (nbits + 7) // 8 * 8 == int(g['real_bits'])
    bad = freqs[1:3].astype(np.uint32).copy()
    bad[0, int(syms[1])] = 0
    with pytest.raises(ValueError):
        ac.ArithmeticEncoder().write_u32(bad, np.ascontiguousarray(syms[1:3].astype(np.uint8)))
