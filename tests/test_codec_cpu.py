"""Container of codec.py (host logic only)."""
import numpy as np
import pytest

from imgcomp_cvpr_b200 import codec


def test_container_round_trip():
    blob = codec.pack(b'\x01\x02\x03', first_sym=4, C=32, L=6, H=511, W=767)
    it = codec.unpack(blob)
    assert it == {'stream': b'\x01\x02\x03', 'first_sym': 4, 'C': 32, 'L': 6, 'H': 511, 'W': 767}
    assert codec.padded_size(511, 767) == (512, 768)
    assert codec.padded_size(512, 768) == (512, 768)


@pytest.mark.parametrize('mutate', [
    lambda b: b'JUNK' + b[4:],                      # foreign magic
    lambda b: b[:10],                               # shorter than a header
    lambda b: b[:-1],                               # truncated stream
    lambda b: b + b'\0',                            # trailing bytes
    lambda b: b[:4] + bytes([9]) + b[5:],           # unknown version
])
def test_container_rejects(mutate):
    blob = codec.pack(b'\x01\x02\x03', first_sym=0, C=32, L=6, H=64, W=64)
    with pytest.raises(ValueError):
        codec.unpack(mutate(blob))


def test_first_symbol_must_be_a_symbol():
    with pytest.raises(ValueError):
        codec.unpack(codec.pack(b'', first_sym=6, C=32, L=6, H=8, W=8))


def test_padding_offsets_match_add_padding():
    from imgcomp_cvpr_b200.val import add_padding
    rng = np.random.RandomState(0)
    for H, W in [(61, 64), (64, 59), (50, 77), (64, 64)]:
        im = rng.randint(1, 255, (H, W, 3)).astype(np.uint8)
        p, undo = add_padding(im, 8)
        Hp, Wp = codec.padded_size(H, W)
        assert p.shape[:2] == (Hp, Wp)
        t, l = (Hp - H) // 2, (Wp - W) // 2
        assert np.array_equal(p[t:t + H, l:l + W], im)
        assert np.array_equal(undo(p), im)


def test_header_fields_are_bounded_before_anything_is_allocated():
    """H / W / stream length come from an untrusted blob: implausible values are rejected in unpack()"""
    import struct
    with pytest.raises(ValueError):
        codec.unpack(codec.pack(b'\0' * 8, first_sym=0, C=32, L=6, H=1 << 20, W=64))          # side > 65536
    with pytest.raises(ValueError):
        codec.unpack(codec.pack(b'\0' * 8, first_sym=0, C=2000, L=6, H=64, W=64))             # channels
    with pytest.raises(ValueError):
        codec.unpack(codec.pack(b'\0' * (4 * 32 * 8 * 8 + 65), first_sym=0, C=32, L=6, H=64, W=64))   # > 4 bytes / symbol
    assert codec.unpack(codec.pack(b'\0' * (4 * 32 * 8 * 8 + 64), first_sym=0, C=32, L=6, H=64, W=64))['C'] == 32
