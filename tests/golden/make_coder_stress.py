"""Golden bitstreams of the REFERENCE's own range coder (code/arithmetic_coding.py, imported from /root/reference) on
tables chosen to reach the corners of ArithmeticCoderBase.update (:80-115): up to 30 bits per symbol, intervals that
collapse to one value (32 shared leading bits), underflow runs of hundreds of bits, the maximal total 2^30 + 2, a
one-symbol alphabet.  tests/test_cabi_cpu.py::test_coder_stress_goldens replays them through the C coder.

    python tests/golden/make_coder_stress.py        # writes tests/golden/coder_stress.npz (needs /root/reference)
"""
import importlib.util
import io
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def cases(n=1500):
    rng = np.random.RandomState(11)
    out = []
    f = np.ones((n, 6), np.int64); f[:, 0] = 10 ** 9
    out.append(('rare', f, rng.randint(0, 6, n)))
    f = np.full((n, 2), 2 ** 29, np.int64)
    out.append(('half', f, np.tile([0, 1, 1, 0], n // 4)))
    j = rng.randint(-3, 4, n)
    out.append(('jitter', np.stack([2 ** 29 + j, 2 ** 29 - j], 1).astype(np.int64), np.arange(n) % 2))
    a, b = rng.randint(1, 2 ** 30, n), rng.randint(1, 2 ** 30, n)
    lo, hi = np.minimum(a, b), np.maximum(a, b) + 1
    f = np.stack([lo, hi - lo, 2 ** 30 + 2 - hi], 1).astype(np.int64)
    f = f[f.min(1) >= 1]
    out.append(('maxtotal', f, rng.randint(0, 3, len(f))))
    out.append(('single', np.ones((300, 1), np.int64), np.zeros(300, np.int64)))
    f = np.zeros((n, 2), np.int64); f[:, 0] = 1; f[:, 1] = 2 ** 30 + 1
    s = np.zeros(n, np.int64); s[::7] = 1
    out.append(('collapse', f, s))
    # the tiny middle symbol again and again: the interval straddles 1/2 and only underflow bits accumulate (hundreds),
    # released by one outer symbol now and then
    f = np.zeros((n, 3), np.int64); f[:, 0] = 2 ** 29 - 1; f[:, 1] = 2; f[:, 2] = 2 ** 29 - 1
    s = np.ones(n, np.int64); s[97::211] = 0; s[150::211] = 2
    out.append(('underflow', f, s))
    return out


def main():
    spec = importlib.util.spec_from_file_location('ref_arithmetic_coding', '/root/reference/code/arithmetic_coding.py')
    rac = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(rac)
    arrays = {}
    for name, f, s in cases():
        buf = io.BytesIO()
        buf.close = lambda: None
        bo = rac.BitOutputStream(buf)
        e = rac.ArithmeticEncoder(bo)
        for row, sym in zip(f, s):
            e.write(rac.SimpleFrequencyTable([int(v) for v in row]), int(sym))
        e.finish()
        bo.close()
        arrays[name + '_freqs'] = f
        arrays[name + '_symbols'] = s.astype(np.int64)
        arrays[name + '_stream'] = np.frombuffer(buf.getvalue(), np.uint8)
        print('%-10s %5d symbols -> %6d bytes' % (name, len(s), len(buf.getvalue())))
    np.savez_compressed(os.path.join(HERE, 'coder_stress.npz'), **arrays)


if __name__ == '__main__':
    main()
