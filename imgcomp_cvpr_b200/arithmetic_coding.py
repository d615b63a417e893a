"""Host arithmetic coder (csrc/coder.cpp), mirroring the parts of
code/arithmetic_coding.py the --real_bpp path uses (ArithmeticEncoder /
ArithmeticDecoder over per-symbol SimpleFrequencyTables)."""
import ctypes

import numpy as np

from . import _lib


class ArithmeticEncoder(object):
    """write(freqs, symbols) / finish() -> bytes.  Batched: freqs (n,L) int64, symbols (n,)."""

    def __init__(self):
        self._h = _lib.c_void_p()
        _lib.check(_lib.lib().ic_ac_enc_create(self._h))

    def write(self, freqs, symbols):
        f = np.ascontiguousarray(np.asarray(freqs, np.int64))
        s = np.ascontiguousarray(np.atleast_1d(np.asarray(symbols, np.int64)))
        f = f.reshape(s.size, -1)
        rc = _lib.lib().ic_ac_enc_write(self._h, f.ctypes.data, f.shape[1], s.ctypes.data, s.size)
        if rc == -1:
            raise ValueError(_lib.lib().ic_last_error().decode())     # the reference raises ValueError (:94-97)
        _lib.check(rc)

    def write_u32(self, freqs, symbols):
        """the same over uint32 tables (n,L) and uint8 symbols (n,), e.g. views into pinned staging buffers: no copies"""
        f = np.asarray(freqs)
        s = np.asarray(symbols)
        assert f.dtype == np.uint32 and s.dtype == np.uint8 and f.flags.c_contiguous and s.flags.c_contiguous
        f = f.reshape(s.size, -1)
        rc = _lib.lib().ic_ac_enc_write_u32(self._h, f.ctypes.data, f.shape[1], s.ctypes.data, s.size)
        if rc == -1:
            raise ValueError(_lib.lib().ic_last_error().decode())
        _lib.check(rc)

    def finish(self):
        """-> (bytes, num_bits before byte padding)"""
        p = ctypes.POINTER(ctypes.c_uint8)()
        nb, nbits = ctypes.c_int64(), ctypes.c_int64()
        _lib.check(_lib.lib().ic_ac_enc_finish(self._h, p, nb, nbits))
        return ctypes.string_at(p, nb.value), nbits.value

    def __del__(self):
        try:
            _lib.lib().ic_ac_enc_destroy(self._h)
        except Exception:
            pass


class ArithmeticDecoder(object):
    def __init__(self, data):
        self._buf = np.frombuffer(bytes(data), np.uint8).copy()
        self._h = _lib.c_void_p()
        _lib.check(_lib.lib().ic_ac_dec_create(self._buf.ctypes.data, self._buf.size, self._h))

    def read(self, freqs):
        """freqs (n,L) int64 -> symbols (n,) int64 (one table per symbol)."""
        f = np.ascontiguousarray(np.asarray(freqs, np.int64))
        f = f.reshape(-1, f.shape[-1])
        out = np.empty(f.shape[0], np.int64)
        _lib.check(_lib.lib().ic_ac_dec_read(self._h, f.ctypes.data, f.shape[1], out.ctypes.data, f.shape[0]))
        return out

    def __del__(self):
        try:
            _lib.lib().ic_ac_dec_destroy(self._h)
        except Exception:
            pass
