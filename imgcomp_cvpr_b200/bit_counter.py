"""Host mirror of code/bit_counter.py: arithmetic-code a symbol volume with the
context model's probabilities, decode the stream again and check the round trip.

The reference evaluates the network once per symbol (two sess.run per symbol on
encode, one on decode; README.md:65-66: ~350 s + ~200 s per Kodak image).  Here
  encode: ONE batched probclass pass emits every position's table
          (PredictionNetwork.get_all_freqs(codec=True)), the host range coder
          consumes them in the same raster order;
  decode: a real sequential decode on the device -- only the bitstream and the
          first symbol go in (PredictionNetwork.decode_symbols): context model
          with cached activations + range decoder in one kernel, bit-consistent
          with the encoder's tables."""
import numpy as np

from . import arithmetic_coding as ac


def encode_decode_to_file_ctx(syms, prediction_net, syms_format='HWC', verbose=False, return_stream=False):
    """Encode symbols with arithmetic coding, decode again, assert equality.
    :param syms: HWC / CHW (or BHWC / BCHW: summed over the batch) symbols of one image
    :return: number of bits to encode all symbols in `syms` (code/bit_counter.py:13-74)"""
    syms = np.asarray(syms)
    if syms.ndim == 4:
        return np.sum([encode_decode_to_file_ctx(syms[b], prediction_net, syms_format, verbose)
                       for b in range(syms.shape[0])])
    assert syms.ndim == 3, 'Expected HWC or CHW'
    assert syms_format in ('HWC', 'CHW')
    if syms_format == 'HWC':
        syms = np.transpose(syms, (2, 0, 1))
    freqs, theoretical_bit_cost = prediction_net.get_all_freqs(syms, codec=True)   # (C,h,w,L), raster C -> H -> W
    flat_f = freqs.reshape(-1, freqs.shape[-1])
    flat_s = syms.reshape(-1).astype(np.int64)
    first_sym = flat_s[0]            # the first symbol is side information, not coded (bit_counter.py:118-121)
    enc = ac.ArithmeticEncoder()
    enc.write(flat_f[1:], flat_s[1:])
    stream, nbits = enc.finish()
    virtual_num_bits = nbits + (8 - nbits % 8) % 8                          # CountingBitOutputStream.close
    assert abs(virtual_num_bits - theoretical_bit_cost) < 50, 'Virtual: {} -- Theoretical: {}'.format(
        virtual_num_bits, theoretical_bit_cost)                             # bit_counter.py:51
    actual_num_bits = len(stream) * 8
    assert actual_num_bits == virtual_num_bits, '{} != {}'.format(actual_num_bits, virtual_num_bits)   # :56
    # decode from the stream alone (bit_counter.py:59-68): the decoder re-derives every table from
    # the symbols it has already produced
    syms_dec = prediction_net.decode_symbols(stream, first_sym, syms.shape)
    np.testing.assert_array_equal(syms, syms_dec)                           # :68
    if return_stream:
        return actual_num_bits, stream
    return actual_num_bits
