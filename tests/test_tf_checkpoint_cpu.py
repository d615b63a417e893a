"""imgcomp_cvpr_b200/tf_checkpoint.py: files written here by a minimal writer of the same specification (LevelDB table +
BundleEntryProto + raw data shard) must read back exactly.  No file written by TensorFlow itself is available offline
(tf_checkpoint.py says so): this pins the reader's handling of prefix compression, multiple data blocks, restart
arrays, varints, shapes (scalars included), dtypes and the optimizer-slot filter."""
import os
import struct

import numpy as np
import pytest

from imgcomp_cvpr_b200 import config, tf_checkpoint, weights

DT_ENUM = {np.dtype(np.float32): 1, np.dtype(np.int32): 3, np.dtype(np.int64): 9, np.dtype(np.float64): 2}


def _vi(v):
    out = bytearray()
    while True:
        b = v & 0x7f
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _field(num, wt, payload):
    return _vi((num << 3) | wt) + payload


def _entry_proto(arr, offset):
    shape = b''.join(_field(2, 2, _vi(len(d)) + d) for d in (_field(1, 0, _vi(s)) for s in arr.shape))
    return (_field(1, 0, _vi(DT_ENUM[arr.dtype])) + _field(2, 2, _vi(len(shape)) + shape) + _field(4, 0, _vi(offset)) +
            _field(5, 0, _vi(arr.nbytes)) + _field(6, 5, struct.pack('<I', tf_checkpoint.masked_crc32c(arr.tobytes()))))


def _block(items, restart_interval=16):
    out, restarts, prev = bytearray(), [], b''
    for i, (k, v) in enumerate(items):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            while shared < min(len(prev), len(k)) and prev[shared] == k[shared]:
                shared += 1
        out += _vi(shared) + _vi(len(k) - shared) + _vi(len(v)) + k[shared:] + v
        prev = k
    for r in restarts:
        out += struct.pack('<I', r)
    out += struct.pack('<I', len(restarts))
    return bytes(out)


def write_checkpoint(prefix, tensors, entries_per_block=7):
    """minimal tensor-bundle writer: one data shard, uncompressed table blocks"""
    data, items = bytearray(), [(b'', _field(1, 0, _vi(1)) + _field(2, 0, _vi(0)))]      # header: num_shards 1, little endian
    for name in sorted(tensors):
        arr = np.asarray(tensors[name])          # (ascontiguousarray would turn a scalar into shape (1,))
        items.append((name.encode(), _entry_proto(arr, len(data))))
        data += arr.tobytes()
    with open(prefix + '.data-00000-of-00001', 'wb') as f:
        f.write(bytes(data))
    table, index_items = bytearray(), []
    for i in range(0, len(items), entries_per_block):
        chunk = items[i:i + entries_per_block]
        blk = _block(chunk)
        index_items.append((chunk[-1][0], _vi(len(table)) + _vi(len(blk))))
        table += blk + b'\x00' + struct.pack('<I', 0)
    meta = _block([])
    meta_handle = _vi(len(table)) + _vi(len(meta))
    table += meta + b'\x00' + struct.pack('<I', 0)
    idx = _block(index_items, restart_interval=1)
    idx_handle = _vi(len(table)) + _vi(len(idx))
    table += idx + b'\x00' + struct.pack('<I', 0)
    footer = meta_handle + idx_handle
    footer += b'\x00' * (40 - len(footer)) + struct.pack('<Q', tf_checkpoint.TABLE_MAGIC)
    table += footer
    with open(prefix + '.index', 'wb') as f:
        f.write(bytes(table))


def test_round_trip_of_the_model_variables(tmp_path):
    a, p = config.ae_config('cvpr/low'), config.pc_config('cvpr/res_shallow')
    W = weights.synthetic_weights(a.num_chan_bn, a.num_centers, p.arch_param__k, a.arch_param_B)
    extra = {'global_step': np.asarray(1234567, np.int64), 'beta1_power': np.asarray(0.5, np.float32),
             'autoencoder/encoder/h1/weights/Adam': np.zeros((5, 5, 3, 64), np.float32),
             'autoencoder/encoder/h1/weights/Adam_1': np.ones((5, 5, 3, 64), np.float32)}
    prefix = str(tmp_path / 'ckpt-1234567')
    write_checkpoint(prefix, dict(W, **extra))
    got = tf_checkpoint.load(prefix)
    assert sorted(got) == sorted(W)                      # optimizer slots / counters filtered out
    for k in W:
        assert got[k].dtype == np.float32 and got[k].shape == np.asarray(W[k]).shape and np.array_equal(got[k], W[k]), k
    everything = tf_checkpoint.load(prefix, include=lambda n: True)
    assert everything['global_step'].shape == () and int(everything['global_step']) == 1234567
    assert np.array_equal(everything['autoencoder/encoder/h1/weights/Adam_1'], extra['autoencoder/encoder/h1/weights/Adam_1'])
    # directory form and the generic loader
    assert tf_checkpoint.resolve_prefix(str(tmp_path)) == prefix
    assert sorted(tf_checkpoint.load_weights(str(tmp_path))) == sorted(W)
    np.savez(str(tmp_path / 'w.npz'), **W)
    assert sorted(tf_checkpoint.load_weights(str(tmp_path / 'w.npz'))) == sorted(W)


def test_snappy_blocks_and_errors(tmp_path):
    raw = b'hello hello hello hello, snappy snappy snappy!' * 3
    # literal + copies hand-encoded: varint length, one literal of everything (the simplest valid stream)
    lit = tf_checkpoint._snappy_decompress(_vi(len(raw)) + bytes([(60 << 2)]) + bytes([len(raw) - 1]) + raw)
    assert lit == raw
    # a stream with a back-reference: "abcd" then copy 8 bytes from offset 4 -> "abcdabcdabcd"
    s = _vi(12) + bytes([(3 << 2)]) + b'abcd' + bytes([((8 - 4) << 2) | 1, 4])
    assert tf_checkpoint._snappy_decompress(s) == b'abcdabcdabcd'
    bad = tmp_path / 'x.index'
    bad.write_bytes(b'\x00' * 64)
    with pytest.raises(ValueError):
        tf_checkpoint.read_index(str(bad))
    empty = tmp_path / 'empty_ckpts'
    empty.mkdir()
    with pytest.raises(FileNotFoundError):
        tf_checkpoint.resolve_prefix(str(empty))


def test_crc32c_known_answers_and_corruption_is_detected(tmp_path):
    """CRC-32C (Castagnoli) check values: RFC 3720 B.4 / the usual "123456789" vector; a flipped byte in the data shard
    must be caught by the per-tensor checksum of the index"""
    from imgcomp_cvpr_b200 import _lib

    def crc(b):
        a = np.frombuffer(b, np.uint8)
        return int(_lib.lib().ic_crc32c(a.ctypes.data, a.size)) & 0xFFFFFFFF
    assert crc(b'123456789') == 0xE3069283
    assert crc(bytes(32)) == 0x8A9136AA and crc(b'\xff' * 32) == 0x62A8AB43
    assert crc(bytes(range(32))) == 0x46DD794E
    W = {'a/weights': np.arange(1000, dtype=np.float32), 'b/biases': np.ones(7, np.float32)}
    prefix = str(tmp_path / 'ckpt-1')
    write_checkpoint(prefix, W)
    assert np.array_equal(tf_checkpoint.load(prefix)['a/weights'], W['a/weights'])
    data = prefix + '.data-00000-of-00001'
    raw = bytearray(open(data, 'rb').read())
    raw[100] ^= 0x40
    open(data, 'wb').write(bytes(raw))
    with pytest.raises(ValueError, match='CRC'):
        tf_checkpoint.load(prefix)
    assert tf_checkpoint.load(prefix, verify_crc=False)['a/weights'].shape == (1000,)


def test_reference_optimizer_slot_names_are_filtered(tmp_path):
    """code/train.py:339-349 names its optimizers Adam_AE / Adam_PC: slots '<var>/Adam_AE', '<var>/Adam_AE_1',
    '<var>/Adam_PC[_1]' and the accumulators 'beta1_power[_1]', 'beta2_power[_1]' are not model variables"""
    W = {'autoencoder/encoder/h1/weights': np.ones((5, 5, 3, 64), np.float32),
         'probclass3d/logits/conv3d_conv0_mask/biases': np.zeros(24, np.float32)}
    extra = {}
    for k, v in W.items():
        opt = 'Adam_AE' if k.startswith('autoencoder') else 'Adam_PC'
        extra[k + '/' + opt] = np.zeros_like(v)
        extra[k + '/' + opt + '_1'] = np.zeros_like(v)
    for n in ('beta1_power', 'beta2_power', 'beta1_power_1', 'beta2_power_1'):
        extra[n] = np.asarray(0.5, np.float32)
    extra['global_step'] = np.asarray(7, np.int64)
    prefix = str(tmp_path / 'ckpt-7')
    write_checkpoint(prefix, dict(W, **extra))
    assert sorted(tf_checkpoint.load(prefix)) == sorted(W)
    assert all(not tf_checkpoint.is_model_variable(n) for n in extra)
    assert all(tf_checkpoint.is_model_variable(n) for n in weights.synthetic_weights())
