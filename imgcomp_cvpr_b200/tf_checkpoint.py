"""Reader for TensorFlow V2 checkpoints ("tensor bundles": `<prefix>.index` + `<prefix>.data-00000-of-00001`), the
format `tf.train.Saver` writes and the reference restores from (code/saver.py:60-128, code/val.py:96,147) -- without
TensorFlow.  -> dict *variable name -> ndarray*, the form every class of this package takes its weights in (the variable
names of the reference graph ARE the keys: SURVEY.md Appendix B).

    W = tf_checkpoint.load('ckpts/ckpt-1234567')        # or a ckpts/ directory: the newest prefix in it
    ae = autoencoder.get_network_cls(cfg)(cfg, weights=W)

Format (tensorflow/core/util/tensor_bundle, tensorflow/core/lib/io/table*: the LevelDB table format):
  .index   = sorted string table; key "" -> BundleHeaderProto, key <tensor name> -> BundleEntryProto
             {1 dtype, 2 shape {2 dim {1 size}}, 3 shard_id, 4 offset, 5 size, 6 crc32c}
             table = data blocks ... metaindex block, index block, 48-byte footer (two block handles, magic
             0xdb4775248b80fb57); block = prefix-compressed entries (shared, non_shared, value_len varints) + restart
             array; every block is followed by 1 byte compression type (0 none, 1 snappy) + 4 bytes crc.
  .data-XXXXX-of-YYYYY = raw little-endian tensor bytes at [offset, offset + size).
NOT validated against a file written by TensorFlow (none is available offline; the reference's checkpoints are a
download): tests/test_tf_checkpoint_cpu.py round-trips files produced by a writer that follows the same specification.
"""
import glob
import os
import re
import struct

import numpy as np

TABLE_MAGIC = 0xdb4775248b80fb57
# tensorflow/core/framework/types.proto
DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_,
          17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}


def _varint(buf, pos):
    result, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7f) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _snappy_decompress(buf):
    """raw snappy block format (only needed if a writer compressed the index blocks; TF's BundleWriter does not)"""
    n, pos = _varint(buf, 0)
    out = bytearray()
    while pos < len(buf):
        tag = buf[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:                                   # literal
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(buf[pos:pos + nb], 'little')
                pos += nb
            ln += 1
            out += buf[pos:pos + ln]
            pos += ln
            continue
        if kind == 1:
            ln = ((tag >> 2) & 7) + 4
            off = ((tag >> 5) << 8) | buf[pos]
            pos += 1
        elif kind == 2:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 2], 'little')
            pos += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 4], 'little')
            pos += 4
        if off == 0 or off > len(out):
            raise ValueError('corrupt snappy stream')
        for _ in range(ln):                             # may overlap its own output
            out.append(out[-off])
    if len(out) != n:
        raise ValueError('corrupt snappy stream: %d bytes, expected %d' % (len(out), n))
    return bytes(out)


def _read_block(f, offset, size):
    f.seek(offset)
    raw = f.read(size + 5)
    if len(raw) != size + 5:
        raise ValueError('truncated table block at %d' % offset)
    body, ctype = raw[:size], raw[size]
    if ctype == 1:
        body = _snappy_decompress(body)
    elif ctype != 0:
        raise ValueError('unknown block compression type %d' % ctype)
    return body


def _block_entries(block):
    """-> list of (key bytes, value bytes) of one table block"""
    if len(block) < 4:
        raise ValueError('corrupt table block')
    num_restarts = struct.unpack_from('<I', block, len(block) - 4)[0]
    limit = len(block) - 4 - 4 * num_restarts
    if limit < 0:
        raise ValueError('corrupt table block (restart array)')
    pos, key, out = 0, b'', []
    while pos < limit:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        if shared > len(key) or pos + non_shared + vlen > limit:
            raise ValueError('corrupt table block (entry)')
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        out.append((key, block[pos:pos + vlen]))
        pos += vlen
    return out


def _proto_fields(buf):
    """minimal protobuf wire-format walk -> list of (field number, wire type, value)"""
    pos, out = 0, []
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = buf[pos:pos + 8]
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            v = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError('unsupported protobuf wire type %d' % wt)
        out.append((field, wt, v))
    return out


def _signed64(v):
    return v - (1 << 64) if v >= (1 << 63) else v


# optimizer slots and counters tf.train.Saver stores beside the model variables.  The reference names its optimizers
# 'Adam_AE' / 'Adam_PC' (code/train.py:339-349), so the slots are '<var>/Adam_AE', '<var>/Adam_AE_1', '<var>/Adam_PC[_1]',
# and the second optimizer's accumulators 'beta1_power_1' / 'beta2_power_1'.
_NOT_A_MODEL_VARIABLE = re.compile(r'/(Adam\w*|Momentum\w*|RMSProp\w*)(_\d+)?$|^beta\d_power(_\d+)?$|^global_step$|ExponentialMovingAverage')


def is_model_variable(name):
    return _NOT_A_MODEL_VARIABLE.search(name) is None


def masked_crc32c(raw):
    """tensorflow/core/lib/hash/crc32c.h Mask(): what BundleEntryProto.crc32c holds for the tensor's bytes"""
    from . import _lib
    buf = np.frombuffer(raw, np.uint8)
    crc = int(_lib.lib().ic_crc32c(buf.ctypes.data, buf.size)) & 0xFFFFFFFF
    return (((crc >> 15) | (crc << 17)) + 0xa282ead8) & 0xFFFFFFFF


def _parse_entry(buf):
    e = {'dtype': 0, 'shape': [], 'shard_id': 0, 'offset': 0, 'size': 0, 'slices': 0, 'crc32c': None}
    for field, wt, v in _proto_fields(buf):
        if field == 1:
            e['dtype'] = v
        elif field == 2:                                 # TensorShapeProto
            for f2, _, v2 in _proto_fields(v):
                if f2 == 2:                              # Dim
                    size = 0
                    for f3, _, v3 in _proto_fields(v2):
                        if f3 == 1:
                            size = _signed64(v3)
                    e['shape'].append(size)
        elif field == 3:
            e['shard_id'] = v
        elif field == 4:
            e['offset'] = v
        elif field == 5:
            e['size'] = v
        elif field == 6:
            e['crc32c'] = struct.unpack('<I', v)[0] if wt == 5 else int(v)
        elif field == 7:
            e['slices'] += 1
    return e


def read_index(index_path):
    """-> (header dict, {tensor name: entry dict}) of a `.index` file"""
    with open(index_path, 'rb') as f:
        f.seek(0, os.SEEK_END)
        total = f.tell()
        if total < 48:
            raise ValueError('%s: too short for a table footer' % index_path)
        f.seek(total - 48)
        footer = f.read(48)
        if struct.unpack_from('<Q', footer, 40)[0] != TABLE_MAGIC:
            raise ValueError('%s: not a TensorFlow checkpoint index (bad table magic)' % index_path)
        pos = 0
        _, pos = _varint(footer, pos)                    # metaindex handle
        _, pos = _varint(footer, pos)
        idx_off, pos = _varint(footer, pos)
        idx_size, pos = _varint(footer, pos)
        entries = {}
        header = {'num_shards': 1}
        for _, handle in _block_entries(_read_block(f, idx_off, idx_size)):
            off, p2 = _varint(handle, 0)
            size, _ = _varint(handle, p2)
            for key, value in _block_entries(_read_block(f, off, size)):
                if key == b'':
                    for field, _, v in _proto_fields(value):
                        if field == 1:
                            header['num_shards'] = v
                        elif field == 2 and v != 0:
                            raise ValueError('big-endian checkpoints are not supported')
                else:
                    entries[key.decode()] = _parse_entry(value)
    return header, entries


def resolve_prefix(path):
    """a checkpoint prefix, or a directory (code/saver.py: <log_dir>/ckpts) -> the newest prefix in it"""
    if os.path.isdir(path):
        idx = sorted(glob.glob(os.path.join(path, '*.index')), key=os.path.getmtime)
        if not idx:
            raise FileNotFoundError('no *.index file in %s' % path)
        return idx[-1][:-len('.index')]
    return path[:-len('.index')] if path.endswith('.index') else path


def load(path, include=None, verify_crc=True):
    """-> dict tensor name -> ndarray.  include: optional predicate on the name (default: is_model_variable, i.e. without
    the optimizer slots and counters tf.train.Saver stores beside the variables).  verify_crc: check every tensor's
    bytes against the masked CRC-32C of its index entry (when the writer stored one)."""
    prefix = resolve_prefix(path)
    header, entries = read_index(prefix + '.index')
    if include is None:
        include = is_model_variable
    shards = {}
    out = {}
    for name, e in sorted(entries.items()):
        if not include(name):
            continue
        if e['slices']:
            raise ValueError('%s: partitioned (sliced) variables are not supported' % name)
        if e['dtype'] not in DTYPES:
            raise ValueError('%s: unsupported dtype enum %d' % (name, e['dtype']))
        sid = e['shard_id']
        if sid not in shards:
            shards[sid] = open('%s.data-%05d-of-%05d' % (prefix, sid, header['num_shards']), 'rb')
        f = shards[sid]
        f.seek(e['offset'])
        raw = f.read(e['size'])
        dt = np.dtype(DTYPES[e['dtype']])
        n = int(np.prod(e['shape'])) if e['shape'] else 1
        if len(raw) != e['size'] or n * dt.itemsize != e['size']:
            raise ValueError('%s: %d bytes in the data file, shape %s of %s needs %d' % (name, len(raw), e['shape'], dt, n * dt.itemsize))
        if verify_crc and e['crc32c'] is not None and masked_crc32c(raw) != e['crc32c']:
            raise ValueError('%s: CRC-32C mismatch (checkpoint data corrupted)' % name)
        out[name] = np.frombuffer(raw, dtype=dt.newbyteorder('<')).astype(dt).reshape(e['shape'])
    for f in shards.values():
        f.close()
    return out


def load_weights(path):
    """.npz (dict name -> array) or a TF checkpoint prefix / directory -> dict of float32 model variables"""
    if path.endswith('.npz'):
        return dict(np.load(path))
    return load(path)
