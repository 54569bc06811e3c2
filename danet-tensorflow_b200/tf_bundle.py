"""
TensorFlow checkpoint bundles ("V2" checkpoints: `<prefix>.index` + `<prefix>.data-00000-of-00001`), read and written
without TensorFlow -- the format `tf.train.Saver` produces at main.py:192-206 of the reference, so weights trained there
can be loaded here and vice versa.

Implemented from the published formats, nothing else available in this image:
  * `.index` is a LevelDB-style sorted string table (tensorflow/core/lib/io/table*.cc, a port of leveldb/table):
    data blocks of prefix-compressed (key, value) entries with a restart array, each block followed by a 5-byte trailer
    (compression type, masked CRC32C), an index block of block handles, and a 48-byte footer ending in the magic
    0xdb4775248b80fb57;
  * keys are variable names; the value of key "" is a BundleHeaderProto, every other value a BundleEntryProto
    (tensorflow/core/protobuf/tensor_bundle.proto): dtype, shape, shard, offset and size into the data file, masked
    CRC32C of the bytes;
  * the data shard is the raw little-endian tensor contents back to back.
Only what a dense float checkpoint needs is handled: uncompressed or snappy-compressed blocks on read, one data shard,
no tensor slices.  **Unverified against TensorFlow's own reader/writer** (TensorFlow cannot be installed here): the tests
pin the primitives to published known answers (CRC32C check value, LevelDB footer magic, varints) and the reader to the
writer; DESIGN.md lists this as parity-unpinned.
"""
import os
import struct

import numpy as np

MAGIC = 0xdb4775248b80fb57
_MASK_DELTA = 0xa282ead8

# tensorflow/core/framework/types.proto
DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_,
          17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
_DT_OF = {np.dtype(v): k for k, v in DTYPES.items()}


# ------------------------------------------------------------------ CRC32C (Castagnoli), LevelDB masking
def _make_table():
    tab = []
    for n in range(256):
        c = n
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        tab.append(c)
    return tab


_TAB = _make_table()


def _crc32c_py(data, crc=0):
    crc ^= 0xFFFFFFFF
    for b in bytes(data):
        crc = _TAB[(crc ^ b) & 0xFF] ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def crc32c(data, crc=0):
    """CRC32C of a bytes-like object; through libdanet_sm100.so's host helper when the library is built (a 36 MB
    checkpoint takes ~40 ms instead of ~30 s), else in pure Python"""
    data = bytes(data) if not isinstance(data, (bytes, bytearray)) else data
    if len(data) >= 256:
        try:
            from . import _lib
            return int(_lib.load().danet_crc32c(bytes(data), len(data), crc))
        except Exception:
            pass
    return _crc32c_py(data, crc)


def mask_crc(crc):
    return (((crc >> 15) | (crc << 17)) + _MASK_DELTA) & 0xFFFFFFFF


def unmask_crc(m):
    rot = (m - _MASK_DELTA) & 0xFFFFFFFF
    return ((rot >> 17) | (rot << 15)) & 0xFFFFFFFF


# ------------------------------------------------------------------ varints / minimal protobuf
def put_varint(n):
    out = bytearray()
    n &= (1 << 64) - 1
    while n >= 0x80:
        out.append((n & 0x7F) | 0x80)
        n >>= 7
    out.append(n)
    return bytes(out)


def get_varint(buf, pos):
    shift, val = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        val |= (b & 0x7F) << shift
        if b < 0x80:
            return val, pos
        shift += 7


def _pb_fields(buf):
    """yield (field number, wire type, value) of one protobuf message"""
    pos, n = 0, len(buf)
    while pos < n:
        tag, pos = get_varint(buf, pos)
        f, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = get_varint(buf, pos)
        elif wt == 1:
            v = buf[pos:pos + 8]
            pos += 8
        elif wt == 2:
            ln, pos = get_varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            v = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError('unsupported protobuf wire type %d' % wt)
        yield f, wt, v


def _pb_varint(field, v):
    return put_varint(field << 3) + put_varint(v)


def _pb_bytes(field, b):
    return put_varint((field << 3) | 2) + put_varint(len(b)) + b


def _encode_entry(dtype, shape, offset, size, crc):
    dims = b''.join(_pb_bytes(2, _pb_varint(1, int(d))) for d in shape)          # TensorShapeProto.dim{size}
    msg = _pb_varint(1, _DT_OF[np.dtype(dtype)]) + _pb_bytes(2, dims)
    if offset:
        msg += _pb_varint(4, offset)
    msg += _pb_varint(5, size) + put_varint((6 << 3) | 5) + struct.pack('<I', crc)
    return msg


def _decode_entry(buf):
    e = dict(dtype=0, shape=[], shard_id=0, offset=0, size=0, crc32c=None, slices=False)
    for f, wt, v in _pb_fields(buf):
        if f == 1:
            e['dtype'] = v
        elif f == 2:
            for f2, _, v2 in _pb_fields(v):
                if f2 == 2:
                    size = 0
                    for f3, _, v3 in _pb_fields(v2):
                        if f3 == 1:
                            size = v3
                    e['shape'].append(size)
        elif f == 3:
            e['shard_id'] = v
        elif f == 4:
            e['offset'] = v
        elif f == 5:
            e['size'] = v
        elif f == 6:
            e['crc32c'] = struct.unpack('<I', v)[0]
        elif f == 7:
            e['slices'] = True
    return e


# ------------------------------------------------------------------ snappy (read side only; block format 1.1)
def _snappy_uncompress(buf):
    n, pos = get_varint(buf, 0)
    out = bytearray()
    while pos < len(buf):
        tag = buf[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:
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
            off = buf[pos] | (buf[pos + 1] << 8)
            pos += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 4], 'little')
            pos += 4
        for _ in range(ln):
            out.append(out[-off])
    if len(out) != n:
        raise ValueError('snappy: length mismatch')
    return bytes(out)


# ------------------------------------------------------------------ table blocks
def _read_block(f, offset, size, verify=True):
    f.seek(offset)
    raw = f.read(size + 5)
    if len(raw) != size + 5:
        raise IOError('truncated table block at %d' % offset)
    body, ctype, crc = raw[:size], raw[size], struct.unpack('<I', raw[size + 1:])[0]
    if verify and unmask_crc(crc) != crc32c(raw[:size + 1]):
        raise IOError('table block checksum mismatch at %d' % offset)
    if ctype == 1:
        body = _snappy_uncompress(body)
    elif ctype != 0:
        raise IOError('unknown block compression %d' % ctype)
    return body


def _block_entries(block):
    n_restarts = struct.unpack('<I', block[-4:])[0]
    end = len(block) - 4 - 4 * n_restarts
    pos, key = 0, b''
    while pos < end:
        shared, pos = get_varint(block, pos)
        non_shared, pos = get_varint(block, pos)
        vlen, pos = get_varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        yield key, block[pos:pos + vlen]
        pos += vlen


def _build_block(items, restart_interval=16):
    out, restarts, last = bytearray(), [], b''
    for i, (k, v) in enumerate(items):
        if i % restart_interval == 0:
            restarts.append(len(out))
            shared = 0
        else:
            shared = 0
            while shared < min(len(k), len(last)) and k[shared] == last[shared]:
                shared += 1
        out += put_varint(shared) + put_varint(len(k) - shared) + put_varint(len(v)) + k[shared:] + v
        last = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack('<I', r)
    out += struct.pack('<I', len(restarts))
    return bytes(out)


def _emit_block(f, body):
    off = f.tell()
    f.write(body)
    f.write(b'\x00')
    f.write(struct.pack('<I', mask_crc(crc32c(body + b'\x00'))))
    return off, len(body)


def _handle(off, size):
    return put_varint(off) + put_varint(size)


# ------------------------------------------------------------------ public API
def read_bundle(prefix, verify=True):
    """`prefix` as given to tf.train.Saver.save (no extension) -> {variable name: numpy array}"""
    index_path = prefix + '.index'
    if not os.path.exists(index_path):
        raise IOError('no such checkpoint: %s' % index_path)
    with open(index_path, 'rb') as f:
        f.seek(0, 2)
        total = f.tell()
        if total < 48:
            raise IOError('%s is too short to be a table' % index_path)
        f.seek(total - 48)
        footer = f.read(48)
        if struct.unpack('<Q', footer[40:])[0] != MAGIC:
            raise IOError('%s: bad table magic' % index_path)
        pos = 0
        _, pos = get_varint(footer, pos)          # metaindex handle
        _, pos = get_varint(footer, pos)
        ioff, pos = get_varint(footer, pos)
        isize, pos = get_varint(footer, pos)
        entries = {}
        for _, hv in _block_entries(_read_block(f, ioff, isize, verify)):
            boff, p2 = get_varint(hv, 0)
            bsize, _ = get_varint(hv, p2)
            for k, v in _block_entries(_read_block(f, boff, bsize, verify)):
                entries[k] = v
    header = entries.pop(b'', None)
    num_shards = 1
    if header is not None:
        for fnum, _, v in _pb_fields(header):
            if fnum == 1:
                num_shards = v
            elif fnum == 2 and v != 0:
                raise IOError('big-endian bundles are not supported')
    shards = {}
    out = {}
    for k, v in entries.items():
        e = _decode_entry(v)
        if e['slices']:
            raise IOError('sliced (partitioned) variable %r is not supported' % k)
        if e['dtype'] not in DTYPES:
            raise IOError('variable %r has unsupported dtype %d' % (k, e['dtype']))
        sid = e['shard_id']
        if sid not in shards:
            shards[sid] = np.fromfile('%s.data-%05d-of-%05d' % (prefix, sid, num_shards), dtype=np.uint8)
        raw = shards[sid][e['offset']:e['offset'] + e['size']]
        if len(raw) != e['size']:
            raise IOError('variable %r runs past the end of its data shard' % k)
        if verify and e['crc32c'] is not None and e['size'] <= (4 << 20):
            if unmask_crc(e['crc32c']) != crc32c(raw.tobytes()):
                raise IOError('variable %r: data checksum mismatch' % k)
        out[k.decode()] = raw.view(DTYPES[e['dtype']]).reshape(e['shape']).copy()
    return out


def write_bundle(prefix, tensors, entries_per_block=64):
    """{variable name: array} -> `<prefix>.index` + `<prefix>.data-00000-of-00001` (one shard, uncompressed blocks,
    every tensor with its masked CRC32C as TensorFlow's reader checks it on restore)"""
    names = sorted(tensors)
    d = os.path.dirname(os.path.abspath(prefix))
    if not os.path.exists(d):
        os.makedirs(d)                                       # main.py:194-196
    data_path = '%s.data-00000-of-00001' % prefix
    items = []
    offset = 0
    with open(data_path, 'wb') as df:
        for name in names:
            a = np.asarray(tensors[name])
            a = a.copy(order='C') if not a.flags['C_CONTIGUOUS'] else a
            if np.dtype(a.dtype) not in _DT_OF:
                raise ValueError('variable %r has unsupported dtype %s' % (name, a.dtype))
            raw = a.tobytes()
            df.write(raw)
            items.append((name.encode(), _encode_entry(a.dtype, a.shape, offset, len(raw), mask_crc(crc32c(raw)))))
            offset += len(raw)
    header = _pb_varint(1, 1) + _pb_bytes(3, _pb_varint(1, 1))       # num_shards = 1, little endian, version.producer = 1
    items = [(b'', header)] + items
    with open(prefix + '.index', 'wb') as f:
        index = []
        for i in range(0, len(items), entries_per_block):
            chunk = items[i:i + entries_per_block]
            boff, bsize = _emit_block(f, _build_block(chunk))
            index.append((chunk[-1][0], _handle(boff, bsize)))       # separator key = last key of the block
        moff, msize = _emit_block(f, _build_block([]))
        ioff, isize = _emit_block(f, _build_block(index, restart_interval=1))
        footer = _handle(moff, msize) + _handle(ioff, isize)
        f.write(footer + b'\x00' * (40 - len(footer)) + struct.pack('<Q', MAGIC))
    return prefix
