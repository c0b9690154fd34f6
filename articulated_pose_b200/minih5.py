"""minih5 -- a minimal HDF5 writer / reader for the per-cloud prediction files (lib/prediction_io.py:65-95).

h5py is not part of this image, and the reference's evaluation scripts open `<basename>.h5` with `h5py.File(...)[name][()]`.
This module writes files in the HDF5 file format (specification version 3.0, the structures libhdf5 >= 1.8 reads) with
exactly what those files need and nothing else:

  * superblock version 2 (48 bytes, 8-byte offsets / lengths, lookup3 checksum);
  * a root group as ONE version-2 object header ("OHDR") holding a Link Info message (compact storage: no fractal heap),
    a Group Info message, one hard Link message per dataset and one Attribute message per (string) attribute;
  * one version-2 object header per dataset: Dataspace (v2, simple), Datatype (v1: IEEE little-endian f32 / f64, two's
    complement i8..i64 / u8..u64), Fill Value (v3, none), Data Layout (v3, contiguous);
  * the raw data, C order, little endian, 8-byte aligned.
No chunking, compression, nested groups, variable-length types or dense storage -- `save_batch_nn` uses none of them.
Attributes are fixed-length UTF-8 strings.

`read(path)` parses the same subset (it is what `prediction_io.load_prediction` falls back to without h5py) and verifies
every checksum; files written by other producers with older structures (version-0 superblock, symbol-table groups) are
rejected with a clear error -- open those with h5py.

There is no libhdf5 in the build image to cross-check against: the layout follows the published format specification, the
lookup3 checksum is pinned by the known-answer vectors of Bob Jenkins' lookup3.c (tests/test_minih5_cpu.py), and
write -> read round trips are exact.
"""
import struct

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF

MSG_DATASPACE, MSG_LINK_INFO, MSG_DATATYPE, MSG_FILL_VALUE, MSG_LINK, MSG_LAYOUT, MSG_GROUP_INFO, MSG_ATTRIBUTE = \
    0x01, 0x02, 0x03, 0x05, 0x06, 0x08, 0x0A, 0x0C


def _rot(x, k):
    return ((x << k) | (x >> (32 - k))) & 0xFFFFFFFF


def lookup3(data, initval=0):
    """Bob Jenkins' lookup3 hashlittle() -- HDF5's metadata checksum (H5_checksum_lookup3)."""
    data = bytes(data)
    length = len(data)
    a = b = c = (0xDEADBEEF + length + initval) & 0xFFFFFFFF
    k = 0
    M = 0xFFFFFFFF
    while length > 12:
        a = (a + int.from_bytes(data[k:k + 4], "little")) & M
        b = (b + int.from_bytes(data[k + 4:k + 8], "little")) & M
        c = (c + int.from_bytes(data[k + 8:k + 12], "little")) & M
        a = (a - c) & M; a ^= _rot(c, 4); c = (c + b) & M
        b = (b - a) & M; b ^= _rot(a, 6); a = (a + c) & M
        c = (c - b) & M; c ^= _rot(b, 8); b = (b + a) & M
        a = (a - c) & M; a ^= _rot(c, 16); c = (c + b) & M
        b = (b - a) & M; b ^= _rot(a, 19); a = (a + c) & M
        c = (c - b) & M; c ^= _rot(b, 4); b = (b + a) & M
        k += 12
        length -= 12
    if length == 0:
        return c
    tail = data[k:] + b"\x00" * (12 - length)
    a = (a + int.from_bytes(tail[0:4], "little")) & M
    b = (b + int.from_bytes(tail[4:8], "little")) & M
    c = (c + int.from_bytes(tail[8:12], "little")) & M
    c ^= b; c = (c - _rot(b, 14)) & M
    a ^= c; a = (a - _rot(c, 11)) & M
    b ^= a; b = (b - _rot(a, 25)) & M
    c ^= b; c = (c - _rot(b, 16)) & M
    a ^= c; a = (a - _rot(c, 4)) & M
    b ^= a; b = (b - _rot(a, 14)) & M
    c ^= b; c = (c - _rot(b, 24)) & M
    return c


# ---- messages ---------------------------------------------------------------------------------------------------
def _datatype(dt):
    dt = np.dtype(dt)
    if dt.kind == "f" and dt.itemsize in (4, 8):
        # class 1 (floating point), version 1; bit field: little endian, mantissa normalisation 2 (implied msb), sign location
        size = dt.itemsize
        sign = 8 * size - 1
        exp_size, mant_size, bias = (8, 23, 127) if size == 4 else (11, 52, 1023)
        head = bytes([0x11, 0x20, sign, 0x00]) + struct.pack("<I", size)
        return head + struct.pack("<HHBBBBI", 0, 8 * size, mant_size, exp_size, 0, mant_size, bias)
    if dt.kind in "iu" and dt.itemsize in (1, 2, 4, 8):
        # class 0 (fixed point), version 1; bit field bit 3 = signed
        head = bytes([0x10, 0x08 if dt.kind == "i" else 0x00, 0x00, 0x00]) + struct.pack("<I", dt.itemsize)
        return head + struct.pack("<HH", 0, 8 * dt.itemsize)
    raise TypeError("minih5: unsupported dtype %s" % dt)


def _string_type(nbytes):
    # class 3 (string), version 1; bit field: null terminated (0), character set UTF-8 (1 << 4)
    return bytes([0x13, 0x10, 0x00, 0x00]) + struct.pack("<I", nbytes)


def _dataspace(shape):
    if len(shape) == 0:
        return bytes([2, 0, 0, 0])                                   # version 2, rank 0, no max dims, scalar
    return bytes([2, len(shape), 0, 1]) + b"".join(struct.pack("<Q", int(d)) for d in shape)


def _message(mtype, body, flags=0):
    return struct.pack("<BHB", mtype, len(body), flags) + body


def _object_header(messages):
    """Version-2 object header, one chunk: 'OHDR', version 2, flags 0x02 (4-byte chunk-0 size; no times, no attribute
    phase-change values, no creation order), the messages, lookup3 checksum over everything before it."""
    body = b"".join(messages)
    head = b"OHDR" + bytes([2, 0x02]) + struct.pack("<I", len(body))
    blob = head + body
    return blob + struct.pack("<I", lookup3(blob))


def _align8(n):
    return (n + 7) & ~7


def write(path, datasets, attrs=None):
    """datasets: {name: array} (written in this order), attrs: {name: str} on the root group."""
    attrs = attrs or {}
    names = list(datasets)
    arrays = []
    for k in names:
        a = np.asarray(datasets[k])
        if a.dtype == np.bool_:
            a = a.astype(np.uint8)
        a = np.ascontiguousarray(a.astype(a.dtype.newbyteorder("<"), copy=False)).reshape(a.shape)   # (keeps 0-d arrays 0-d)
        _datatype(a.dtype)
        arrays.append(a)

    def dataset_header(a, addr):
        return _object_header([
            _message(MSG_DATASPACE, _dataspace(a.shape)),
            _message(MSG_DATATYPE, _datatype(a.dtype), flags=0x01),                # constant message
            _message(MSG_FILL_VALUE, bytes([3, 0x0A])),                            # v3: allocate late, write fill if set, none defined
            _message(MSG_LAYOUT, bytes([3, 1]) + struct.pack("<QQ", addr, a.nbytes)),   # v3, contiguous
        ])

    def root_header(addrs):
        msgs = [_message(MSG_LINK_INFO, bytes([0, 0]) + struct.pack("<QQ", UNDEF, UNDEF)),
                _message(MSG_GROUP_INFO, bytes([0, 0]))]
        for k, addr in zip(names, addrs):
            nm = k.encode("utf-8")
            if len(nm) > 255:
                raise ValueError("minih5: link name too long")
            # version 1, flags 0x10 (character-set field present, 1-byte name length, hard link), UTF-8
            msgs.append(_message(MSG_LINK, bytes([1, 0x10, 1, len(nm)]) + nm + struct.pack("<Q", addr)))
        for k, v in attrs.items():
            nm = k.encode("utf-8") + b"\x00"
            val = (v if isinstance(v, bytes) else str(v).encode("utf-8")) + b"\x00"
            dtm, dsm = _string_type(len(val)), _dataspace(())
            msgs.append(_message(MSG_ATTRIBUTE, struct.pack("<BBHHHB", 3, 0, len(nm), len(dtm), len(dsm), 1) + nm + dtm + dsm + val))
        return _object_header(msgs)

    # layout: superblock | root header | dataset headers | data (sizes do not depend on the addresses)
    root_addr = 48
    off = root_addr + len(root_header([0] * len(names)))
    hdr_addrs = []
    for a in arrays:
        hdr_addrs.append(off)
        off += len(dataset_header(a, 0))
    off = _align8(off)
    data_addrs = []
    for a in arrays:
        data_addrs.append(off if a.nbytes else UNDEF)
        off = _align8(off + a.nbytes)
    eof = off
    sb = SIGNATURE + bytes([2, 8, 8, 0]) + struct.pack("<QQQQ", 0, UNDEF, eof, root_addr)
    sb += struct.pack("<I", lookup3(sb))
    out = bytearray(eof)
    out[0:48] = sb
    blob = root_header(hdr_addrs)
    out[root_addr:root_addr + len(blob)] = blob
    for a, ha, da in zip(arrays, hdr_addrs, data_addrs):
        blob = dataset_header(a, da)
        out[ha:ha + len(blob)] = blob
        if a.nbytes:
            out[da:da + a.nbytes] = a.tobytes()
    with open(path, "wb") as f:
        f.write(out)


# ---- reader (same subset) ---------------------------------------------------------------------------------------
class FormatError(ValueError):
    pass


def _parse_header(buf, addr):
    if buf[addr:addr + 4] != b"OHDR" or buf[addr + 4] != 2:
        raise FormatError("minih5: not a version-2 object header at %d (file written with older HDF5 structures: use h5py)" % addr)
    flags = buf[addr + 5]
    p = addr + 6
    if flags & 0x20:
        p += 16
    if flags & 0x10:
        p += 4
    nsz = 1 << (flags & 3)
    size = int.from_bytes(buf[p:p + nsz], "little")
    p += nsz
    end = p + size
    if lookup3(buf[addr:end]) != struct.unpack_from("<I", buf, end)[0]:
        raise FormatError("minih5: object header checksum mismatch at %d" % addr)
    msgs = []
    while p + 4 <= end:
        mtype, msize, _mflags = struct.unpack_from("<BHB", buf, p)
        p += 4
        if flags & 0x04:
            p += 2
        msgs.append((mtype, bytes(buf[p:p + msize])))
        p += msize
    return msgs


def _parse_type(b):
    cls, ver = b[0] & 0x0F, b[0] >> 4
    size = struct.unpack_from("<I", b, 4)[0]
    if ver != 1:
        raise FormatError("minih5: datatype version %d" % ver)
    if cls == 1:
        return np.dtype("<f%d" % size)
    if cls == 0:
        return np.dtype("<%s%d" % ("i" if b[1] & 0x08 else "u", size))
    if cls == 3:
        return np.dtype("S%d" % size)
    raise FormatError("minih5: datatype class %d" % cls)


def _parse_space(b):
    if b[0] != 2:
        raise FormatError("minih5: dataspace version %d" % b[0])
    rank = b[1]
    return tuple(struct.unpack_from("<Q", b, 4 + 8 * i)[0] for i in range(rank))


def read(path):
    """-> (datasets {name: ndarray}, attrs {name: str}) of a file written by `write`."""
    with open(path, "rb") as f:
        buf = f.read()
    if buf[:8] != SIGNATURE:
        raise FormatError("minih5: not an HDF5 file")
    if buf[8] not in (2, 3) or buf[9] != 8 or buf[10] != 8:
        raise FormatError("minih5: superblock version %d (only the version-2 layout this module writes is supported: use h5py)" % buf[8])
    if lookup3(buf[:44]) != struct.unpack_from("<I", buf, 44)[0]:
        raise FormatError("minih5: superblock checksum mismatch")
    root = struct.unpack_from("<Q", buf, 36)[0]
    datasets, attrs = {}, {}
    for mtype, body in _parse_header(buf, root):
        if mtype == MSG_LINK:
            flags = body[1]
            p = 2
            if flags & 0x08:
                if body[p] != 0:
                    raise FormatError("minih5: only hard links")
                p += 1
            if flags & 0x04:
                p += 8
            if flags & 0x10:
                p += 1
            nsz = 1 << (flags & 3)
            n = int.from_bytes(body[p:p + nsz], "little")
            p += nsz
            name = body[p:p + n].decode("utf-8")
            addr = struct.unpack_from("<Q", body, p + n)[0]
            dt = shape = layout = None
            for t2, b2 in _parse_header(buf, addr):
                if t2 == MSG_DATATYPE:
                    dt = _parse_type(b2)
                elif t2 == MSG_DATASPACE:
                    shape = _parse_space(b2)
                elif t2 == MSG_LAYOUT:
                    if b2[0] != 3 or b2[1] != 1:
                        raise FormatError("minih5: only contiguous layouts")
                    layout = struct.unpack_from("<QQ", b2, 2)
            if dt is None or shape is None or layout is None:
                raise FormatError("minih5: incomplete dataset %r" % name)
            n_el = int(np.prod(shape)) if shape else 1
            datasets[name] = (np.frombuffer(buf, dt, n_el, layout[0]).reshape(shape).copy() if n_el else np.zeros(shape, dt))
        elif mtype == MSG_ATTRIBUTE:
            ver, _fl, nsz, dsz, ssz = struct.unpack_from("<BBHHH", body, 0)
            p = 9 if ver == 3 else 8
            name = body[p:p + nsz].rstrip(b"\x00").decode("utf-8")
            p += nsz
            dt = _parse_type(body[p:p + dsz])
            p += dsz + ssz
            attrs[name] = body[p:p + dt.itemsize].rstrip(b"\x00").decode("utf-8") if dt.kind == "S" else np.frombuffer(body, dt, 1, p)[0]
    return datasets, attrs
