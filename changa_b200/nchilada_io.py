"""N-Chilada snapshot directories: the reference's native format next to Tipsy.

Layout (TreePiece::loadNChilada, InOutput.cpp:843-956; load_NC_base / _dark / _gas / _star,
InOutput.cpp:507-840; writers InOutput.cpp:1997-2330): one directory per snapshot, one sub-directory per
particle family (`gas`, `dark`, `star`), one FILE PER ATTRIBUTE (`pos`, `vel`, `mass`, `soft`, `pot`,
... ; ChaNGa's `-n 0` force dump adds `acc2`, AccOutputParams, InOutput.h:386-405).  Every attribute
file is XDR (big-endian):

    FieldHeader  28 bytes   int magic = 1062053; double time; uint high word of the count;
                            uint numParticles (low word); uint dimensions (1 | 3); int type code
    minimum, maximum        `dimensions` values each, of the field's type
    data                    numParticles x dimensions values, particle-major

Type codes (utility/structures tree_xdr.h): int8 1, uint8 2, int16 3, uint16 4, int32 5, uint32 6,
int64 7, uint64 8, float32 9, float64 10.  The gravity path reads pos / mass / soft of every family in
the order gas, dark, star (the particle order of the merged array, InOutput.cpp:900-938)."""
import os
import struct

import numpy as np

MAGIC = 1062053
HEADER_BYTES = 28
FAMILIES = ("gas", "dark", "star")
_DTYPES = {1: ">i1", 2: ">u1", 3: ">i2", 4: ">u2", 5: ">i4", 6: ">u4", 7: ">i8", 8: ">u8", 9: ">f4", 10: ">f8"}
_CODES = {np.dtype(v).newbyteorder("=").name: k for k, v in _DTYPES.items()}


def read_field(path, start=0, count=None):
    """-> (array (N,) or (N,3) in native byte order, header dict); start / count select a particle range
    as the reference's readFieldData does for one TreePiece's share (InOutput.cpp:474-503)"""
    with open(path, "rb") as f:
        raw = f.read(HEADER_BYTES)
        if len(raw) < HEADER_BYTES:
            raise ValueError(f"{path}: shorter than a field header")
        magic, time, high, low, dims, code = struct.unpack(">idIIIi", raw)
        if magic != MAGIC:
            raise ValueError(f"{path}: not a field file (magic number doesn't match)")  # InOutput.cpp:462-464
        if dims not in (1, 3):
            raise ValueError(f"{path}: wrong dimension {dims}")                           # InOutput.cpp:465-467
        if code not in _DTYPES:
            raise ValueError(f"{path}: unknown type code {code}")
        n = (high << 32) | low
        dt = np.dtype(_DTYPES[code])
        lo = np.frombuffer(f.read(dims * dt.itemsize), dtype=dt)
        hi = np.frombuffer(f.read(dims * dt.itemsize), dtype=dt)
        count = n - start if count is None else count
        if start < 0 or count < 0 or start + count > n:
            raise ValueError(f"{path}: particles [{start}, {start + count}) outside the file's {n}")
        f.seek(start * dims * dt.itemsize, 1)
        data = np.frombuffer(f.read(count * dims * dt.itemsize), dtype=dt)
        if len(data) != count * dims:
            raise ValueError(f"{path}: truncated")
    out = data.astype(dt.newbyteorder("="))
    hdr = {"time": time, "numParticles": n, "dimensions": dims, "code": code,
           "min": lo.astype(dt.newbyteorder("=")), "max": hi.astype(dt.newbyteorder("="))}
    return (out.reshape(count, 3) if dims == 3 else out), hdr


def write_field(path, values, time=0.0, dtype=np.float32):
    """one attribute file; values (N,) or (N,3)"""
    v = np.ascontiguousarray(values, dtype=dtype)
    dims = 3 if v.ndim == 2 else 1
    assert v.ndim == 1 or v.shape[1] == 3
    n = v.shape[0]
    code = _CODES[np.dtype(dtype).name]
    be = np.dtype(_DTYPES[code])
    lo = v.min(axis=0) if n else np.zeros(dims, dtype=dtype)
    hi = v.max(axis=0) if n else np.zeros(dims, dtype=dtype)
    with open(path, "wb") as f:
        f.write(struct.pack(">idIIIi", MAGIC, float(time), n >> 32, n & 0xffffffff, dims, code))
        f.write(np.atleast_1d(lo).astype(be).tobytes())
        f.write(np.atleast_1d(hi).astype(be).tobytes())
        f.write(v.astype(be).tobytes())


def family_counts(dirname):
    """particles per family, from <family>/pos (ncGetCount, InOutput.cpp:447-472; a missing family counts 0)"""
    out = {}
    for fam in FAMILIES:
        p = os.path.join(dirname, fam, "pos")
        if not os.path.exists(p):
            out[fam] = 0
            continue
        with open(p, "rb") as f:
            magic, _, high, low, dims, _ = struct.unpack(">idIIIi", f.read(HEADER_BYTES))
        if magic != MAGIC:
            raise ValueError(f"{p}: not a field file (magic number doesn't match)")
        out[fam] = (high << 32) | low
    return out


def read_nchilada(dirname):
    """(pos (N,3), mass (N,), soft (N,), counts) of every particle, families in the order gas, dark, star --
    what the gravity path needs of a snapshot"""
    counts = family_counts(dirname)
    if sum(counts.values()) == 0:
        raise ValueError(f"{dirname}: no gas/pos, dark/pos or star/pos")
    pos, mass, soft = [], [], []
    for fam in FAMILIES:
        if counts[fam] == 0:
            continue
        d = os.path.join(dirname, fam)
        p, _ = read_field(os.path.join(d, "pos"))
        m, _ = read_field(os.path.join(d, "mass"))
        s, _ = read_field(os.path.join(d, "soft"))
        if not (len(p) == len(m) == len(s) == counts[fam]):
            raise ValueError(f"{d}: pos / mass / soft disagree on the particle count")
        pos.append(p.astype(np.float64)); mass.append(m.astype(np.float64)); soft.append(s.astype(np.float64))
    return np.concatenate(pos), np.concatenate(mass), np.concatenate(soft), counts


def write_nchilada(dirname, pos, mass, soft, family="dark", time=0.0, extra=None):
    """a snapshot with one family; extra: {attribute name: array} written beside pos / mass / soft (e.g. the
    `acc2` and `pot` arrays of a force step, in the snapshot's particle order)"""
    d = os.path.join(dirname, family)
    os.makedirs(d, exist_ok=True)
    write_field(os.path.join(d, "pos"), pos, time)
    write_field(os.path.join(d, "mass"), mass, time)
    write_field(os.path.join(d, "soft"), soft, time)
    for name, arr in (extra or {}).items():
        write_field(os.path.join(d, name), arr, time)
