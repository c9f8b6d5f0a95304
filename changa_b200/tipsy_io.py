"""Tipsy I/O around the force path: the standard (big-endian XDR) snapshot the
reference's fixtures use (testdata/tipsydefs.h:4-110; reader in changa_b200.workloads.read_tipsy),
a writer for it, and the Tipsy ARRAY files ChaNGa writes accelerations into with `-n 0`
(`<snapshot>.acc2`, AccOutputParams, InOutput.h:386-405): ASCII = particle count, then every x,
every y, every z one value per line in "%.14g" (TreePiece::outputASCII, InOutput.cpp:1755-1870);
binary = XDR int count + big-endian floats in the same component-major order
(InOutput.cpp:2165-2175, 2297-2330).  With these a force step of this repository can be diffed
against a real ChaNGa run of the same snapshot."""
import struct

import numpy as np


def write_tipsy(path, pos, mass, soft, vel=None, time=0.0):
    """all particles as DARK (9 floats: mass, pos, vel, eps, phi), standard big-endian, padded header"""
    n = len(pos)
    vel = np.zeros((n, 3)) if vel is None else np.asarray(vel)
    rec = np.zeros((n, 9), dtype=">f4")
    rec[:, 0] = mass
    rec[:, 1:4] = pos
    rec[:, 4:7] = vel
    rec[:, 7] = soft
    with open(path, "wb") as f:
        f.write(struct.pack(">diiiii", float(time), n, 3, 0, n, 0))
        f.write(b"\0\0\0\0")  # the pad most writers add (sizeof(struct dump) = 32)
        f.write(rec.tobytes())


def write_array(path, values, binary=False):
    """values: (N,) scalar or (N,3) vector array in file particle order"""
    v = np.asarray(values, dtype=np.float64)
    n = v.shape[0]
    flat = v.T.reshape(-1) if v.ndim == 2 else v  # component-major
    if binary:
        with open(path, "wb") as f:
            f.write(struct.pack(">i", n))
            f.write(flat.astype(">f4").tobytes())
    else:
        with open(path, "w") as f:
            f.write("%d\n" % n)
            f.write("".join("%.14g\n" % x for x in flat))


def read_array(path, vector=None):
    """-> (N,) or (N,3); ASCII or XDR binary detected from the content, as the reference does
    (InOutput.cpp:360-372: a leading binary int equal to the particle count)"""
    raw = open(path, "rb").read()
    n_bin = struct.unpack(">i", raw[:4])[0] if len(raw) >= 4 else -1
    if n_bin > 0 and len(raw) - 4 in (4 * n_bin, 12 * n_bin):
        flat = np.frombuffer(raw, dtype=">f4", offset=4).astype(np.float64)
        n = n_bin
    else:
        tok = raw.split()
        n = int(tok[0])
        flat = np.array(tok[1:], dtype=np.float64)
    if vector is None:
        vector = len(flat) == 3 * n
    assert len(flat) == (3 * n if vector else n), "array length does not match the particle count"
    return flat.reshape(3, n).T.copy() if vector else flat


def accelerations_in_file_order(acc_sorted, order):
    """rows of a force step (tree order) -> the snapshot's particle order: out[order[i]] = acc[i]"""
    out = np.empty_like(np.asarray(acc_sorted))
    out[np.asarray(order)] = acc_sorted
    return out
