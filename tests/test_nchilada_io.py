"""N-Chilada attribute files and snapshot directories (InOutput.cpp:445-956, 1997-2330)"""
import struct

import numpy as np
import pytest

from changa_b200 import nchilada_io as nc


def test_field_file_layout_is_the_xdr_field_header(tmp_path):
    """28-byte big-endian header {magic 1062053, double time, high word, count, dimensions, type code},
    then minimum and maximum, then particle-major data"""
    pos = np.array([[1.0, -2.0, 3.0], [0.5, 4.0, -6.0]], dtype=np.float32)
    p = tmp_path / "pos"
    nc.write_field(p, pos, time=0.25)
    raw = open(p, "rb").read()
    assert len(raw) == 28 + 2 * 12 + 2 * 12
    assert struct.unpack(">idIIIi", raw[:28]) == (1062053, 0.25, 0, 2, 3, 9)
    assert struct.unpack(">3f", raw[28:40]) == (0.5, -2.0, -6.0) and struct.unpack(">3f", raw[40:52]) == (1.0, 4.0, 3.0)
    assert struct.unpack(">6f", raw[52:]) == (1.0, -2.0, 3.0, 0.5, 4.0, -6.0)
    back, hdr = nc.read_field(p)
    assert np.array_equal(back, pos) and hdr["numParticles"] == 2 and hdr["dimensions"] == 3 and hdr["time"] == 0.25


def test_partial_reads_and_other_types(tmp_path):
    v = np.arange(100, dtype=np.float64) * 0.5
    nc.write_field(tmp_path / "mass", v, dtype=np.float64)
    part, hdr = nc.read_field(tmp_path / "mass", start=10, count=5)   # one TreePiece's share
    assert hdr["code"] == 10 and np.array_equal(part, v[10:15])
    ids = np.arange(7, dtype=np.int64) * 2 ** 33
    nc.write_field(tmp_path / "iord", ids, dtype=np.int64)
    assert np.array_equal(nc.read_field(tmp_path / "iord")[0], ids)
    with pytest.raises(ValueError):
        nc.read_field(tmp_path / "mass", start=99, count=5)


def test_bad_files_are_refused(tmp_path):
    (tmp_path / "junk").write_bytes(b"\0" * 64)
    with pytest.raises(ValueError, match="magic"):
        nc.read_field(tmp_path / "junk")
    (tmp_path / "dim2").write_bytes(struct.pack(">idIIIi", nc.MAGIC, 0.0, 0, 1, 2, 9) + b"\0" * 64)
    with pytest.raises(ValueError, match="dimension"):
        nc.read_field(tmp_path / "dim2")
    (tmp_path / "short").write_bytes(struct.pack(">idIIIi", nc.MAGIC, 0.0, 0, 50, 1, 9) + b"\0" * 20)
    with pytest.raises(ValueError):
        nc.read_field(tmp_path / "short")
    with pytest.raises(ValueError):
        nc.read_nchilada(str(tmp_path))          # no family directories


def test_snapshot_round_trip_feeds_the_tree(tmp_path):
    """families come back in the order gas, dark, star; a force step's accelerations are written beside them"""
    rng = np.random.default_rng(5)
    d = str(tmp_path / "snap.000000")
    for fam, n in (("dark", 300), ("gas", 120), ("star", 40)):
        nc.write_nchilada(d, rng.uniform(-0.5, 0.5, (n, 3)), rng.uniform(1, 2, n) / 460, np.full(n, 0.01), family=fam, time=1.5)
    pos, mass, soft, counts = nc.read_nchilada(d)
    assert counts == {"gas": 120, "dark": 300, "star": 40} and pos.shape == (460, 3)
    first_gas, _ = nc.read_field(tmp_path / "snap.000000" / "gas" / "pos")
    assert np.array_equal(pos[:120], first_gas.astype(np.float64))
    from changa_b200.tree import Tree
    t = Tree(pos, mass, soft, max_bucket=12)
    assert t.n == 460 and np.isclose(t.moments[0][2], mass.sum())
    acc = rng.normal(size=(300, 3))
    nc.write_nchilada(d, pos[120:420], mass[120:420], soft[120:420], family="dark", time=1.5, extra={"acc2": acc})
    back, hdr = nc.read_field(tmp_path / "snap.000000" / "dark" / "acc2")
    assert hdr["dimensions"] == 3 and np.array_equal(back, acc.astype(np.float32))
