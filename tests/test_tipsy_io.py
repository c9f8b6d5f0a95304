"""Tipsy snapshot / array files: round trips and the reference's own fixtures"""
import os

import numpy as np
import pytest

from changa_b200 import tipsy_io
from changa_b200.workloads import read_tipsy


def test_snapshot_round_trip(tmp_path):
    rng = np.random.default_rng(3)
    pos, mass, soft = rng.uniform(-0.5, 0.5, (257, 3)), rng.uniform(1, 2, 257), np.full(257, 0.01)
    p = tmp_path / "box.tbin"
    tipsy_io.write_tipsy(p, pos, mass, soft)
    pos2, mass2, soft2 = read_tipsy(str(p))
    f32 = lambda a: np.asarray(a, dtype=np.float32).astype(np.float64)
    assert np.array_equal(pos2, f32(pos)) and np.array_equal(mass2, f32(mass)) and np.array_equal(soft2, f32(soft))


@pytest.mark.parametrize("binary", [False, True])
def test_acc2_round_trip(tmp_path, binary):
    rng = np.random.default_rng(4)
    acc = rng.normal(size=(100, 3)) * 1e3
    p = tmp_path / "box.acc2"
    tipsy_io.write_array(p, acc, binary=binary)
    back = tipsy_io.read_array(p)
    assert back.shape == (100, 3)
    if binary:
        assert np.array_equal(back, acc.astype(np.float32).astype(np.float64))
    else:
        np.testing.assert_allclose(back, acc, rtol=1e-13)
        lines = open(p).read().split("\n")
        assert lines[0] == "100" and lines[1] == "%.14g" % acc[0, 0] and lines[101] == "%.14g" % acc[0, 1]
    pot = rng.normal(size=100)
    tipsy_io.write_array(p, pot, binary=binary)
    assert tipsy_io.read_array(p).shape == (100,)


def test_file_order():
    order = np.array([2, 0, 3, 1])
    acc = np.arange(12.0).reshape(4, 3)
    out = tipsy_io.accelerations_in_file_order(acc, order)
    assert np.array_equal(out[2], acc[0]) and np.array_equal(out[1], acc[3])


@pytest.mark.skipif(not os.path.exists("/root/reference/teststep/king_soft.bin"), reason="reference tree not present")
def test_reads_the_reference_fixtures():
    pos, mass, soft = read_tipsy("/root/reference/teststep/king_soft.bin")
    assert len(pos) == 36000 and np.isfinite(pos).all() and (mass > 0).all() and (soft > 0).all()
    pos, mass, soft = read_tipsy("/root/reference/testcosmo/cube300.tbin")
    assert len(pos) == 48 ** 3 and np.abs(pos).max() < 0.52  # a drifted periodic unit box
