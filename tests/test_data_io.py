"""On-disk formats either side of the path (PFM height maps, 170-value RPC text): our readers /
writers against files and arrays produced by the reference's own functions
(`oracle/make_golden_io.py` exec-extracts dataset/data_io.py:17-92 and writes tests/golden/io_*)."""
import os

import numpy as np
import pytest

from satmvs_b200 import data_io

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def ref():
    return np.load(os.path.join(G, "io_formats.npz"))


def test_load_pfm_matches_reference_reader(ref):
    for name, key in (("io_grey.pfm", "grey_loaded"), ("io_color.pfm", "color_loaded"), ("io_big_endian.pfm", "big_loaded")):
        got = data_io.load_pfm(os.path.join(G, name))
        assert got.shape == ref[key].shape
        assert np.array_equal(np.asarray(got, dtype=np.float32), ref[key].astype(np.float32))
    assert np.array_equal(data_io.load_pfm(os.path.join(G, "io_grey.pfm")), ref["grey"])


def test_save_pfm_is_byte_identical_to_reference_writer(ref, tmp_path):
    for name, key in (("io_grey.pfm", "grey"), ("io_color.pfm", "color")):
        out = tmp_path / name
        data_io.save_pfm(str(out), ref[key])
        assert out.read_bytes() == open(os.path.join(G, name), "rb").read()
    one = tmp_path / "one.pfm"                        # H x W x 1 is written as greyscale
    data_io.save_pfm(str(one), ref["grey"][:, :, None])
    assert np.array_equal(data_io.load_pfm(str(one)), ref["grey"])


def test_pfm_errors_like_the_reference(tmp_path):
    with pytest.raises(Exception, match="float32"):
        data_io.save_pfm(str(tmp_path / "x.pfm"), np.zeros((2, 2), np.float64))
    with pytest.raises(Exception, match="dimensions"):
        data_io.save_pfm(str(tmp_path / "x.pfm"), np.zeros((2, 2, 2), np.float32))
    bad = tmp_path / "bad.pfm"
    bad.write_bytes(b"P6\n2 2\n-1.0\n")
    with pytest.raises(Exception, match="Not a PFM"):
        data_io.load_pfm(str(bad))
    bad.write_bytes(b"Pf\n2 x\n-1.0\n")
    with pytest.raises(Exception, match="Malformed"):
        data_io.load_pfm(str(bad))


def test_load_rpc_as_array(ref):
    data, h_max, h_min = data_io.load_rpc_as_array(os.path.join(G, "io_rpc.rpc"))
    assert data.dtype == np.float64 and data.shape == (170,)
    assert np.array_equal(data, ref["rpc"])
    assert h_max == float(ref["h_max"]) and h_min == float(ref["h_min"])
    with pytest.raises(Exception, match="RPC not found"):
        data_io.load_rpc_as_array(os.path.join(G, "missing.rpc"))


def test_rpc_file_feeds_the_camera_pack(ref):
    """The loaded 170-vector is what the sweep's host-side camera pack consumes (layout of SURVEY.md §8 a5)."""
    data, h_max, h_min = data_io.load_rpc_as_array(os.path.join(G, "io_rpc.rpc"))
    assert h_max - h_min == pytest.approx(2 * data[9])
    assert (h_max + h_min) / 2 == pytest.approx(data[4])
