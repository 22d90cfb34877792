"""On-disk formats against bytes written by upstream's own functions
(tests/golden/make_golden_io.py)."""
import os

import numpy as np

import effimvs_b200  # noqa: F401
from effimvs_b200 import data_io

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "io_formats.npz")


def test_pfm_bytes_and_roundtrip(tmp_path):
    z = np.load(GOLDEN)
    for name in ("depth", "color"):
        p = str(tmp_path / (name + ".pfm"))
        data_io.save_pfm(p, z[name])
        assert open(p, "rb").read() == z[name + "_pfm"].tobytes()        # byte-identical to upstream's save_pfm
        back, scale = data_io.read_pfm(p)
        assert np.array_equal(back, z[name]) and scale == 1.0
    # a file written by upstream reads back identically
    p = str(tmp_path / "up.pfm")
    open(p, "wb").write(z["depth_pfm"].tobytes())
    assert np.array_equal(data_io.read_pfm(p)[0], z["depth"])


def test_cam_text(tmp_path):
    z = np.load(GOLDEN)
    p = str(tmp_path / "cam.txt")
    data_io.write_cam(p, z["cam"], 935.0, 425.0)
    assert open(p, "rb").read() == z["cam_txt"].tobytes()
    K, E, tail = data_io.read_cam_file(p)
    assert np.array_equal(K, z["cam"][1, :3, :3]) and np.array_equal(E, z["cam"][0])
    assert tail == [425.0, 2.5, 935.0, 425.0]


def test_ply_layout_and_roundtrip(tmp_path):
    g = np.random.default_rng(0)
    pts = g.random((11, 3)).astype(np.float32) * 100
    col = (g.random((11, 3)) * 255).astype(np.uint8)
    p = str(tmp_path / "cloud.ply")
    data_io.write_ply(p, pts, col)
    raw = open(p, "rb").read()
    head, body = raw.split(b"end_header\n")
    assert head.startswith(b"ply\nformat binary_little_endian 1.0\nelement vertex 11\nproperty float x\n")
    assert len(body) == 11 * 15
    back_p, back_c = data_io.read_ply(p)
    assert np.array_equal(back_p, pts) and np.array_equal(back_c, col)
