"""Golden bytes for the on-disk formats, written by UPSTREAM's own functions
(datasets/data_io.py save_pfm; test_tank.py write_cam, compiled from the upstream file because the script
parses arguments at import) in the build container.

    python tests/golden/make_golden_io.py        # needs /root/reference; CPU only
"""
import ast
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
from datasets.data_io import read_pfm, save_pfm  # noqa: E402  (upstream)


def upstream_write_cam():
    tree = ast.parse(open("/root/reference/test_tank.py").read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "write_cam"]
    ns = {}
    exec(compile(ast.Module(body=fn, type_ignores=[]), "test_tank.py", "exec"), ns)
    return ns["write_cam"]


def main():
    g = np.random.default_rng(5)
    depth = (g.random((7, 9)) * 500 + 425).astype(np.float32)
    color = g.random((5, 6, 3)).astype(np.float32)
    cam = np.zeros((2, 4, 4), dtype=np.float32)
    cam[0] = np.eye(4, dtype=np.float32) + (g.random((4, 4)) * 0.1).astype(np.float32)
    cam[1, :3, :3] = np.array([[2892.33, 0, 823.2], [0, 2883.18, 619.07], [0, 0, 1]], dtype=np.float32)
    cam[1, 3, :2] = [425.0, 2.5]
    out = {"depth": depth, "color": color, "cam": cam}
    with tempfile.TemporaryDirectory() as d:
        for name, arr in (("depth", depth), ("color", color)):
            p = os.path.join(d, name + ".pfm")
            save_pfm(p, arr)
            out[name + "_pfm"] = np.frombuffer(open(p, "rb").read(), dtype=np.uint8)
            back, scale = read_pfm(p)
            assert np.array_equal(back, arr) and scale == 1.0
        p = os.path.join(d, "cam.txt")
        upstream_write_cam()(p, cam, 935.0, 425.0)
        out["cam_txt"] = np.frombuffer(open(p, "rb").read(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "io_formats.npz"), **out)
    print("wrote io_formats", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
