"""On-disk formats on either side of the hot path: PFM depth / confidence maps,
MVSNet camera text files and the fused point cloud as a binary PLY.  Pure host I/O, byte-compatible with
what upstream reads and writes:

    read_pfm / save_pfm     upstream datasets/data_io.py:61-126
    write_cam               upstream test_tank.py:176-193
    read_cam_file           upstream datasets/general_eval.py:60-80 and test_dtu_dypcd.py:120-131 (no /4, no interval logic)
    write_ply               upstream test_tank.py:551-570 / test_dtu_dypcd.py:339-350 (plyfile's binary_little_endian layout)
"""
from __future__ import annotations

import sys

import numpy as np


class PFMError(ValueError):
    """A file that is not a PFM image, or an array that PFM cannot hold."""


_PFM_KINDS = {b"PF": 3, b"Pf": 1}          # magic -> channels


def read_pfm(filename):
    """-> (data (H,W) or (H,W,3) float32, top row first; scale).  PFM: three text lines (magic, "width height", scale whose
    sign gives the byte order) followed by the rows bottom-up."""
    with open(filename, "rb") as f:
        magic = f.readline().strip()
        if magic not in _PFM_KINDS:
            raise PFMError("{}: magic {!r} is neither PF nor Pf".format(filename, magic))
        dims = f.readline().split()
        if len(dims) != 2 or not all(d.isdigit() for d in dims):
            raise PFMError("{}: expected 'width height' on the second line, got {!r}".format(filename, b" ".join(dims)))
        width, height = int(dims[0]), int(dims[1])
        scale = float(f.readline().strip())
        data = np.fromfile(f, dtype="<f4" if scale < 0 else ">f4")
    channels = _PFM_KINDS[magic]
    shape = (height, width, 3) if channels == 3 else (height, width)
    if data.size != height * width * channels:
        raise PFMError("{}: {} values for a {} x {} x {} image".format(filename, data.size, height, width, channels))
    return data.reshape(shape)[::-1], abs(scale)


def save_pfm(filename, image, scale=1):
    """float32 (H,W), (H,W,1) or (H,W,3) -> PFM, rows written bottom-up, the scale line signed by the array's byte order."""
    image = np.asarray(image)
    if image.dtype != np.float32:
        raise PFMError("PFM stores float32, got {}".format(image.dtype))
    if image.ndim == 2 or (image.ndim == 3 and image.shape[2] == 1):
        magic = b"Pf"
    elif image.ndim == 3 and image.shape[2] == 3:
        magic = b"PF"
    else:
        raise PFMError("PFM holds (H,W), (H,W,1) or (H,W,3), got shape {}".format(image.shape))
    little = image.dtype.byteorder == "<" or (image.dtype.byteorder == "=" and sys.byteorder == "little")
    with open(filename, "wb") as f:
        f.write(magic + b"\n")
        f.write("{} {}\n".format(image.shape[1], image.shape[0]).encode("ascii"))
        f.write(("%f\n" % (-scale if little else scale)).encode("ascii"))
        image[::-1].tofile(f)


def write_cam(filename, cam, depth_max, depth_min):
    """cam (2,4,4): [0] extrinsic, [1][:3,:3] intrinsic, [1][3][:2] the two trailing numbers upstream carries there.
    MVSNet camera text: every matrix entry followed by one blank, a blank line between the blocks."""
    def rows(m, n):
        return "".join("".join(str(m[i][j]) + " " for j in range(n)) + "\n" for i in range(n))   # str(): numpy's shortest repr
    text = "extrinsic\n" + rows(cam[0], 4) + "\nintrinsic\n" + rows(cam[1], 3)
    text += "\n" + " ".join(str(t) for t in (cam[1][3][0], cam[1][3][1], depth_max, depth_min)) + "\n"
    with open(filename, "w") as f:
        f.write(text)


def read_cam_file(filename):
    """-> (intrinsics (3,3) float32, extrinsics (4,4) float32, last-line numbers as floats)."""
    with open(filename) as f:
        lines = [line.rstrip() for line in f.readlines()]
    extrinsics = np.array(" ".join(lines[1:5]).split(), dtype=np.float32).reshape(4, 4)
    intrinsics = np.array(" ".join(lines[7:10]).split(), dtype=np.float32).reshape(3, 3)
    tail = [float(t) for t in lines[11].split()] if len(lines) > 11 else []
    return intrinsics, extrinsics, tail


def write_ply(filename, points, colors):
    """points (N,3) float, colors (N,3) uint8 -> binary little-endian PLY with vertex x y z red green blue,
    the layout plyfile's ``PlyData([PlyElement.describe(vertex_all, 'vertex')]).write`` produces."""
    points = np.asarray(points, dtype=np.float32).reshape(-1, 3)
    colors = np.asarray(colors, dtype=np.uint8).reshape(-1, 3)
    if len(points) != len(colors):
        raise ValueError("points and colors differ in length")
    vertex = np.empty(len(points), dtype=[("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("red", "u1"), ("green", "u1"), ("blue", "u1")])
    vertex["x"], vertex["y"], vertex["z"] = points[:, 0], points[:, 1], points[:, 2]
    vertex["red"], vertex["green"], vertex["blue"] = colors[:, 0], colors[:, 1], colors[:, 2]
    header = ("ply\nformat binary_little_endian 1.0\nelement vertex {}\nproperty float x\nproperty float y\nproperty float z\n"
              "property uchar red\nproperty uchar green\nproperty uchar blue\nend_header\n").format(len(points))
    with open(filename, "wb") as f:
        f.write(header.encode("ascii"))
        vertex.tofile(f)


def read_ply(filename):
    """The inverse of write_ply (tests, tools): -> (points (N,3) float32, colors (N,3) uint8)."""
    with open(filename, "rb") as f:
        n = None
        while True:
            line = f.readline().decode("ascii").strip()
            if line.startswith("element vertex"):
                n = int(line.split()[-1])
            if line == "end_header":
                break
        v = np.fromfile(f, dtype=[("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("red", "u1"), ("green", "u1"), ("blue", "u1")], count=n)
    return np.stack([v["x"], v["y"], v["z"]], axis=1), np.stack([v["red"], v["green"], v["blue"]], axis=1)
