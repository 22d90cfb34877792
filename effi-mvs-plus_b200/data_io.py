"""On-disk formats on either side of the hot path: PFM depth / confidence maps,
MVSNet camera text files and the fused point cloud as a binary PLY.  Pure host I/O, byte-compatible with
what upstream reads and writes:

    read_pfm / save_pfm     upstream datasets/data_io.py:61-126
    write_cam               upstream test_tank.py:176-193
    read_cam_file           upstream datasets/general_eval.py:60-80 and test_dtu_dypcd.py:120-131 (no /4, no interval logic)
    write_ply               upstream test_tank.py:551-570 / test_dtu_dypcd.py:339-350 (plyfile's binary_little_endian layout)
"""
from __future__ import annotations

import re
import sys

import numpy as np


def read_pfm(filename):
    """-> (data (H,W) or (H,W,3) float32, top row first; scale)."""
    with open(filename, "rb") as f:
        header = f.readline().decode("utf-8").rstrip()
        if header == "PF":
            color = True
        elif header == "Pf":
            color = False
        else:
            raise Exception("Not a PFM file.")
        m = re.match(r"^(\d+)\s(\d+)\s$", f.readline().decode("utf-8"))
        if not m:
            raise Exception("Malformed PFM header.")
        width, height = map(int, m.groups())
        scale = float(f.readline().rstrip())
        endian = "<" if scale < 0 else ">"
        scale = abs(scale)
        data = np.fromfile(f, endian + "f")
    data = np.reshape(data, (height, width, 3) if color else (height, width))
    return np.flipud(data), scale


def save_pfm(filename, image, scale=1):
    image = np.asarray(image)
    if image.dtype.name != "float32":
        raise Exception("Image dtype must be float32.")
    if image.ndim == 3 and image.shape[2] == 3:
        color = True
    elif image.ndim == 2 or (image.ndim == 3 and image.shape[2] == 1):
        color = False
    else:
        raise Exception("Image must have H x W x 3, H x W x 1 or H x W dimensions.")
    image = np.flipud(image)
    endian = image.dtype.byteorder
    if endian == "<" or (endian == "=" and sys.byteorder == "little"):
        scale = -scale
    with open(filename, "wb") as f:
        f.write(b"PF\n" if color else b"Pf\n")
        f.write("{} {}\n".format(image.shape[1], image.shape[0]).encode("utf-8"))
        f.write(("%f\n" % scale).encode("utf-8"))
        image.tofile(f)


def write_cam(filename, cam, depth_max, depth_min):
    """cam (2,4,4): [0] extrinsic, [1][:3,:3] intrinsic, [1][3][:2] the two trailing numbers upstream carries there."""
    with open(filename, "w") as f:
        f.write("extrinsic\n")
        for i in range(4):
            for j in range(4):
                f.write(str(cam[0][i][j]) + " ")
            f.write("\n")
        f.write("\n")
        f.write("intrinsic\n")
        for i in range(3):
            for j in range(3):
                f.write(str(cam[1][i][j]) + " ")
            f.write("\n")
        f.write("\n" + str(cam[1][3][0]) + " " + str(cam[1][3][1]) + " " + str(depth_max) + " " + str(depth_min) + "\n")


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
