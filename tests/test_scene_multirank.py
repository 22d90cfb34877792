"""Sharded scene runner on 2 ranks (gloo, CPU): reference views round-robin over ranks, ONE
all-gather of depth maps, fusion per owned view -- must equal the single-rank run bit for bit.
The depth inference and the fusion are injected (rendered depths / oracle fusion): this test covers
the host-side sharding + collective logic; the CUDA kernels behind the same callables are covered
by test_gpu_parity."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import effimvs_b200  # noqa: F401
from effimvs_b200 import scene, synthetic

N_VIEWS, H, W = 7, 24, 32


def _inputs():
    E, K = synthetic.camera_ring(N_VIEWS, W, H)
    depths = synthetic.render_plane_scene(E, K, W, H, noise=0.05, seed=1)
    cams = synthetic.stage_cameras(E, K, 1)["stage4"]
    pairs = [[(i + k) % N_VIEWS for k in (1, 2, 3)] for i in range(N_VIEWS)]
    return depths, cams, pairs


def _run(rank, world, sharding="round_robin"):
    from oracle import fusion as ofu
    depths, cams, pairs = _inputs()

    def infer(i, srcs):
        return depths[i], torch.full((H // 2, W // 2), 0.9)

    def fuse(i, ref_depth, conf, srcs, src_depths):
        return ofu.fuse_view(ref_depth, conf, src_depths, cams[:, i], cams[:, srcs], 1, 0.5, 2, 0.3)
    return scene.run_scene(infer, fuse, N_VIEWS, pairs, rank, world, device="cpu", sharding=sharding)


def _worker(rank, world, port, out_dir, sharding="round_robin"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res = _run(rank, world, sharding)
        torch.save({k: (p, d) for k, (p, d) in res.items()}, os.path.join(out_dir, "rank{}.pt".format(rank)))
    finally:
        dist.destroy_process_group()


import pytest  # noqa: E402


@pytest.mark.parametrize("sharding", ["round_robin", "block"])
def test_two_ranks_equal_one(tmp_path, sharding):
    single = _run(0, 1)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path), sharding), nprocs=2, join=True)
    merged = {}
    for r in range(2):
        part = torch.load(os.path.join(str(tmp_path), "rank{}.pt".format(r)))
        assert sorted(part) == scene.shard_views(N_VIEWS, r, 2, sharding)
        merged.update(part)
    assert sorted(merged) == list(range(N_VIEWS))
    for i in range(N_VIEWS):
        assert torch.equal(merged[i][0], single[i][0]) and torch.equal(merged[i][1], single[i][1])
    assert sum(v[0].shape[0] for v in single.values()) > 0.3 * N_VIEWS * H * W


def test_gather_layout_roundtrip():
    for n, world in ((7, 2), (49, 8), (5, 4), (3, 1)):
        slots = scene.slots_per_rank(n, world)
        blocks = []
        for r in range(world):
            blk = torch.zeros(slots, 2, 2)
            for i in scene.shard_views(n, r, world):
                blk[i // world] = float(i)
            blocks.append(blk)
        full = torch.stack(blocks).permute(1, 0, 2, 3).reshape(slots * world, 2, 2)[:n]
        assert [int(full[i, 0, 0]) for i in range(n)] == list(range(n))


def test_block_gather_layout_roundtrip():
    for n, world in ((7, 2), (49, 8), (5, 4), (3, 1)):
        local_all = {}
        for r in range(world):
            local_all[r] = {i: torch.full((2, 2), float(i)) for i in scene.shard_views(n, r, world, "block")}
        assert sorted(i for d in local_all.values() for i in d) == list(range(n))
        slots = scene.slots_per_rank(n, world)
        blocks = []
        for r in range(world):
            blk = torch.zeros(slots, 2, 2)
            for i, d in local_all[r].items():
                blk[i % slots] = d
            blocks.append(blk)
        full = torch.stack(blocks).reshape(world * slots, 2, 2)[:n]
        assert [int(full[i, 0, 0]) for i in range(n)] == list(range(n))
