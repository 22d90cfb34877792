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


def _inputs(n_views=N_VIEWS):
    E, K = synthetic.camera_ring(n_views, W, H)
    depths = synthetic.render_plane_scene(E, K, W, H, noise=0.05, seed=1)
    cams = synthetic.stage_cameras(E, K, 1)["stage4"]
    pairs = [[(i + k) % n_views for k in range(1, min(4, n_views))] for i in range(n_views)]
    return depths, cams, pairs


def _run(rank, world, sharding="round_robin", n_views=N_VIEWS):
    from oracle import fusion as ofu
    depths, cams, pairs = _inputs(n_views)
    thres_view = min(2, n_views - 1)

    def infer(i, srcs):
        return depths[i], torch.full((H // 2, W // 2), 0.9)

    def fuse(i, ref_depth, conf, srcs, src_depths):
        return ofu.fuse_view(ref_depth, conf, src_depths, cams[:, i], cams[:, srcs], 1, 0.5, thres_view, 0.3)
    return scene.run_scene(infer, fuse, n_views, pairs, rank, world, device="cpu", sharding=sharding)


def _worker(rank, world, port, out_dir, sharding="round_robin", n_views=N_VIEWS):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res = _run(rank, world, sharding, n_views)
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


@pytest.mark.parametrize("sharding", ["round_robin", "block"])
def test_rank_without_a_view_still_joins_the_collective(tmp_path, sharding):
    """2 views on 3 ranks: rank 2 owns nothing, learns the map size from its peers, contributes a zero block to the
    all-gather and returns an empty result instead of raising while the others wait in the collective."""
    n, world = 2, 3
    single = _run(0, 1, n_views=n)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, str(tmp_path), sharding, n), nprocs=world, join=True)
    parts = [torch.load(os.path.join(str(tmp_path), "rank{}.pt".format(r))) for r in range(world)]
    assert [sorted(p) for p in parts] == [[0], [1], []]
    for i in range(n):
        assert torch.equal(parts[i][i][0], single[i][0]) and torch.equal(parts[i][i][1], single[i][1])


@pytest.mark.parametrize("n,world", [(49, 8), (5, 4), (3, 8), (7, 2), (3, 1), (16, 8)])
@pytest.mark.parametrize("mode", ["round_robin", "block"])
def test_sharding_is_balanced_and_gather_layout_roundtrips(n, world, mode):
    shards = [scene.shard_views(n, r, world, mode) for r in range(world)]
    assert sorted(i for sh in shards for i in sh) == list(range(n))                  # a partition of the views
    assert max(map(len, shards)) - min(map(len, shards)) <= 1                          # balanced: (49, 8) -> 7,6,6,6,6,6,6,6
    assert all(sh for sh in shards) or n < world                                       # empty ranks only when n < world
    if mode == "block":
        assert all(sh == list(range(sh[0], sh[0] + len(sh))) for sh in shards if sh)   # contiguous
    slots = scene.slots_per_rank(n, world)
    blocks = []
    for r in range(world):       # what every rank would contribute, gathered by hand
        blk = torch.zeros(slots, 2, 2)
        for i in shards[r]:
            owner, slot = scene.view_slot(i, n, world, mode)
            assert owner == r and 0 <= slot < slots
            blk[slot] = float(i) + 1
        blocks.append(blk)
    flat = torch.cat(blocks)
    index = [o * slots + s for o, s in (scene.view_slot(i, n, world, mode) for i in range(n))]
    assert [int(flat[j, 0, 0]) for j in index] == [i + 1 for i in range(n)]
    # the single-process path of gather_depths uses the same indexing
    if world == 1:
        got = scene.gather_depths({i: torch.full((2, 2), float(i)) for i in range(n)}, n, 0, 1, 2, 2, "cpu", mode)
        assert [int(got[i, 0, 0]) for i in range(n)] == list(range(n))
