"""Whole-scene run (BASELINE.json configs[4]): depth maps for every view of a synthetic N-view scene at
the DTU shape, reference views sharded round-robin over the ranks, ONE NCCL all-gather of the depth
maps, then the geometric-consistency fusion of the views each rank owns.

    python tools/scene_bench.py [--views 49] [--precision bf16x3]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/scene_bench.py
"""
import argparse
import json
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import effimvs_b200  # noqa: E402,F401
from effimvs_b200 import hotpath, scene, synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--views", type=int, default=49)
    ap.add_argument("--width", type=int, default=1600)
    ap.add_argument("--height", type=int, default=1184)
    ap.add_argument("--precision", default="bf16x3")
    ap.add_argument("--feature-cache", action="store_true", help="encode each image once per rank (block sharding)")
    ap.add_argument("--graph", action="store_true", help="feature cache + encode / cascade replayed as CUDA graphs")
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.benchmark = True
    from util import dtu_model
    model = dtu_model(hotpath.CudaHotPath(a.precision, native_projection=True), dev)
    for m in model.modules():
        if isinstance(m, torch.nn.Conv2d):
            m.weight.data = m.weight.data.contiguous(memory_format=torch.channels_last)
    N, W, H = a.views, a.width, a.height
    g = torch.Generator().manual_seed(0)
    imgs = torch.rand(N, 3, H, W, generator=g).to(dev)
    E, K = synthetic.camera_arc(N, W, H)
    cams = {k: v[0].to(dev) for k, v in synthetic.stage_cameras(E, K, 1).items()}
    dv = torch.linspace(1 / 935.0, 1 / 425.0, 384, device=dev)
    nb = lambda i, n: [j for d in range(1, n // 2 + 1) for j in ((i - d) % N, (i + d) % N)]      # noqa: E731
    pairs, fpairs = [nb(i, 4) for i in range(N)], [nb(i, 10) for i in range(N)]
    a.feature_cache = a.feature_cache or a.graph
    infer, fuse = scene.cuda_scene_callables(model, imgs, cams, dv, 2.0, 6.0, 2, 0.3, feature_cache=a.feature_cache,
                                             graphed_src_views=4 if a.graph else 0)
    sharding = "block" if a.feature_cache else "round_robin"
    with torch.no_grad():
        infer(rank % N, pairs[rank % N])            # warm-up (cuDNN autotune, lazy init)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    tm = {}
    t0 = time.perf_counter()
    out = scene.run_scene(infer, fuse, N, pairs, rank, world, dev, fuse_pairs=fpairs, timings=tm, sharding=sharding)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = torch.tensor([time.perf_counter() - t0], device=dev)
    pts = torch.tensor([float(sum(v[0].shape[0] for v in out.values()))], device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        dist.all_reduce(pts)
    if rank == 0:
        print(json.dumps({"scene_views": N, "n_gpus": world, "shape": [W, H], "seconds": float(dt), "depth_maps_per_sec_incl_fusion": N / float(dt),
                          "all_gather_ms_rank0": tm.get("all_gather_ms"), "fusion_ms_rank0": tm.get("fusion_ms"),
                          "fused_points": int(pts), "precision": a.precision, "feature_cache": a.feature_cache, "cuda_graphs": a.graph, "sharding": sharding,
                          "note": "4 source views for depth, 10 for fusion"}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
