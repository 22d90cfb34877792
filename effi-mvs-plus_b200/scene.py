"""Per-scene runner: reference views sharded round-robin over ranks (one process per GPU), no
collective during depth inference, ONE all-gather of the final depth maps before fusion
(upstream hands depth maps over through PFM files, test_tank.py:304-373).
"""
from __future__ import annotations

from typing import Callable, Dict, List, Sequence

import torch
import torch.distributed as dist


def block_start(n_views: int, rank: int, world: int) -> int:
    """First view of rank's block under balanced block sharding (rank == world gives n_views)."""
    q, r = divmod(n_views, world)
    return rank * q + min(rank, r)


def shard_views(n_views: int, rank: int, world: int, mode: str = "round_robin") -> List[int]:
    """Reference views owned by `rank`.  "round_robin": rank, rank+world, ... ; "block": a contiguous run --
    neighbouring reference views share their source images, so a rank that caches encoded features (section 8(f)
    row 1) encodes each image at most once.  Both are balanced: every rank owns floor(n/world) views and the first
    n % world ranks one more, so a rank is empty only when n_views < world (49 views on 8 ranks: 7,6,6,6,6,6,6,6)."""
    if mode == "block":
        return list(range(block_start(n_views, rank, world), block_start(n_views, rank + 1, world)))
    return list(range(rank, n_views, world))


def slots_per_rank(n_views: int, world: int) -> int:
    return (n_views + world - 1) // world


def view_slot(i: int, n_views: int, world: int, mode: str = "round_robin"):
    """(owner rank, slot inside the owner's all-gather block) of reference view i."""
    if mode != "block":
        return i % world, i // world
    q, r = divmod(n_views, world)
    owner = i // (q + 1) if i < r * (q + 1) else r + (i - r * (q + 1)) // max(q, 1)
    return owner, i - block_start(n_views, owner, world)


def gather_depths(local: Dict[int, torch.Tensor], n_views: int, rank: int, world: int, h: int, w: int, device,
                  mode: str = "round_robin") -> torch.Tensor:
    """All-gather of the per-rank depth maps into one (n_views,h,w) tensor on every rank.

    Each rank contributes a (slots,h,w) block, zero padded where it owns fewer than `slots` views (a rank that owns
    none still contributes its zero block: every rank must enter the collective); view i lives at
    view_slot(i) = (owner, slot) of the rank-major concatenation."""
    slots = slots_per_rank(n_views, world)
    mine = torch.zeros(slots, h, w, device=device, dtype=torch.float32)
    for i, d in local.items():
        owner, slot = view_slot(i, n_views, world, mode)
        if owner != rank:
            raise ValueError("view {} belongs to rank {}, not {} ({} sharding)".format(i, owner, rank, mode))
        mine[slot] = d
    if world == 1 or not (dist.is_available() and dist.is_initialized()):
        flat = mine
    else:
        flat = torch.empty(world * slots, h, w, device=device, dtype=torch.float32)   # rank-major concatenation
        dist.all_gather_into_tensor(flat, mine)
    index = [o * slots + s for o, s in (view_slot(i, n_views, world, mode) for i in range(n_views))]
    if index == list(range(n_views)):
        return flat[:n_views]
    return flat.index_select(0, torch.as_tensor(index, device=flat.device))


def agree_on_size(local_hw, n_views: int, world: int, device):
    """(h, w) of the depth maps on every rank.  Only when n_views < world can a rank own no view and not know the
    size; then (and only then -- the condition is the same on all ranks) one 2-element MAX all-reduce tells it."""
    if n_views >= world or world == 1 or not (dist.is_available() and dist.is_initialized()):
        if local_hw is None:
            raise ValueError("no depth map on this rank and no peer to learn its size from")
        return local_hw
    t = torch.tensor(list(local_hw or (0, 0)), device=device, dtype=torch.int64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return int(t[0]), int(t[1])


@torch.no_grad()
def run_scene(infer: Callable, fuse: Callable, n_views: int, pairs: Sequence[Sequence[int]], rank: int = 0, world: int = 1,
              device="cuda", fuse_pairs: Sequence[Sequence[int]] = None, timings: dict = None, sharding: str = "round_robin"):
    """infer(ref_view, src_views) -> (depth (h,w), conf (hc,wc)) on `device`;
    fuse(ref_view, ref_depth (1,1,h,w), conf (1,hc,wc), src_views, src_depths (1,v,1,h,w)) -> dict with
    'final' (1,1,h,w) bool and 'points' (1,3,h,w).
    pairs[i]: source views used for the depth map of view i; fuse_pairs[i] (default: pairs[i]): source
    views whose depth maps the consistency filter of view i reads (upstream: 4 and up to 10).
    Returns {ref_view: (points (k,3), depth (h,w))} for the views this rank owns."""
    fuse_pairs = pairs if fuse_pairs is None else fuse_pairs
    mine = shard_views(n_views, rank, world, sharding)
    depths, confs = {}, {}
    timed = timings is not None and torch.cuda.is_available() and str(device).startswith("cuda")
    if timed:
        ev_start = torch.cuda.Event(enable_timing=True)
        ev_start.record()
    for i in mine:
        d, c = infer(i, list(pairs[i]))
        depths[i], confs[i] = d, c
    # a rank without a view (n_views < world) still enters the collective below with a zero block
    h, w = agree_on_size(tuple(next(iter(depths.values())).shape[-2:]) if depths else None, n_views, world, device)
    ev = None
    if timed:
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
    all_depths = gather_depths(depths, n_views, rank, world, h, w, device, sharding)
    if ev:
        ev[1].record()
    out = {}
    flat = [j for i in mine for j in fuse_pairs[i]]
    flat_dev = torch.as_tensor(flat, device=all_depths.device, dtype=torch.long) if flat else None      # ONE host -> device copy
    at = 0
    for i in mine:
        src = list(fuse_pairs[i])
        sel = flat_dev[at:at + len(src)]
        at += len(src)
        res = fuse(i, depths[i].reshape(1, 1, h, w), confs[i].unsqueeze(0), src, all_depths.index_select(0, sel).reshape(1, len(src), 1, h, w))
        m = res["final"][0, 0]
        out[i] = (res["points"][0][:, m].t().contiguous(), depths[i])
    if ev:
        ev[2].record()
        torch.cuda.synchronize()
        timings["depth_maps_ms"] = ev_start.elapsed_time(ev[0])
        timings["all_gather_ms"] = ev[0].elapsed_time(ev[1])
        timings["fusion_ms"] = ev[1].elapsed_time(ev[2])
    return out


class GraphedViewRunner:
    """Depth inference of one reference view from cached feature pyramids as two CUDA graphs: `encode` (one image
    through the FPN) and `forward` (the cascade on static feature / camera buffers).  The hot path has no host
    synchronisation, so both capture; per view the host only issues a handful of device-to-device copies and
    one replay instead of ~320 launches (section 8(f) row 1 + SURVEY section 8(e) "CUDA-graph the forward")."""

    def __init__(self, model, imgs: torch.Tensor, cams: Dict[str, torch.Tensor], depth_values: torch.Tensor, n_src: int):
        self.model, self.imgs, self.cams = model, imgs, cams
        self.stages = [k for k in ("stage1", "stage2", "stage3") if k in cams]
        V = n_src + 1
        dev = imgs.device
        self.cache: Dict[int, list] = {}
        self._sel: Dict[tuple, torch.Tensor] = {}
        self.stream = torch.cuda.Stream(dev)
        self.img_in = imgs[:1].reshape(1, 1, *imgs.shape[1:]).clone()
        self.ref_in = imgs[:1].clone()
        self.cam_in = {k: cams[k][:V].unsqueeze(0).clone() for k in self.stages}
        self.dv = depth_values.reshape(1, -1).clone()
        with torch.no_grad():
            torch.cuda.synchronize(dev)
            with torch.cuda.stream(self.stream):
                for _ in range(2):                                        # warm-up: cuDNN autotune, lazy init, workspace sizes
                    enc = [st[0] for st in model.encode(self.img_in)]
                self.feat_in = [[torch.empty_like(e) for _ in range(V)] for e in enc]
                for row, e in zip(self.feat_in, enc):
                    for t in row:
                        t.copy_(e)
                for _ in range(2):
                    model.forward_from_features(self.feat_in, self.ref_in, self.cam_in, self.dv)
            self.stream.synchronize()
            self.g_enc = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.g_enc, stream=self.stream):
                self.enc_out = [st[0] for st in model.encode(self.img_in)]
            self.g_fwd = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.g_fwd, stream=self.stream):
                out = model.forward_from_features(self.feat_in, self.ref_in, self.cam_in, self.dv)
            self.out = (out["depth"][-1], out["photometric_confidence"])

    def _index(self, idx):
        """device index tensor of a view list, built once: a host list -> device copy is a stream synchronisation, and one per
        view serialises the host's enqueue work with the previous view's graph"""
        key = tuple(idx)
        if key not in self._sel:
            self._sel[key] = torch.as_tensor(list(idx), device=self.imgs.device)
        return self._sel[key]

    @torch.no_grad()
    def encoded(self, j: int):
        if j not in self.cache:
            self.img_in.copy_(self.imgs[j].reshape(self.img_in.shape))
            self.g_enc.replay()
            self.cache[j] = [e.clone() for e in self.enc_out]
        return self.cache[j]

    @torch.no_grad()
    def infer(self, i: int, srcs: Sequence[int]):
        idx = [i] + list(srcs)
        if len(idx) != len(self.feat_in[0]):
            raise ValueError("GraphedViewRunner was captured for {} source views, got {}".format(len(self.feat_in[0]) - 1, len(srcs)))
        with torch.cuda.stream(self.stream):
            for v, j in enumerate(idx):
                for s, e in enumerate(self.encoded(j)):
                    self.feat_in[s][v].copy_(e)
            self.ref_in.copy_(self.imgs[i:i + 1])
            sel = self._index(idx)
            for k in self.stages:
                self.cam_in[k].copy_(self.cams[k].index_select(0, sel).unsqueeze(0))
            self.g_fwd.replay()
            depth, conf = self.out[0][0].clone(), self.out[1][0].clone()
        torch.cuda.current_stream().wait_stream(self.stream)
        return depth, conf


def cuda_scene_callables(model, imgs: torch.Tensor, cams: Dict[str, torch.Tensor], depth_values: torch.Tensor,
                         dist_base: float, rel_diff_base: float, thres_view: int, prob_threshold: float,
                         feature_cache: bool = False, graphed_src_views: int = 0):
    """infer / fuse callables for `run_scene` on the CUDA path.
    imgs (Nv,3,H,W) on the device; cams {"stage1".."stage4": (Nv,2,4,4)}; depth_values (Dv).
    feature_cache: encode every image once per rank and reuse its feature pyramid for each reference view
    that lists it as a source (section 8(f) row 1; eval-mode results are those of re-encoding it).
    graphed_src_views > 0: run encode / cascade as CUDA graphs captured for that many source views (implies the
    feature cache)."""
    from . import fusion
    cache: Dict[int, list] = {}
    runner = GraphedViewRunner(model, imgs, cams, depth_values, graphed_src_views) if graphed_src_views > 0 else None

    def encoded(j):
        if j not in cache:
            cache[j] = [stage[0] for stage in model.encode(imgs[j].reshape(1, 1, *imgs.shape[1:]))]
        return cache[j]

    def infer(i, srcs):
        if runner is not None:
            return runner.infer(i, srcs)
        idx = [i] + list(srcs)
        stage_cams = {k: v[idx].unsqueeze(0) for k, v in cams.items() if k != "stage4"}
        if feature_cache:
            per_view = [encoded(j) for j in idx]
            feats = [[pv[s] for pv in per_view] for s in range(3)]
            out = model.forward_from_features(feats, imgs[i].unsqueeze(0), stage_cams, depth_values.unsqueeze(0))
        else:
            out = model(imgs[idx].unsqueeze(0), stage_cams, depth_values.unsqueeze(0))
        return out["depth"][-1][0], out["photometric_confidence"][0]

    def clear_cache():
        """forget the encoded feature pyramids (a new scene, or a benchmark pass that must encode again)"""
        cache.clear()
        if runner is not None:
            runner.cache.clear()
    infer.clear_cache = clear_cache

    sel_cache: Dict[tuple, torch.Tensor] = {}

    def fuse(i, ref_depth, conf, srcs, src_depths):
        full = cams["stage4"]
        key = tuple(srcs)
        if key not in sel_cache:        # one host -> device copy (a synchronisation) per distinct source list, not per call
            sel_cache[key] = torch.as_tensor(list(srcs), device=full.device)
        return fusion.filter_view(ref_depth, conf, src_depths, full[i].unsqueeze(0), full.index_select(0, sel_cache[key]).unsqueeze(0),
                                  dist_base, rel_diff_base, thres_view, prob_threshold, torch_inverse=False)
    return infer, fuse
