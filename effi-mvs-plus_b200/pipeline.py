"""DepthMapPipeline: the end-to-end entry point for host-resident inputs.

A caller hands over one 5-view sample in host memory and gets the depth map back in host memory.
The pipeline keeps `slots` copies of the static device inputs, each with its own CUDA graph of the
whole cascade (possible because the hot path has no host synchronisation), and overlaps

    host -> device copy of sample k+1   (copy stream)
    forward of sample k                 (compute stream, one graph replay)
    device -> host read of depth k      (read-back stream, overlapping the forward of sample k+1)

so that the PCIe transfer of the 113 MB of images per sample hides behind the previous sample's
compute.  Upstream's drivers do the same steps strictly one after the other
(test_dtu_dypcd.py:424-451: tocuda, model(...), tensor2numpy).
"""
from __future__ import annotations

from typing import Dict, List

import torch


class _Slot:
    def __init__(self):
        self.inputs = None          # static device tensors the graph reads
        self.graph = None
        self.out = None             # device outputs of the graph
        self.host_depth = None      # pinned host outputs
        self.host_conf = None
        self.copied = torch.cuda.Event()
        self.computed = torch.cuda.Event()
        self.done = torch.cuda.Event()
        self.busy = False


class DepthMapPipeline:
    def __init__(self, model, example: Dict, slots: int = 2, use_graph: bool = True, before_replay=None):
        """example: {"imgs" (B,V,3,H,W), "proj_matrices" {stage: (B,V,2,4,4)}, "depth_values" (B,Dv)} on the host;
        fixes the shapes the pipeline accepts.  "imgs" may be fp32 in [0,1] (what upstream's loaders hand over) or uint8
        (what the image files hold): 8-bit images cross PCIe as they are -- a quarter of the bytes -- and are divided by
        255 on the device with the loaders' own arithmetic (effimvs_images_u8_to_f32), so the depth maps are the same
        bit for bit.  `submit` accepts either, whatever the example was."""
        self.model = model
        self.device = next(model.parameters()).device
        self.compute = torch.cuda.Stream(self.device)
        self.copy = torch.cuda.Stream(self.device)
        self.readback = torch.cuda.Stream(self.device)
        self.slots: List[_Slot] = []
        self.graphed = use_graph
        self.before_replay = before_replay      # optional callable enqueued on the compute stream before each forward
        self._n = 0
        stages = [k for k in example["proj_matrices"] if k in ("stage1", "stage2", "stage3")]
        with torch.no_grad():
            for _ in range(slots):
                s = _Slot()
                first = example["imgs"].to(self.device)
                s.u8 = torch.empty(first.shape, dtype=torch.uint8, device=self.device)
                if first.dtype == torch.uint8:
                    from . import ops
                    s.u8.copy_(first)
                    first = torch.empty(first.shape, dtype=torch.float32, device=self.device)
                    ops.images_u8_to_f32(s.u8, first)
                s.inputs = {"imgs": first, "depth_values": example["depth_values"].to(self.device),
                            "proj_matrices": {k: example["proj_matrices"][k].to(self.device) for k in stages}}
                torch.cuda.synchronize(self.device)
                with torch.cuda.stream(self.compute):
                    for _ in range(2):                                   # warm-up (cuDNN autotune, lazy init)
                        out = model(s.inputs["imgs"], s.inputs["proj_matrices"], s.inputs["depth_values"])
                self.compute.synchronize()
                if use_graph:
                    try:
                        g = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(g, stream=self.compute):
                            out = model(s.inputs["imgs"], s.inputs["proj_matrices"], s.inputs["depth_values"])
                        s.graph = g
                    except Exception:                                    # capture is an optimisation only
                        self.graphed = False
                        s.graph = None
                        torch.cuda.synchronize(self.device)
                s.out = (out["depth"][-1], out["photometric_confidence"])
                s.host_depth = torch.empty(s.out[0].shape, dtype=torch.float32).pin_memory()
                s.host_conf = torch.empty(s.out[1].shape, dtype=torch.float32).pin_memory()
                self.slots.append(s)

    def h2d_bytes(self, sample: Dict) -> int:
        """bytes `submit(sample)` copies host -> device"""
        return (sample["imgs"].numel() * sample["imgs"].element_size() + 4 * sample["depth_values"].numel() +
                4 * sum(sample["proj_matrices"][k].numel() for k in self.slots[0].inputs["proj_matrices"]))

    @property
    def d2h_bytes(self) -> int:
        return 4 * (self.slots[0].host_depth.numel() + self.slots[0].host_conf.numel())

    @torch.no_grad()
    def submit(self, sample: Dict) -> int:
        """Enqueue one host sample (pinned memory recommended).  Returns a ticket for `result`."""
        i = self._n % len(self.slots)
        s = self.slots[i]
        if s.busy:
            s.done.synchronize()            # the slot's previous outputs have reached the host
        with torch.cuda.stream(self.copy):
            self.copy.wait_event(s.done)    # do not overwrite inputs a previous replay may still read
            as_u8 = sample["imgs"].dtype == torch.uint8
            (s.u8 if as_u8 else s.inputs["imgs"]).copy_(sample["imgs"], non_blocking=True)
            s.inputs["depth_values"].copy_(sample["depth_values"], non_blocking=True)
            for k, v in s.inputs["proj_matrices"].items():
                v.copy_(sample["proj_matrices"][k], non_blocking=True)
            s.copied.record(self.copy)
        with torch.cuda.stream(self.compute):
            self.compute.wait_event(s.copied)
            self.compute.wait_event(s.done)     # the slot's previous outputs have left the device (no-op the first time)
            if as_u8:
                from . import ops
                ops.images_u8_to_f32(s.u8, s.inputs["imgs"])
            if self.before_replay is not None:
                self.before_replay()
            if s.graph is not None:
                s.graph.replay()
            else:
                out = self.model(s.inputs["imgs"], s.inputs["proj_matrices"], s.inputs["depth_values"])
                s.out = (out["depth"][-1], out["photometric_confidence"])
            s.computed.record(self.compute)
        with torch.cuda.stream(self.readback):  # device -> host on its own stream: the next sample's forward starts at once
            self.readback.wait_event(s.computed)
            if s.graph is None:
                for t in s.out:
                    t.record_stream(self.readback)   # eager outputs come from the caching allocator
            s.host_depth.copy_(s.out[0], non_blocking=True)
            s.host_conf.copy_(s.out[1], non_blocking=True)
            s.done.record(self.readback)
        s.busy = True
        self._n += 1
        return i

    def result(self, ticket: int):
        """Blocks until the sample of `ticket` is in host memory; returns (depth, confidence) pinned host tensors."""
        s = self.slots[ticket]
        s.done.synchronize()
        s.busy = False
        return s.host_depth, s.host_conf
