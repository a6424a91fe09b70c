"""CUDA-graph replay of `MipNerf.forward` for a fixed batch shape.

A 4096-ray forward is two kernel launches (~0.5 ms each on a B200) plus a 12 KB constant upload; the host side of
one call (ctypes marshalling, output allocation, argument checks) costs ~85 us, and at 512 rays per GPU (the
4096-ray batch split over 8 GPUs) the launch path is as long as the kernels.  `GraphedForward` captures ONE forward
— and, for a ray-sharded batch, the NCCL all-gather of the rendered pixels that follows it (the reference's only
inference-time collective is implicit in its single-GPU loop; its training collective is DDP's, train.py:60) — into
a CUDA graph on static buffers and replays it: one `cudaGraphLaunch` per step.

The rays live in a `RayStaging` device buffer (one H2D copy per step refreshes them), the outputs in the tensors
the captured forward returned; both keep their addresses for the life of the object.
"""
from __future__ import annotations

from typing import Optional

import torch

from .rays import Rays, RayStaging


class GraphedForward:
    def __init__(self, model, staging: RayStaging, white_bkgd: bool = True, device=None, world: int = 1,
                 group=None, gather: str = "fine_rgb", warmup: int = 3):
        """`staging`: this rank's rays (its shard of the global batch).  world > 1: every replay ends with one
        all_gather_into_tensor of this rank's fine RGB ([B,3]) or of all pixels (`gather='pixels'`: rgb, distance,
        acc of both levels, 10 floats per ray) into `self.gathered`."""
        self.model, self.staging, self.white = model, staging, bool(white_bkgd)
        self.device = torch.device(device) if device is not None else next(model.parameters()).device
        self.world, self.group, self.gather = int(world), group, gather
        b = staging.num_rays
        self.rays: Rays = staging.to(self.device)            # static input buffer (views of one allocation)
        self.gathered: Optional[torch.Tensor] = None
        if self.world > 1:
            width = 3 if gather == "fine_rgb" else 10
            self.gathered = torch.empty(self.world * b * width, device=self.device)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(1, warmup)):                   # kernels loaded, NCCL communicator built, caches warm
                self._run()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.ret = self._run()

    def _run(self):
        ret = self.model(self.rays, False, self.white)
        if self.world > 1:
            import torch.distributed as dist
            src = ret[-1][0] if self.gather == "fine_rgb" else ret.pixels
            dist.all_gather_into_tensor(self.gathered, src.reshape(-1), group=self.group)
        return ret

    def load(self, non_blocking: bool = True) -> None:
        """One H2D copy of the staging buffer's current host contents into the graph's input."""
        self.staging.to(self.device, non_blocking=non_blocking)

    def replay(self):
        """One graph launch on the current stream; returns the (static) LevelOutputs of the captured forward."""
        self.graph.replay()
        return self.ret

    def __call__(self, rays: Optional[Rays] = None):
        if rays is not None:
            self.staging.fill(rays)
            self.load()
        return self.replay()
