"""Multi-GPU sharding of the image and the single framebuffer gather.

Replaces the reference's dynamic row-band queue (manager.rs:100, job.rs:66-88) and flux-node's TCP tile
distribution: rank r of N owns rows with (row // tile_rows) % N == r; after rendering, ONE
all_gather of the packed per-rank slices (padded to equal size) and an index_copy un-interleave.
Works on any torch device/backend (NCCL on GPUs; gloo on CPU for the tests).
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np

from .worker import shard_rows


class FramePlan:
    def __init__(self, height: int, width: int, tile_rows: int, world: int):
        self.height, self.width, self.tile_rows, self.world = height, width, tile_rows, world
        self.rows: List[np.ndarray] = [shard_rows(height, tile_rows, r, world) for r in range(world)]
        self.max_rows = max(len(r) for r in self.rows)

    def my_rows(self, rank: int) -> np.ndarray:
        return self.rows[rank]


class FrameGather:
    """Owns the device buffers of one rank: packed slice, gathered slices, assembled frame."""

    def __init__(self, plan: FramePlan, rank: int, device, dist=None):
        import torch
        self.plan, self.rank, self.dist = plan, rank, dist
        self.mine = torch.zeros((plan.max_rows, plan.width, 3), dtype=torch.float64, device=device)
        # flat [world*max_rows] leading dim: the shape both NCCL and gloo accept for all_gather_into_tensor
        self.gathered_flat = (torch.empty((plan.world * plan.max_rows, plan.width, 3), dtype=torch.float64, device=device)
                              if plan.world > 1 else None)
        self.gathered = (self.gathered_flat.view(plan.world, plan.max_rows, plan.width, 3)
                         if plan.world > 1 else None)
        self.frame = torch.empty((plan.height, plan.width, 3), dtype=torch.float64, device=device)
        self.row_index = [torch.from_numpy(r.astype(np.int64)).to(device) for r in plan.rows]

    def gather(self):
        """all_gather the packed slices and un-interleave into self.frame (every rank gets the frame)."""
        p = self.plan
        if p.world > 1:
            self.dist.all_gather_into_tensor(self.gathered_flat, self.mine)
            for r in range(p.world):
                self.frame.index_copy_(0, self.row_index[r], self.gathered[r, :len(p.rows[r])])
        else:
            self.frame.index_copy_(0, self.row_index[0], self.mine[:len(p.rows[0])])
        return self.frame
