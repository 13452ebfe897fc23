"""Multi-GPU sharding of the image and the assembly of the one frame.

Replaces the reference's dynamic row-band queue (manager.rs:100, job.rs:66-88) and flux-node's TCP tile
distribution: rank r of N owns rows with (row // tile_rows) % N == r.

Assembly, product path (``PeerFrame``): rank 0 owns ONE framebuffer on its GPU (``flux_frame_create``), hands its
64-byte CUDA IPC handle to the other ranks, and every rank's render kernel stores its pixels straight into that
buffer over NVLink / NVSwitch peer memory (``flux_render_row_list_into_frame``) — no gather collective, no packed
slices, no un-interleave.  What is left of "communication" is one 4-byte all-reduce used as a stream-ordered barrier.

``FrameGather`` is the collective form of the same assembly (ONE all_gather of packed, padded slices and an
index_copy un-interleave).  It works on any torch device/backend, so the world-size-2 gloo tests exercise the plan
with it on CPU, and ``bench.py --gather nccl`` keeps it for A/B against the peer-memory path.
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


def exchange_handle(dist, rank: int, handle: Optional[bytes], src: int = 0) -> bytes:
    """Rank `src` passes its frame handle (64 opaque bytes) to every rank; works on NCCL and gloo groups alike."""
    if dist is None:
        if handle is None:
            raise ValueError("no process group and no handle")
        return handle
    box = [handle if rank == src else None]
    dist.broadcast_object_list(box, src=src)
    h = box[0]
    if not isinstance(h, (bytes, bytearray)) or len(h) != 64:
        raise RuntimeError("frame handle exchange failed")
    return bytes(h)


class PeerFrame:
    """The frame every rank renders into (module docstring).  ``render()`` queues this rank's rows; ``barrier()``
    orders rank 0's later reads after every rank's render, on the device (stream-ordered, no host wait)."""

    def __init__(self, ctx, plan: FramePlan, rank: int, device=None, dist=None, owner_rank: int = 0):
        self.ctx, self.plan, self.rank, self.dist, self.owner_rank = ctx, plan, rank, dist, owner_rank
        self.is_owner = rank == owner_rank
        handle = None
        if self.is_owner:
            self.frame = ctx.frame_create(plan.width, plan.height)
            if plan.world > 1:
                handle = self.frame.export()
        if plan.world > 1:
            handle = exchange_handle(dist, rank, handle, owner_rank)
            if not self.is_owner:
                self.frame = ctx.frame_open_ipc(handle, plan.width, plan.height)
        self._flag = None
        if plan.world > 1:
            import torch
            self._flag = torch.zeros(1, dtype=torch.int32, device=device)
        self.my_rows = plan.my_rows(rank)

    @classmethod
    def try_create(cls, ctx, plan: FramePlan, rank: int, device=None, dist=None, owner_rank: int = 0):
        """PeerFrame on every rank, or (None, reason) on every rank: a rank that cannot create / export / map the
        frame (peer access or CUDA IPC unavailable) must not leave the others waiting in a collective, so every step
        that can fail is followed by an agreement.  Returns (frame, None) or (None, reason)."""
        self = cls.__new__(cls)
        self.ctx, self.plan, self.rank, self.dist, self.owner_rank = ctx, plan, rank, dist, owner_rank
        self.is_owner = rank == owner_rank
        self.frame, self._flag = None, None
        self.my_rows = plan.my_rows(rank)
        msg = None
        if self.is_owner:
            try:
                self.frame = ctx.frame_create(plan.width, plan.height)
                msg = self.frame.export() if plan.world > 1 else b""
            except Exception as e:   # noqa: BLE001 — the reason travels to every rank
                msg = f"rank {rank}: {e}"
        if plan.world == 1:
            return (self, None) if self.frame is not None else (None, msg)
        box = [msg if self.is_owner else None]
        dist.broadcast_object_list(box, src=owner_rank)
        msg = box[0]
        reason = None
        if not isinstance(msg, (bytes, bytearray)) or len(msg) != 64:
            reason = str(msg)
        elif not self.is_owner:
            try:
                self.frame = ctx.frame_open_ipc(bytes(msg), plan.width, plan.height)
            except Exception as e:   # noqa: BLE001
                reason = f"rank {rank}: {e}"
        import torch
        ok = torch.tensor([0 if reason else 1], dtype=torch.int32, device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok[0]) == 0:
            reasons = [None] * plan.world
            dist.all_gather_object(reasons, reason)
            self.close()   # mappings first, then (after a barrier) the owner's allocation
            return None, "; ".join(r for r in reasons if r) or "unknown"
        self._flag = torch.zeros(1, dtype=torch.int32, device=device)
        return self, None

    def render(self, stream_ptr: int = 0):
        self.ctx.render_row_list_into_frame(self.my_rows, self.frame, stream_ptr)

    def barrier(self):
        if self.plan.world > 1:
            self.dist.all_reduce(self._flag)

    def read(self, out=None):
        return self.frame.read(out)

    def close(self):
        """Unmap / free the frame.  The owner frees its allocation only after every other rank has closed its IPC
        mapping (freeing exported memory that another process still has open is undefined)."""
        if getattr(self, "_closed", False):
            return
        self._closed = True
        if self.plan.world > 1:   # every rank takes part in the barrier, with or without a mapping of its own
            if not self.is_owner and self.frame is not None:
                self.frame.close()
            self.dist.barrier()
            if self.is_owner and self.frame is not None:
                self.frame.close()
        elif self.frame is not None:
            self.frame.close()
        self.frame = None
