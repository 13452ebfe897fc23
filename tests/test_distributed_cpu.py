"""N>1 host path on CPU: world_size-2 gloo group; each rank fills its shard (rendered by the CPU
oracle standing in for the GPU) and the FrameGather assembles the frame.  Checks the sharding plan,
padding, gather and un-interleave used by bench.py at N GPUs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from flux_b200 import JobConfiguration, SceneData
from flux_b200.sharding import FrameGather, FramePlan

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, height, width, tile, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle_py as O
        from tests import helpers as Hp
        sd = SceneData.from_yaml(os.path.join(ROOT, "scenes", "demo1.yml")).with_size(width, height)
        cfg = JobConfiguration(2, 4, 50)
        flat = sd.flatten()
        ss = Hp.oracle_samples(17, cfg, width, height)
        plan = FramePlan(height, width, tile, world)
        fg = FrameGather(plan, rank, torch.device("cpu"), dist)
        rows = plan.my_rows(rank)
        part = O.render_row_list(flat, cfg, ss, rows, threads=1)
        fg.mine[:len(rows)] = torch.from_numpy(part)
        frame = fg.gather().numpy().copy()
        full = O.render_rows(flat, cfg, ss, 0, height - 1, threads=1)
        ok = np.array_equal(frame.view(np.uint64), full.view(np.uint64))
        q.put((rank, ok, len(rows)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("height,tile", [(30, 4), (33, 4), (7, 8)])
def test_two_rank_gather_reassembles_frame(height, tile):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, height, 24, tile, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    assert sum(n for _, _, n in res) == height


def _handle_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from flux_b200.sharding import exchange_handle
        secret = bytes(range(64))
        got = exchange_handle(dist, rank, secret if rank == 0 else None, src=0)
        # the stream-ordered barrier of PeerFrame is a 4-byte all-reduce
        flag = torch.ones(1, dtype=torch.int32)
        dist.all_reduce(flag)
        q.put((rank, got == secret, int(flag[0])))
    finally:
        dist.destroy_process_group()


def test_two_rank_frame_handle_exchange():
    """PeerFrame's host side without a GPU: rank 0's 64-byte frame handle reaches rank 1 unchanged."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_handle_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True, 2), (1, True, 2)]


class _NoPeerCtx:
    """Stands in for a GpuContext on a box without peer access / CUDA IPC."""

    def __init__(self, fail_on):
        self.fail_on = fail_on

    def frame_create(self, w, h):
        if self.fail_on == "create":
            raise RuntimeError("cudaMalloc: no device")

        class _F:
            def export(self_inner):
                return bytes(64)

            def close(self_inner):
                pass
        return _F()

    def frame_open_ipc(self, handle, w, h):
        raise RuntimeError("cudaIpcOpenMemHandle: invalid device context")


def _fallback_worker(rank, world, port, fail_on, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from flux_b200.sharding import FramePlan, PeerFrame
        plan = FramePlan(12, 8, 1, world)
        pf, why = PeerFrame.try_create(_NoPeerCtx(fail_on), plan, rank, torch.device("cpu"), dist)
        q.put((rank, pf is None, why))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("fail_on", ["create", "open"])
def test_peer_frame_failure_is_agreed_by_every_rank(fail_on):
    """If the owner cannot create the frame, or another rank cannot map it, no rank is left waiting in a collective:
    all of them get (None, reason) and bench.py falls back to the NCCL gather together."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_fallback_worker, args=(r, world, port, fail_on, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(none for _, none, _ in res)
    assert res[0][2] == res[1][2] and ("cudaMalloc" in res[0][2] or "cudaIpcOpenMemHandle" in res[0][2])
