"""Frame assembly over peer memory (flux_frame_*): every GPU / process renders its rows straight into ONE
framebuffer.  On a one-GPU box the cross-process path is exercised with two processes sharing cuda:0 (the child opens
the parent's frame through its CUDA IPC handle); the bytes must be those of a single full-frame render."""
import os
import subprocess
import sys

import numpy as np
import pytest

from flux_b200 import JobConfiguration
from flux_b200.worker import FluxError, shard_rows
from tests import helpers as Hp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import sys, numpy as np
sys.path.insert(0, sys.argv[1])
from flux_b200 import JobConfiguration, SceneData
from flux_b200.worker import GpuContext, shard_rows
handle = bytes.fromhex(sys.argv[2])
W, H, root, tile = (int(a) for a in sys.argv[3:7])
sd = SceneData.from_yaml(sys.argv[1] + "/scenes/demo2.yml").with_size(W, H)
cfg = JobConfiguration(root, 5, 50)
ctx = GpuContext(0)
ctx.set_scene(sd.flatten(), cfg)
ctx.generate_samples(7, W)
frame = ctx.frame_open_ipc(handle, W, H)
ctx.render_row_list_into_frame(shard_rows(H, tile, 1, 2), frame)
ctx.sync()
frame.close()
ctx.close()
print("child done")
"""


@pytest.mark.parametrize("root,tile", [(4, 1), (16, 4)])
def test_two_processes_render_into_one_frame_over_ipc(gpu_ctx, demo2, root, tile):
    W, H = 48, 36
    sd = demo2.with_size(W, H)
    cfg = JobConfiguration(root, 5, 50)
    gpu_ctx.set_kernel_mode(0)
    gpu_ctx.set_scene(sd.flatten(), cfg)
    gpu_ctx.generate_samples(7, W)
    full = gpu_ctx.render_rows(0, H - 1, W)
    frame = gpu_ctx.frame_create(W, H)
    try:
        handle = frame.export()
        assert len(handle) == 64
        gpu_ctx.render_row_list_into_frame(shard_rows(H, tile, 0, 2), frame)
        gpu_ctx.sync()
        half = frame.read()
        mine = shard_rows(H, tile, 0, 2)
        theirs = shard_rows(H, tile, 1, 2)
        assert np.array_equal(half[mine].view(np.uint64), full[mine].view(np.uint64))
        assert not half[theirs].any()                       # untouched rows are still zero
        r = subprocess.run([sys.executable, "-c", CHILD, ROOT, handle.hex(), str(W), str(H), str(root), str(tile)],
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        whole = frame.read()
        assert np.array_equal(whole.view(np.uint64), full.view(np.uint64))
    finally:
        frame.close()


def test_peer_alias_in_one_process_and_errors(gpu_ctx, demo2):
    """flux_frame_open_peer on the owner's own device is a plain alias (the in-process GpuWorker path with one GPU);
    every kernel variant scatters rows to their place; wrong sizes are refused."""
    W, H = 40, 24
    sd = demo2.with_size(W, H)
    cfg = JobConfiguration(16, 5, 50)
    gpu_ctx.set_scene(sd.flatten(), cfg)
    gpu_ctx.generate_samples(3, W)
    frame = gpu_ctx.frame_create(W, H)
    alias = gpu_ctx.frame_open_peer(frame)
    try:
        for mode in (1, 2, 4):
            gpu_ctx.set_kernel_mode(mode)
            ref = gpu_ctx.render_rows(0, H - 1, W)
            rows = np.array([1, 2, 5, 9, 23], np.uint32)
            gpu_ctx.render_row_list_into_frame(rows, alias)
            gpu_ctx.sync()
            got = frame.read()
            assert np.array_equal(got[rows].view(np.uint64), ref[rows].view(np.uint64)), mode
        assert alias.device_ptr() == frame.device_ptr()
        bad = gpu_ctx.frame_create(W + 1, H)
        with pytest.raises(FluxError):
            gpu_ctx.render_row_list_into_frame(np.array([0], np.uint32), bad)
        bad.close()
    finally:
        gpu_ctx.set_kernel_mode(0)
        alias.close()
        frame.close()
