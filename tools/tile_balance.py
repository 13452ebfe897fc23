#!/usr/bin/env python
"""Shard imbalance of the headline frame on ONE GPU: renders every rank's row list of an N-way split by itself and
compares the kernel times (max / mean = what the slowest of N GPUs would cost against a perfect split).
usage: tools/tile_balance.py [root=128] [world=8] -> one JSON line per tile size"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from flux_b200 import JobConfiguration, SceneData  # noqa: E402
from flux_b200.worker import GpuContext, shard_rows  # noqa: E402


def main():
    root = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    sd = SceneData.from_yaml(os.path.join(ROOT, "scenes", "demo2.yml"))
    W, H = sd.output_settings.image_width, sd.output_settings.image_height
    ctx = GpuContext(0)
    ctx.set_scene(sd.flatten(), JobConfiguration(root, 5, 50))
    ctx.generate_samples(1, W)
    frame = ctx.frame_create(W, H)
    for tile in (1, 2, 4, 8):
        ms = []
        for r in range(world):
            rows = shard_rows(H, tile, r, world)
            best = None
            for _ in range(2):
                ctx.render_row_list_into_frame(rows, frame)
                ctx.sync()
                v = ctx.last_kernel_ms()
                best = v if best is None else min(best, v)
            ms.append(best)
        ms = np.array(ms)
        print(json.dumps({"tile_rows": tile, "world": world, "root": root, "kernel_ms": [round(float(v), 3) for v in ms],
                          "max_over_mean": float(ms.max() / ms.mean())}), flush=True)
    frame.close()
    ctx.close()


if __name__ == "__main__":
    main()
