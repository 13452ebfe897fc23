#!/usr/bin/env python
"""Throughput of the BASELINE.json configs that are parity cases, not the bench line.
One GPU, through the C-ABI.  Prints one JSON line per config.  usage: tools/bench_configs.py [c1] [c3] [c4] [c5]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from flux_b200 import JobConfiguration, SceneData, synth  # noqa: E402
from flux_b200.opsmodel import algorithmic_ops  # noqa: E402
from flux_b200.worker import GpuContext  # noqa: E402


def render_config(ctx, name, sd, root, depth=5, seed=1, reps=2):
    W, H = sd.output_settings.image_width, sd.output_settings.image_height
    cfg = JobConfiguration(root, depth, 50)
    t0 = time.perf_counter()
    flat = sd.flatten()
    t_flat = time.perf_counter() - t0
    t0 = time.perf_counter()
    ctx.set_scene(flat, cfg)
    t_scene = time.perf_counter() - t0
    ctx.generate_samples(seed, W)
    best = None
    for _ in range(reps):
        ctx.render_rows(0, H - 1, W)
        ms = ctx.last_kernel_ms()
        best = ms if best is None else min(best, ms)
    ctx.enable_counters(True)
    ctx.reset_counters()
    ctx.render_row_list(np.arange(0, H, 8, dtype=np.uint32), W)   # counters on every 8th row (instrumented kernel is slower)
    cn = ctx.counters()
    ctx.enable_counters(False)
    n = W * H * root * root
    out = {"config": name, "image": [W, H], "spp": root * root, "shapes": flat.n_shapes, "kernel_ms": best,
           "Msamples_per_s": n / (best * 1e-3) / 1e6, "flatten_s": t_flat, "set_scene_s": t_scene,
           "segments_per_sample": cn["segments"] / max(1, cn["samples"]),
           "ops_per_sample": algorithmic_ops(cn) / max(1, cn["samples"]),
           "nodes_per_segment": cn["nodes_visited"] / max(1, cn["segments"]),
           "prim_tests_per_segment": (cn["bbox_tests"] + cn["tri_tests"]) / max(1, cn["segments"])}
    # SURVEY.md §8d for BVH configs: algorithmic bytes = nodes_visited * node_bytes + prims_tested * prim_bytes
    # (128-byte nodes, 112-byte sphere records, 96-byte triangle records), set against the HBM peak
    seg = max(1, cn["segments"])
    bytes_per_segment = (cn["nodes_visited"] * 128 + cn["bbox_tests"] * 112 + cn["tri_tests"] * 96) / seg
    out["alg_bytes_per_segment"] = bytes_per_segment
    out["alg_GB_per_s"] = bytes_per_segment * out["segments_per_sample"] * n / (best * 1e-3) / 1e9
    print(json.dumps(out), flush=True)
    return out


def c5(ctx, n_total=100_000_000, chunk=10_000_000, reps=1):
    sd = synth.sphere_cloud_scene(10_000, seed=5)
    flat = sd.flatten()
    ctx.set_scene(flat, JobConfiguration(1))
    ms_total, hits, csum = 0.0, 0, 0
    for k in range(n_total // chunk):
        o, d = synth.random_rays(chunk, seed=5, chunk_offset=k)
        best = None
        for _ in range(reps):   # the GPU idles while the host makes the next chunk: repeat and keep the fastest launch
            hit, t = ctx.trace_rays(o, d)
            ms = ctx.last_kernel_ms()
            best = ms if best is None else min(best, ms)
        ms_total += best
        hits += int((hit >= 0).sum())
        csum = (csum + int(hit.astype(np.int64).sum())) & 0xFFFFFFFFFFFF
    # event counts on the first chunk (the instrumented instantiation is slower: not part of the timing)
    o, d = synth.random_rays(chunk, seed=5, chunk_offset=0)
    ctx.enable_counters(True)
    ctx.reset_counters()
    ctx.trace_rays(o, d)
    cn = ctx.counters()
    ctx.enable_counters(False)
    bytes_per_ray = (cn["nodes_visited"] * 128 + cn["bbox_tests"] * 112) / max(1, cn["segments"])
    out = {"config": "c5 100M rays x 10K spheres (BVH)", "rays": n_total, "kernel_ms": ms_total,
           "Mrays_per_s": n_total / (ms_total * 1e-3) / 1e6, "hit_fraction": hits / n_total, "hit_id_checksum": csum,
           "nodes_per_ray": cn["nodes_visited"] / max(1, cn["segments"]), "sphere_tests_per_ray": cn["bbox_tests"] / max(1, cn["segments"]),
           "quadratics_per_ray": cn["bbox_pass"] / max(1, cn["segments"]), "alg_bytes_per_ray": bytes_per_ray,
           "alg_GB_per_s": bytes_per_ray * n_total / (ms_total * 1e-3) / 1e9}
    print(json.dumps(out), flush=True)
    # brute force (the reference's algorithm) on 2 M rays for comparison and a bitwise check
    o, d = synth.random_rays(2_000_000, seed=5, chunk_offset=0)
    hb, tb = ctx.trace_rays(o, d)
    ctx.set_accel_mode(1)
    ctx.set_scene(flat, JobConfiguration(1))
    hl, tl = ctx.trace_rays(o, d)
    ms_lin = ctx.last_kernel_ms()
    ctx.set_accel_mode(0)
    same = bool(np.array_equal(hb, hl) and np.array_equal(tb.view(np.uint64), tl.view(np.uint64)))
    print(json.dumps({"config": "c5 linear scan 2M rays", "kernel_ms": ms_lin, "Mrays_per_s": 2e6 / (ms_lin * 1e-3) / 1e6,
                      "G_sphere_tests_per_s": 2e6 * 1e4 / (ms_lin * 1e-3) / 1e9, "bvh_bitwise_equal": same}), flush=True)


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--mode=") and not a.startswith("--accel=")]
    accel = next((int(a.split("=")[1]) for a in sys.argv[1:] if a.startswith("--accel=")), 0)
    mode = next((int(a.split("=")[1]) for a in sys.argv[1:] if a.startswith("--mode=")), 0)
    which = set(args) or {"c1", "c3", "c4", "c5"}
    ctx = GpuContext(0)
    ctx.set_kernel_mode(mode)   # 0 = auto; A/B of render kernels on the same config
    ctx.set_accel_mode(accel)   # 0 = auto, 1 = linear scan, 2 = BVH
    if "c1" in which:
        sd = SceneData.from_yaml(os.path.join(ROOT, "scenes", "demo1.yml")).with_size(512, 512)
        render_config(ctx, "c1 demo1 512x512 @16spp", sd, 4, reps=5)
    if "c5" in which:
        c5(ctx)
    if "c5q" in which:   # quick A/B form: 20 M rays, best of 4 launches per chunk
        c5(ctx, 20_000_000, 10_000_000, reps=4)
    if "c3" in which:
        render_config(ctx, "c3 1M-triangle mesh 800x600 @1024spp (BVH)", synth.mesh_scene(1000, 500, seed=3), 32)
    if "c4" in which:
        render_config(ctx, "c4 glossy 1920x1080 @4096spp", synth.glossy_scene(), 64)
    for a in sorted(which):   # c4gN: the config-4 scene with an N x N sphere grid at 960x540 (linear-scan / BVH break-even)
        if a.startswith("c4g"):
            g = int(a[3:])
            render_config(ctx, f"c4 scene, {g}x{g} grid, 960x540 @4096spp", synth.glossy_scene(960, 540, grid=g), 64)
    ctx.close()


if __name__ == "__main__":
    main()
