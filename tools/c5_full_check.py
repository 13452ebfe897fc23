#!/usr/bin/env python
"""Config 5 in full: all 100 M rays against the 10 K spheres through the BVH AND through the GPU's
brute-force linear scan (the reference's algorithm, scene.rs:156-160), compared bitwise (ids and distances), chunk by
chunk.  Writes one JSON summary (checksums, hit fraction, timings) to the path given (default: stdout only).
usage: tools/c5_full_check.py [out.json] [n_total=100000000]"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from flux_b200 import JobConfiguration, synth  # noqa: E402
from flux_b200.worker import GpuContext  # noqa: E402


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else None
    n_total = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000_000
    chunk = 10_000_000
    flat = synth.sphere_cloud_scene(10_000, seed=5).flatten()
    bvh, lin = GpuContext(0), GpuContext(0)
    bvh.set_accel_mode(2)
    lin.set_accel_mode(1)
    bvh.set_scene(flat, JobConfiguration(1))
    lin.set_scene(flat, JobConfiguration(1))
    h_ids, h_t = hashlib.sha256(), hashlib.sha256()
    hits = mism = 0
    ms_bvh = ms_lin = 0.0
    for k in range(n_total // chunk):
        o, d = synth.random_rays(chunk, seed=5, chunk_offset=k)
        hb, tb = bvh.trace_rays(o, d)
        ms_bvh += bvh.last_kernel_ms()
        hl, tl = lin.trace_rays(o, d)
        ms_lin += lin.last_kernel_ms()
        mism += int((hb != hl).sum()) + int((tb.view(np.uint64) != tl.view(np.uint64)).sum())
        hits += int((hb >= 0).sum())
        h_ids.update(hb.tobytes())
        h_t.update(tb.tobytes())
        print(f"chunk {k}: mismatches so far {mism}, bvh {ms_bvh:.1f} ms, linear {ms_lin:.1f} ms", file=sys.stderr, flush=True)
    rays = (n_total // chunk) * chunk
    res = {"config": "c5: 100 M random rays x 10 K spheres", "rays": rays, "bvh_vs_gpu_linear_mismatches": mism,
           "hit_fraction": hits / rays, "sha256_hit_ids": h_ids.hexdigest(), "sha256_t": h_t.hexdigest(),
           "bvh_kernel_ms": ms_bvh, "bvh_Mrays_per_s": rays / (ms_bvh * 1e-3) / 1e6,
           "linear_kernel_ms": ms_lin, "linear_Mrays_per_s": rays / (ms_lin * 1e-3) / 1e6,
           "linear_G_sphere_tests_per_s": rays * 1e4 / (ms_lin * 1e-3) / 1e9}
    print(json.dumps(res), flush=True)
    if out_path:
        with open(out_path, "w") as f:
            json.dump(res, f, indent=1)
    bvh.close()
    lin.close()
    return 0 if mism == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
