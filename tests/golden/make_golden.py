#!/usr/bin/env python
"""Generates the golden fixtures in this directory.

The reference (Rust) cannot be built or run in this environment and ships no test vectors, so the
fixtures are of two kinds:
  * demo2_reference.png — the reference's own converged render of scenes/demo2.yml at 16384 spp
    (copied verbatim from the reference repository's demo.png; README.md:1-3).  It is the only
    reference-produced artefact and pins the oracle statistically (tests/test_golden.py).
  * oracle_*.npz — outputs of THIS repo's CPU oracle on seeded inputs, committed so that a change in
    the oracle's arithmetic is caught (regression vectors, not reference vectors), and so that the
    GPU path can be checked against fixed files as well as against a live oracle run.

usage: python tests/golden/make_golden.py   (rewrites oracle_*.npz)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from flux_b200 import JobConfiguration, SceneData  # noqa: E402
from oracle import oracle_py as O  # noqa: E402
from tests import helpers as Hp  # noqa: E402

CASES = {
    # name: (scene factory, width, height, root, depth, seed, rows)
    "oracle_demo1_64x48_r3": (lambda: SceneData.from_yaml(os.path.join(ROOT, "scenes", "demo1.yml")).with_size(64, 48), 3, 5, 101),
    "oracle_demo2_rows_r2": (lambda: SceneData.from_yaml(os.path.join(ROOT, "scenes", "demo2.yml")), 2, 5, 102),
    "oracle_mixed_96x64_r4": (lambda: Hp.mixed_material_scene(), 4, 7, 103),
    "oracle_deterministic_96x64_r2": (lambda: Hp.deterministic_scene(), 2, 6, 104),
}


def make(name):
    factory, root, depth, seed = CASES[name]
    sd = factory()
    W, H = sd.output_settings.image_width, sd.output_settings.image_height
    cfg = JobConfiguration(root, depth, 50)
    rows = np.arange(0, H, 40, dtype=np.uint32) if H > 100 else np.arange(H, dtype=np.uint32)
    ss = Hp.oracle_samples(seed, cfg, W, H)
    img, cn = O.render_row_list(sd.flatten(), cfg, ss, rows, counters=True)
    return dict(rows=rows, image=img, seed=np.uint64(seed), root=np.uint32(root), depth=np.uint32(depth),
                counters=np.array([cn[k] for k in sorted(cn)], np.uint64), counter_names=np.array(sorted(cn)))


if __name__ == "__main__":
    for name in CASES:
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **make(name))
        print("wrote", name)
    # ray-batch fixture: demo2 shapes, seeded rays, oracle hit ids and distances
    rng = np.random.default_rng(105)
    o, d = Hp.random_rays(rng, 4096, extent=12.0)
    o[:, 1] = np.abs(o[:, 1])
    flat = SceneData.from_yaml(os.path.join(ROOT, "scenes", "demo2.yml")).flatten()
    hit, t = O.trace_rays(flat, o, d)
    np.savez_compressed(os.path.join(HERE, "oracle_rays_demo2.npz"), origins=o, dirs=d, hit=hit, t=t)
    print("wrote oracle_rays_demo2")
