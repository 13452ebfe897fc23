"""Tiny renders through every kernel — the workload for `compute-sanitizer --tool memcheck|racecheck python
tools/sanitize_small.py` where compute-sanitizer is available (it is closed on the pool used in round 1; the
shared-memory hand-offs of the wavefront kernels are argued barrier by barrier in DESIGN.md instead)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from flux_b200 import JobConfiguration, SceneData, synth  # noqa: E402
from flux_b200.worker import GpuContext  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ctx = GpuContext(0)
sd = SceneData.from_yaml(os.path.join(ROOT, "scenes", "demo2.yml")).with_size(6, 4)
for mode, root in ((1, 2), (2, 8), (3, 64), (4, 16), (4, 64)):
    cfg = JobConfiguration(root, 5, 50)
    ctx.set_kernel_mode(mode)
    ctx.set_scene(sd.flatten(), cfg)
    ctx.generate_samples(1, 6)
    img = ctx.render_rows(0, 3, 6)
    print("mode", mode, "root", root, "mean", float(np.nanmean(img)))
ctx.set_kernel_mode(0)
mesh = synth.mesh_scene(12, 8, seed=3, width=6, height=4)
for root in (2, 8):
    ctx.set_scene(mesh.flatten(), JobConfiguration(root, 5, 50))
    ctx.generate_samples(2, 6)
    print("mesh root", root, "mean", float(np.nanmean(ctx.render_rows(0, 3, 6))))
o = np.random.default_rng(0).uniform(-5, 5, (2000, 3)); d = np.random.default_rng(1).standard_normal((2000, 3))
hit, t = ctx.trace_rays(o, d)
print("trace", int((hit >= 0).sum()))
ctx.close()
