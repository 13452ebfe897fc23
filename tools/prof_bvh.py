"""Workload for an ncu capture of the BVH kernels (traversal is reported with L1/L2 hit rates and
pipe utilisation): 10 M config-5 rays through trace_rays_kernel<true>, then a 200x150 crop-size render of the
config-3 mesh at 1024 spp through render_regen_kernel<false, true>.
  ncu --set full --clock-control none -k regex:'trace_rays|render_regen' -o gpurun_out/bvh python tools/prof_bvh.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flux_b200 import JobConfiguration, synth  # noqa: E402
from flux_b200.worker import GpuContext  # noqa: E402

ctx = GpuContext(0)
ctx.set_scene(synth.sphere_cloud_scene(10_000, seed=5).flatten(), JobConfiguration(1))
o, d = synth.random_rays(10_000_000, seed=5)
hit, t = ctx.trace_rays(o, d)
print("config 5, 10 M rays:", ctx.last_kernel_ms(), "ms", float((hit >= 0).mean()))
sd = synth.mesh_scene(1000, 500, seed=3, width=200, height=150)
ctx.set_scene(sd.flatten(), JobConfiguration(32, 5, 50))
ctx.generate_samples(1, 200)
ctx.render_rows(0, 149, 200)
print("config 3 mesh 200x150 @1024spp:", ctx.last_kernel_ms(), "ms")
ctx.close()
