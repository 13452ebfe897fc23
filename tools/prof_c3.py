"""Config-3 mesh crop (200x150 @1024spp) alone, for a light ncu capture of the BVH render kernel."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flux_b200 import JobConfiguration, synth  # noqa: E402
from flux_b200.worker import GpuContext  # noqa: E402

ctx = GpuContext(0)
sd = synth.mesh_scene(1000, 500, seed=3, width=200, height=150)
ctx.set_scene(sd.flatten(), JobConfiguration(32, 5, 50))
ctx.generate_samples(1, 200)
ctx.render_rows(0, 149, 200)
print("config 3 mesh 200x150 @1024spp:", ctx.last_kernel_ms(), "ms")
ctx.close()
