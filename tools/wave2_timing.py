"""Per-stage cycle accounting of render_wave2.cu (needs a library built with -DWAVE2_TIMING):
  python flux_b200/build.py -o flux_b200/lib/variants/lib_timing.so -DWAVE2_TIMING
  FLUXB200_LIB=$PWD/flux_b200/lib/variants/lib_timing.so python tools/wave2_timing.py"""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
from flux_b200 import JobConfiguration, SceneData
from flux_b200.worker import GpuContext
sd = SceneData.from_yaml("scenes/demo2.yml")
cfg = JobConfiguration(64, 5, 50)
ctx = GpuContext(0)
ctx.set_scene(sd.flatten(), cfg)
ctx.generate_samples(1, 800)
rows = np.arange(0, 600, 4, dtype=np.uint32)
ctx.render_row_list(rows, 800)
ctx.reset_counters()
ctx.render_row_list(rows, 800)
ms = ctx.last_kernel_ms()
c = ctx.counters()
v = list(c.values())[:15]
names = ["stage1 owner", "wait barrier 1", "bin (scan+list)", "wait barrier 2", "stage3 item", "wait barrier 3"]
it = v[6]
tot = sum(v[:6])
print(f"kernel {ms:.1f} ms; warp-iterations {it}; cycles per warp-iteration {tot / it:.0f}")
for n, x in zip(names, v[:6]):
    print(f"  {n:18s} {x / it:8.0f} cycles  {100 * x / tot:5.1f} %")
for name, k in (("matte warp", 7), ("glossy/spec warp", 9), ("term+regen warp", 11), ("idle warp", 13)):
    if v[k + 1]:
        print(f"  item stage, {name:18s} {v[k] / v[k + 1]:8.0f} cycles  ({100 * v[k + 1] / it:4.1f} % of warp-iterations)")
