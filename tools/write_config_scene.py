#!/usr/bin/env python
"""Write the synthetic BASELINE configs as scene files the render driver takes ("inputs ... written to a
file both sides read").  usage: tools/write_config_scene.py c3|c4|c5 OUT.yml
  c3: the 1 M-triangle height-field mesh, 800x600 (61 MB of YAML; host/fluxb200 loads it in about 2 s)  -> -r 32
  c4: the 67-sphere glossy scene, 1920x1080                                                             -> -r 64
  c5: the 10 K-sphere cloud (as a scene; its 100 M-ray batch is tools/bench_configs.py c5)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flux_b200 import synth  # noqa: E402

MAKE = {"c3": lambda: synth.mesh_scene(1000, 500, seed=3), "c4": synth.glossy_scene, "c5": lambda: synth.sphere_cloud_scene(10_000, seed=5)}

if __name__ == "__main__":
    if len(sys.argv) != 3 or sys.argv[1] not in MAKE:
        sys.exit(__doc__)
    MAKE[sys.argv[1]]().to_yaml(sys.argv[2])
    print(f"wrote {sys.argv[2]} ({os.path.getsize(sys.argv[2]) / 1e6:.1f} MB)")
