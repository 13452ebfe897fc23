"""Render scenes/demo2.yml at 16384 spp on cuda:0 and compare with the reference's own render (demo.png of the
reference repository, copied to tests/golden/demo2_reference.png): writes gpurun_out/demo2_16384spp.png and prints
the RMSE (linear 8-bit values, as the reference's PNG stores them)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from PIL import Image  # noqa: E402

from flux_b200 import JobConfiguration, SceneData  # noqa: E402
from flux_b200.worker import GpuWorker  # noqa: E402

sd = SceneData.from_yaml(os.path.join(ROOT, "scenes", "demo2.yml"))
w = GpuWorker(0, seed=1)
img = w.render_image(sd, JobConfiguration(128, 5, 50))
w.stop()
ref = np.asarray(Image.open(os.path.join(ROOT, "tests", "golden", "demo2_reference.png")).convert("RGB"), np.float64) / 255.0
q = np.clip(np.nan_to_num(img), 0.0, 1.0)
rmse = float(np.sqrt(np.mean((q - ref) ** 2)))
print(f"demo2 @16384 spp vs the reference's demo.png: RMSE {rmse:.5f}; channel means ours {q.mean(axis=(0, 1)).round(4)} "
      f"reference {ref.mean(axis=(0, 1)).round(4)}")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
Image.fromarray((q * 255.0 + 0.5).astype(np.uint8)).save(os.path.join(ROOT, "gpurun_out", "demo2_16384spp.png"))
