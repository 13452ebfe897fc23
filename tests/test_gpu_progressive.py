"""Progressive refinement: passes over sample-index ranges accumulate to Camera::render's image."""
import numpy as np
import pytest

from flux_b200 import JobConfiguration, SceneData, WorkUnit
from flux_b200.worker import Camera, FluxError, GpuContext, Scene
from oracle import oracle_py as O
from tests import helpers as Hp

pytestmark = pytest.mark.gpu


def test_passes_accumulate_to_the_one_shot_render_and_to_the_oracle(gpu_ctx):
    sd = Hp.deterministic_scene(48, 32)
    cfg = JobConfiguration(8, 5, 50)
    flat = sd.flatten()
    ss = Hp.oracle_samples(21, cfg, 48, 32)
    ref = O.render_rows(flat, cfg, ss, 4, 27)
    Hp.upload(gpu_ctx, flat, cfg, ss)
    rows = np.arange(4, 28, dtype=np.uint32)
    try:
        gpu_ctx.set_kernel_mode(1)
        one_shot = gpu_ctx.render_rows(4, 27, 48)
        gpu_ctx.progressive_begin(rows)
        whole = gpu_ctx.progressive_pass(0, 64, 48)
        assert np.array_equal(whole.view(np.uint64), one_shot.view(np.uint64))     # one pass == the direct kernel, bit for bit
        gpu_ctx.progressive_begin(rows)
        errs = []
        for a, b in ((0, 1), (1, 4), (4, 5), (5, 37), (37, 64)):                   # ragged passes, incl. single samples
            img = gpu_ctx.progressive_pass(a, b, 48)
            errs.append(float(np.abs(img - ref).mean()))
        assert Hp.rel_err(img, ref) <= 1e-12                                        # = the oracle up to the order of the sum
        assert errs[-1] < errs[1] < errs[0] and errs[0] > 1e-3                     # it refines
        # the image after the first pass is the 1-sample estimate: that sample's radiance, clamped like max_to_one
        gpu_ctx.progressive_begin(rows)
        first = gpu_ctx.progressive_pass(0, 1, 48)
        assert first.max() <= 1.0 and first.min() >= 0.0 and first.std() > 0.01
        # silent passes (no image wanted) accumulate the same
        gpu_ctx.progressive_begin(rows)
        assert gpu_ctx.progressive_pass(0, 30, 48, want_image=False) is None
        again = gpu_ctx.progressive_pass(30, 64, 48)
        gpu_ctx.progressive_begin(rows)
        gpu_ctx.progressive_pass(0, 30, 48)
        assert np.array_equal(again.view(np.uint64), gpu_ctx.progressive_pass(30, 64, 48).view(np.uint64))
    finally:
        gpu_ctx.set_kernel_mode(0)


def test_progressive_on_glossy_lens_and_mesh_scenes(gpu_ctx):
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    from flux_b200 import synth
    for sd, tol in ((SceneData.from_yaml(os.path.join(root, "scenes", "demo2.yml")).with_size(40, 30), 1e-6),
                    (synth.mesh_scene(24, 12, seed=3, width=40, height=30), 1e-12)):     # BVH traversal in the pass kernel
        cfg = JobConfiguration(6, 5, 50)
        flat = sd.flatten()
        ss = Hp.oracle_samples(5, cfg, 40, 30)
        ref = O.render_rows(flat, cfg, ss, 0, 29)
        gpu_ctx.set_accel_mode(2 if "mesh" in sd.scene_name else 0)
        try:
            Hp.upload(gpu_ctx, flat, cfg, ss)
            gpu_ctx.progressive_begin(np.arange(30, dtype=np.uint32))
            for a in range(0, 36, 7):
                img = gpu_ctx.progressive_pass(a, min(36, a + 7), 40)
            assert Hp.rel_err(img, ref) <= tol, sd.scene_name
        finally:
            gpu_ctx.set_accel_mode(0)


def test_progressive_argument_checks(gpu_ctx):
    sd = Hp.deterministic_scene(16, 8)
    cfg = JobConfiguration(3, 5, 50)
    Hp.upload(gpu_ctx, sd.flatten(), cfg, Hp.oracle_samples(1, cfg, 16, 8))
    with pytest.raises(FluxError, match="flux_progressive_begin not called"):
        gpu_ctx._prog_shape = 1
        gpu_ctx.progressive_pass(0, 1, 16)
    with pytest.raises(FluxError, match="strictly ascending"):
        gpu_ctx.progressive_begin([3, 3])
    with pytest.raises(FluxError, match="row out of range"):
        gpu_ctx.progressive_begin([8])
    gpu_ctx.progressive_begin([0, 2, 5])
    with pytest.raises(FluxError, match="continue where the last one ended"):
        gpu_ctx.progressive_pass(1, 2, 16)
    with pytest.raises(FluxError, match="bad sample range"):
        gpu_ctx.progressive_pass(0, 10, 16)
    gpu_ctx.progressive_pass(0, 9, 16)
    with pytest.raises(FluxError, match="bad sample range"):
        gpu_ctx.progressive_pass(9, 10, 16)
    gpu_ctx.set_scene(sd.flatten(), cfg)                       # a new scene ends the progression
    with pytest.raises(FluxError, match="flux_progressive_begin not called"):
        gpu_ctx.progressive_pass(0, 1, 16)


def test_camera_render_progressive_and_cancel():
    sd = Hp.deterministic_scene(32, 16)
    cfg = JobConfiguration(6, 5, 50)
    scene = Scene.from_data(sd, cfg)
    with GpuContext(0) as ctx:
        cam = Camera.new(scene, cfg, 32, seed=4, ctx=ctx)
        unit = WorkUnit(2, 13)
        ctx.set_kernel_mode(1)
        full = cam.render(scene, unit).rows
        ctx.set_kernel_mode(0)
        steps = list(cam.render_progressive(scene, unit, 10))
        assert [s for s, _ in steps] == [10, 20, 30, 36]
        assert Hp.rel_err(steps[-1][1].rows, full) <= 1e-12
        seen = []
        for done, res in cam.render_progressive(scene, unit, 10, cancel=lambda: len(seen) >= 2):
            seen.append(done)
        assert seen == [10, 20]


def _read_ppm(path):
    tok = open(path).read().split()
    assert tok[0] == "P3"
    w, h, mx = int(tok[1]), int(tok[2]), int(tok[3])
    return np.array(tok[4:], dtype=np.int64).reshape(h, w, 3), mx


def test_cli_progressive_refines_to_the_one_shot_frame(tmp_path):
    """`fluxb200 --progressive K` (the preview's refinement without the window): passes on two row shards (two
    contexts on this GPU) end at the frame the one-shot render writes — the 16-bit PPM values differ by at most one
    unit, and only where the different order of the per-pixel sum crosses a quantisation step."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run(["make", "-C", os.path.join(root, "host"), "-s"], check=True)
    cli, scene = os.path.join(root, "host", "fluxb200"), os.path.join(root, "scenes", "demo2.yml")
    common = ["-r", "6", "--width", "64", "--height", "48", "--seed", "5"]
    a = subprocess.run([cli, scene, *common, "--progressive", "10", "--devices", "0,0", "-o", str(tmp_path / "prog.ppm")],
                       capture_output=True, timeout=300)
    assert a.returncode == 0, a.stderr.decode()
    out = a.stdout.decode()
    assert [int(l.split()[2]) for l in out.splitlines() if l.startswith("pass to")] == [10, 20, 30, 36]
    assert "rendering finished" in out and "cancelled" not in out
    b = subprocess.run([cli, scene, *common, "-o", str(tmp_path / "once.ppm")], capture_output=True, timeout=300)
    assert b.returncode == 0, b.stderr.decode()
    p, mx = _read_ppm(tmp_path / "prog.ppm")
    q, _ = _read_ppm(tmp_path / "once.ppm")
    assert mx == 65535 and p.shape == q.shape == (48, 64, 3)
    diff = np.abs(p - q)
    assert diff.max() <= 1 and (diff != 0).mean() < 0.01
    # one pass over everything is the direct kernel's frame exactly; a first pass alone is a coarser image of it
    c = subprocess.run([cli, scene, *common, "--progressive", "36", "-o", str(tmp_path / "single.ppm")], capture_output=True, timeout=300)
    assert c.returncode == 0 and c.stdout.decode().count("pass to") == 1
    s, _ = _read_ppm(tmp_path / "single.ppm")
    assert np.abs(s - q).max() <= 1
