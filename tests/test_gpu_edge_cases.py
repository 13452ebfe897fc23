"""Edge cases of the render path against the oracle: empty and one-shape scenes, 1-pixel and 1-row images, single
samples, NaN pixels, colours above one, zero rows, and the largest sizes — at sizes the oracle cannot render, through
properties that do not depend on the size (sharding, row subsets, pass accumulation)."""
import numpy as np
import pytest

from flux_b200 import (CameraData, CameraSettings, Emissive, JobConfiguration, Matte, OutputSettings, PlaneData, Reflective,
                       SceneData, SphereData)
from flux_b200.worker import FluxError
from oracle import oracle_py as O
from tests import helpers as Hp

pytestmark = pytest.mark.gpu


def _scene(shapes, width, height, background=(0.1, 0.2, 0.3), eye=(0.0, 3.0, -9.0), lens=0.0):
    return SceneData("edge", OutputSettings(width, height, 0.5), background, shapes,
                     CameraSettings(eye, (0.0, 1.0, 0.0), (0.0, 1.0, 0.0)), CameraData(6.0 * width / 800.0, 500.0, 10.0, lens))


def _modes_for(cfg):
    spp = cfg.sample_root ** 2
    return [m for m, need in ((1, 1), (2, 64), (4, 256)) if spp >= need]


def _check_all_modes(gpu_ctx, sd, cfg, seed=3, tol=1e-12):
    w, h = sd.output_settings.image_width, sd.output_settings.image_height
    flat = sd.flatten()
    ss = Hp.oracle_samples(seed, cfg, w, h)
    ref = O.render_rows(flat, cfg, ss, 0, h - 1)
    Hp.upload(gpu_ctx, flat, cfg, ss)
    try:
        for mode in [0] + _modes_for(cfg):
            gpu_ctx.set_kernel_mode(mode)
            img = gpu_ctx.render_rows(0, h - 1, w)
            assert Hp.rel_err(img, ref) <= tol, mode
    finally:
        gpu_ctx.set_kernel_mode(0)
    return ref


@pytest.mark.parametrize("root", [1, 8, 16])
def test_scene_without_shapes_is_the_background(gpu_ctx, root):
    ref = _check_all_modes(gpu_ctx, _scene([], 12, 9), JobConfiguration(root, 5, 50))
    assert np.allclose(ref, [0.1, 0.2, 0.3], rtol=1e-12)


@pytest.mark.parametrize("root", [1, 16])
def test_background_above_one_is_scaled_by_max_to_one(gpu_ctx, root):
    ref = _check_all_modes(gpu_ctx, _scene([], 5, 4, background=(4.0, 2.0, 1.0)), JobConfiguration(root, 5, 50))
    assert np.allclose(ref, [1.0, 0.5, 0.25], rtol=1e-12)       # color.rs:35-44


@pytest.mark.parametrize("width,height", [(1, 1), (1, 7), (9, 1), (33, 2)])
def test_tiny_images_every_kernel(gpu_ctx, width, height):
    shapes = [SphereData((0, 0, 0), 100.0, Emissive((1, 0.9, 0.8), 0.6), True),
              SphereData((0.0, 1.0, 0.0), 1.5, Matte((0.6, 0.6, 0.6), (1, 1, 1), 1.0), False),
              PlaneData((0, 0, 0), (0, 1, 0), Reflective(0.8, (0.9, 0.9, 0.9)))]
    _check_all_modes(gpu_ctx, _scene(shapes, width, height, lens=0.1), JobConfiguration(16, 5, 50))


def test_only_planes_and_only_one_sphere(gpu_ctx):
    plane = PlaneData((0, 0, 0), (0, 1, 0), Matte((0.5, 0.5, 0.5), (1, 1, 1), 1.0))
    _check_all_modes(gpu_ctx, _scene([plane, PlaneData((0, 9, 0), (0, -1, 0), Emissive((1, 1, 1), 2.0))], 16, 12), JobConfiguration(16, 5, 50))
    _check_all_modes(gpu_ctx, _scene([SphereData((0, 1, 0), 2.0, Emissive((1, 0.5, 0.2), 3.0), False)], 16, 12), JobConfiguration(16, 5, 50))


def test_nan_and_infinite_radiance_coincide_with_the_oracle(gpu_ctx):
    """An emitter of infinite power gives inf radiance, which max_to_one turns into inf * (1 / inf) = NaN
    (color.rs:35-44); a NaN colour component poisons its channel only.  Whatever the reference's arithmetic yields,
    the GPU yields the same pattern (rel_err requires the NaN masks to coincide).  Degenerate shapes ride along: a
    zero-radius sphere (its box is never passed: t0 < t1 fails) and a plane with a null normal (t = 0/0 is no hit)."""
    shapes = [SphereData((0, 0, 0), 50.0, Emissive((1, 1, 1), 0.5), True),
              SphereData((-1.5, 1.0, 0.0), 0.8, Emissive((1.0, 0.5, 0.25), float("inf")), False),
              SphereData((1.5, 1.0, 0.0), 0.8, Emissive((0.5, float("nan"), 0.25), 2.0), False),
              SphereData((0.0, 1.0, 0.0), 0.0, Matte((0.6, 0.6, 0.6), (1, 1, 1), 1.0), False),
              PlaneData((0, 5, 0), (0, 0, 0), Matte((0.5, 0.5, 0.5), (1, 1, 1), 1.0)),
              PlaneData((0, 0, 0), (0, 1, 0), Matte((0.5, 0.5, 0.5), (1, 1, 1), 1.0))]
    sd = _scene(shapes, 24, 16)
    ref = _check_all_modes(gpu_ctx, sd, JobConfiguration(16, 5, 50))
    nan = np.isnan(ref)
    assert nan.any() and not nan.all()
    assert (nan[..., 1] & ~nan[..., 0]).any()        # the NaN green channel alone, somewhere on the right sphere


def test_zero_rows_and_single_rows(gpu_ctx):
    sd = Hp.deterministic_scene(32, 16)
    cfg = JobConfiguration(4, 5, 50)
    Hp.upload(gpu_ctx, sd.flatten(), cfg, Hp.oracle_samples(2, cfg, 32, 16))
    assert gpu_ctx.render_row_list(np.zeros(0, np.uint32), 32).shape == (0, 32, 3)
    full = gpu_ctx.render_rows(0, 15, 32)
    for r in (0, 7, 15):
        assert np.array_equal(gpu_ctx.render_rows(r, r, 32).view(np.uint64), full[r:r + 1].view(np.uint64))
    picked = gpu_ctx.render_row_list(np.array([1, 2, 9, 15], np.uint32), 32)
    assert np.array_equal(picked.view(np.uint64), full[[1, 2, 9, 15]].view(np.uint64))


def test_largest_sample_count_on_a_few_pixels(gpu_ctx):
    """sample_root 160 (25 600 spp, beyond BASELINE config 2's 16 384) on a 4x3 image: the wavefront kernel against
    the oracle, which still finishes in seconds at this pixel count."""
    sd = Hp.deterministic_scene(4, 3)
    cfg = JobConfiguration(160, 5, 50)
    flat = sd.flatten()
    ss = Hp.oracle_samples(8, cfg, 4, 3)
    ref = O.render_rows(flat, cfg, ss, 0, 2)
    Hp.upload(gpu_ctx, flat, cfg, ss)
    img = gpu_ctx.render_rows(0, 2, 4)
    assert Hp.rel_err(img, ref) <= 1e-11


def test_largest_image_through_size_independent_properties(gpu_ctx):
    """4096 x 2304 (9.4 M pixels, 85 M samples at 9 spp; the oracle would need minutes): rows rendered alone, in
    shards of interleaved tiles and as progressive passes must be the very pixels of the full render; two whole rows
    from the middle of the frame are checked against the oracle as well."""
    from flux_b200.worker import shard_rows
    w, h = 4096, 2304
    sd = Hp.deterministic_scene(w, h)
    cfg = JobConfiguration(3, 5, 50)
    gpu_ctx.set_scene(sd.flatten(), cfg)
    gpu_ctx.generate_samples(11, w)
    full = gpu_ctx.render_rows(0, h - 1, w)
    assert np.isfinite(full).all() and 0.0 <= full.min() and full.max() <= 1.0 and full.std() > 0.05
    rng = np.random.default_rng(0)
    rows = np.sort(rng.choice(h, 40, replace=False)).astype(np.uint32)
    assert np.array_equal(gpu_ctx.render_row_list(rows, w).view(np.uint64), full[rows].view(np.uint64))
    parts = np.empty_like(full)
    for rank in range(3):
        mine = shard_rows(h, 4, rank, 3)
        parts[mine] = gpu_ctx.render_row_list(mine, w)
    assert np.array_equal(parts.view(np.uint64), full.view(np.uint64))
    gpu_ctx.set_kernel_mode(1)
    try:
        band = gpu_ctx.render_rows(1000, 1099, w)
        gpu_ctx.progressive_begin(np.arange(1000, 1100, dtype=np.uint32))
        gpu_ctx.progressive_pass(0, 4, w, want_image=False)
        assert Hp.rel_err(gpu_ctx.progressive_pass(4, 9, w), band) <= 1e-12
    finally:
        gpu_ctx.set_kernel_mode(0)
    ss_o = O.generate_samples(11, 3, 5, w)
    ss_o.set_index = O.generate_set_index(11, h, w, w)
    ref = O.render_rows(sd.flatten(), cfg, ss_o, 1152, 1153)
    assert Hp.rel_err(full[1152:1154], ref) <= 1e-12


def test_triangles_only_many_planes_and_120_spheres(gpu_ctx):
    """Shape mixes at the edges of the kernel-selection rules: no sphere at all (triangles only: the BVH's linear
    list is empty), forty planes (every one is tested by every ray), and 120 spheres (beyond the 112 the wavefront
    kernel scans linearly: its BVH owner stage)."""
    from flux_b200 import BoxData, RectangleData
    rng = np.random.default_rng(12)
    light = Emissive((1.0, 0.95, 0.9), 4.0)
    grey = Matte((0.6, 0.6, 0.6), (1, 1, 1), 1.0)
    tri_only = _scene([RectangleData((-6, 0, -6), (0, 0, 12), (12, 0, 0), grey), RectangleData((-2, 4, -2), (4, 0, 0), (0, 0, 4), light),
                       BoxData((-1.0, 0.0, -0.5), (0.5, 1.5, 1.0), Reflective(0.8, (0.9, 0.9, 0.9)))], 20, 14)
    planes = [PlaneData(tuple(map(float, rng.uniform(-6, 6, 3))), tuple(map(float, rng.standard_normal(3))),
                        Matte(tuple(map(float, rng.uniform(0.3, 0.9, 3))), (1, 1, 1), 1.0)) for _ in range(39)]
    many_planes = _scene(planes + [PlaneData((0, 9, 0), (0, -1, 0), light)], 20, 14)
    spheres = [SphereData((0, 0, 0), 60.0, Emissive((1, 1, 1), 0.4), True)]
    spheres += [SphereData(tuple(map(float, rng.uniform(-5, 5, 3))), float(rng.uniform(0.2, 0.7)),
                           Matte(tuple(map(float, rng.uniform(0.3, 0.9, 3))), (1, 1, 1), 1.0) if k % 2 else Reflective(0.8, (0.9, 0.9, 0.9)), False)
                for k in range(119)]
    crowd = _scene(spheres, 20, 14)
    cfg = JobConfiguration(16, 5, 50)
    taken = []
    for sd in (tri_only, many_planes, crowd):
        w, h = sd.output_settings.image_width, sd.output_settings.image_height
        flat = sd.flatten()
        ss = Hp.oracle_samples(9, cfg, w, h)
        ref = O.render_rows(flat, cfg, ss, 0, h - 1)
        Hp.upload(gpu_ctx, flat, cfg, ss)
        try:
            imgs = {}
            for mode in (0, 1, 2, 4):
                gpu_ctx.set_kernel_mode(mode)
                try:
                    imgs[mode] = gpu_ctx.render_rows(0, h - 1, w)
                except FluxError as e:      # a forced kernel that does not take this scene says so; auto and direct always do
                    assert mode in (2, 4) and "needs" in str(e), (len(sd.shapes), mode, str(e))
                    continue
                assert Hp.rel_err(imgs[mode], ref) <= 1e-12, (len(sd.shapes), mode)
            assert 0 in imgs and 1 in imgs
            taken.append(sorted(imgs))
        finally:
            gpu_ctx.set_kernel_mode(0)
    # a few triangles without a BVH leave only the direct kernel; the 120-sphere scene (BVH) is taken by every kernel
    assert taken[0] == [0, 1] and taken[2] == [0, 1, 2, 4], taken


def test_stale_set_index_map_is_refused(gpu_ctx):
    """A set-index map names sample sets: it must not outlive the image size or the number of sets it was checked
    against.  set_scene(root 2) + 800 sets + map, then set_scene(root 8) + 4 sets + render used to read sets 4..799
    of allocations that hold 4; it is a call-order error now."""
    shapes = [SphereData((0, 0, 0), 100.0, Emissive((1, 0.9, 0.8), 0.6), True)]
    sd = _scene(shapes, 40, 30)
    flat = sd.flatten()
    cfg2, cfg8 = JobConfiguration(2, 5, 50), JobConfiguration(8, 5, 50)
    ss2 = Hp.oracle_samples(1, cfg2, 40, 30, num_sets=800)
    assert ss2.set_index.max() >= 4
    Hp.upload(gpu_ctx, flat, cfg2, ss2)
    gpu_ctx.render_rows(0, 29, 40)
    gpu_ctx.set_scene(flat, cfg8)
    ss8 = Hp.oracle_samples(1, cfg8, 40, 30, num_sets=4)
    gpu_ctx.set_samples(ss8.root, ss8.max_depth, 4, ss8.pixel, ss8.disc, ss8.hemi)
    with pytest.raises(FluxError) as e:
        gpu_ctx.render_rows(0, 29, 40)
    assert e.value.code == 3   # FLUX_ERR_STATE
    # same root, fewer sets: the map is stale as well
    Hp.upload(gpu_ctx, flat, cfg2, ss2)
    ss2b = Hp.oracle_samples(1, cfg2, 40, 30, num_sets=4)
    gpu_ctx.set_samples(2, 5, 4, ss2b.pixel, ss2b.disc, ss2b.hemi)
    with pytest.raises(FluxError) as e:
        gpu_ctx.render_rows(0, 29, 40)
    assert e.value.code == 3
    # another image size that still fits the old allocation
    Hp.upload(gpu_ctx, flat, cfg2, ss2)
    small = _scene(shapes, 20, 10).flatten()
    gpu_ctx.set_scene(small, cfg2)
    with pytest.raises(FluxError) as e:
        gpu_ctx.render_rows(0, 9, 20)
    assert e.value.code == 3
    # and the proper sequence still renders
    gpu_ctx.set_set_index(np.zeros((10, 20), np.uint32))
    img = gpu_ctx.render_rows(0, 9, 20)
    assert np.isfinite(img).all()


def test_caller_samples_outside_the_unit_square_switch_the_primary_mask_off(gpu_ctx, demo2):
    """The per-pixel primary-ray mask of the wavefront kernel is derived from unit-square pixel samples and unit-disc
    lens samples.  Caller-supplied sets may be anything (flux_set_samples): with samples three pixels wide and a lens
    disc of radius 2 the camera rays leave the pixel's bundle, the mask must stand aside, and the image still equals the
    oracle's on the same sets."""
    sd = demo2.with_size(20, 14)
    cfg = JobConfiguration(16, 5, 50)
    flat = sd.flatten()
    ss = Hp.oracle_samples(77, cfg, 20, 14)
    ss.pixel = ss.pixel * 3.0 - 1.0      # [-1, 2): three pixels wide
    ss.disc = ss.disc * 2.0              # lens samples out to radius 2
    Hp.upload(gpu_ctx, flat, cfg, ss)
    try:
        gpu_ctx.set_kernel_mode(4)
        img = gpu_ctx.render_rows(0, 13, 20)
        gpu_ctx.enable_counters(True)
        counted = gpu_ctx.render_rows(0, 13, 20)
    finally:
        gpu_ctx.enable_counters(False)
        gpu_ctx.set_kernel_mode(0)
    ref = O.render_rows(flat, cfg, ss, 0, 13)
    assert Hp.rel_err(img, ref) <= 1e-6
    assert np.array_equal(img.view(np.uint64), counted.view(np.uint64))
