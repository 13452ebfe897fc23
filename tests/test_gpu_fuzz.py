"""Seeded random scenes: every render kernel and both closest-hit strategies against the oracle.

The hand-made scenes of test_gpu_parity.py exercise the paths the shipped scenes take; these draw sphere / plane
layouts, materials, cameras and lens settings at random (fixed seeds) so that the conservative FP32 box
classification of render_wave2.cu, the BVH's conservative FP32 node tests and the tie rules meet geometry nobody
chose: overlapping and nested spheres, a camera inside a sphere, tiny and huge radii, a zero and a negative radius,
planes through the camera, all four material kinds."""
import numpy as np
import pytest

from flux_b200 import (CameraData, CameraSettings, Emissive, GlossyReflective, JobConfiguration, Matte, OutputSettings,
                       PlaneData, Reflective, SceneData, SphereData, TriangleData)
from oracle import oracle_py as O
from tests import helpers as Hp

pytestmark = pytest.mark.gpu


def _material(rng, allow_glossy):
    k = rng.integers(0, 5 if allow_glossy else 3)
    col = tuple(float(v) for v in rng.uniform(0.2, 1.0, 3))
    if k == 0:
        return Matte(col, (1.0, 1.0, 1.0), float(rng.uniform(0.5, 1.0)))
    if k == 1:
        return Emissive(col, float(rng.uniform(0.5, 6.0)))
    if k == 2:
        return Reflective(float(rng.uniform(0.5, 0.95)), col)
    return GlossyReflective(float(rng.uniform(0.4, 0.9)), col, float(10.0 ** rng.uniform(0.5, 4.0)))


def random_scene(seed, allow_glossy, width=20, height=14):
    rng = np.random.default_rng(seed)
    shapes = []
    if rng.random() < 0.7:   # an environment: the camera is usually inside it
        shapes.append(SphereData((0.0, 0.0, 0.0), float(rng.uniform(30, 120)), Emissive((1.0, 0.97, 0.86), float(rng.uniform(0.2, 1.0))),
                                 True))
    for _ in range(int(rng.integers(2, 20))):
        c = tuple(float(v) for v in rng.uniform(-5, 5, 3))
        r = float(10.0 ** rng.uniform(-1.5, 0.6))
        shapes.append(SphereData(c, r, _material(rng, allow_glossy), bool(rng.random() < 0.15)))
    if rng.random() < 0.5:   # nested / coincident spheres: ties and inside-out hits
        base = shapes[-1]
        shapes.append(SphereData(base.center, base.radius, _material(rng, allow_glossy), False))
        shapes.append(SphereData(base.center, base.radius * 0.5, _material(rng, allow_glossy), True))
    if rng.random() < 0.3:
        shapes.append(SphereData((1.0, 1.0, 1.0), 0.0, _material(rng, allow_glossy), False))       # degenerate
        shapes.append(SphereData((-1.0, 0.5, 2.0), -0.7, _material(rng, allow_glossy), False))     # negative radius: never passes its box
    for _ in range(int(rng.integers(0, 3))):
        nrm = rng.standard_normal(3)
        shapes.append(PlaneData(tuple(float(v) for v in rng.uniform(-3, 3, 3)), tuple(float(v) for v in nrm), _material(rng, allow_glossy)))
    order = rng.permutation(len(shapes))
    shapes = [shapes[i] for i in order]
    eye = tuple(float(v) for v in rng.uniform(-8, 8, 3))
    look = tuple(float(v) for v in rng.uniform(-1, 1, 3))
    cam = CameraData(float(rng.uniform(0.02, 0.2)), 500.0, float(rng.uniform(4, 15)), float(rng.choice([0.0, 0.05, 0.3])))
    return SceneData(f"fuzz{seed}", OutputSettings(width, height, 0.5), tuple(float(v) for v in rng.uniform(0, 0.3, 3)), shapes,
                     CameraSettings(eye, look, (0.0, 1.0, 0.0)), cam)


@pytest.mark.parametrize("seed", range(100, 124))
def test_random_deterministic_scenes_all_kernels(gpu_ctx, seed):
    """No glossy material: nothing transcendental on the path, so every kernel must agree with the oracle to the
    order of the pixel sum (1e-12) and reproduce its event counters exactly."""
    sd = random_scene(seed, allow_glossy=False)
    cfg = JobConfiguration(16, 4, 50)   # 256 spp: direct, regeneration and both wavefront kernels apply
    flat = sd.flatten()
    ss = Hp.oracle_samples(seed, cfg, 20, 14)
    ref, cn_o = O.render_rows(flat, cfg, ss, 0, 13, counters=True)
    Hp.upload(gpu_ctx, flat, cfg, ss)
    try:
        for mode in (1, 2, 4):
            gpu_ctx.set_kernel_mode(mode)
            gpu_ctx.enable_counters(True)
            gpu_ctx.reset_counters()
            img = gpu_ctx.render_rows(0, 13, 20)
            cn = gpu_ctx.counters()
            gpu_ctx.enable_counters(False)
            assert Hp.rel_err(img, ref) <= 1e-12, mode
            for k, v in cn_o.items():
                assert cn[k] == v, (mode, k, cn[k], v)
            again = gpu_ctx.render_rows(0, 13, 20)   # the uninstrumented instantiation computes the same pixels
            assert np.array_equal(again.view(np.uint64), img.view(np.uint64)), mode
    finally:
        gpu_ctx.set_kernel_mode(0)
        gpu_ctx.enable_counters(False)


@pytest.mark.parametrize("seed", range(200, 212))
def test_random_glossy_scenes_all_kernels(gpu_ctx, seed):
    sd = random_scene(seed, allow_glossy=True)
    cfg = JobConfiguration(16, 5, 50)
    flat = sd.flatten()
    ss = Hp.oracle_samples(seed, cfg, 20, 14)
    ref = O.render_rows(flat, cfg, ss, 0, 13)
    Hp.upload(gpu_ctx, flat, cfg, ss)
    try:
        for mode in (1, 2, 4):
            gpu_ctx.set_kernel_mode(mode)
            img = gpu_ctx.render_rows(0, 13, 20)
            assert Hp.rel_err(img, ref) <= 1e-6, mode
    finally:
        gpu_ctx.set_kernel_mode(0)


@pytest.mark.parametrize("seed", range(300, 308))
def test_random_scenes_bvh_equals_linear_and_oracle(gpu_ctx, seed):
    """Random mixed scenes (spheres of wildly different sizes, duplicates, slivers and degenerate triangles, planes):
    BVH = GPU linear scan bitwise on 200 K rays incl. axis-parallel ones; = oracle on a prefix."""
    rng = np.random.default_rng(seed)
    m = Matte((0.5, 0.5, 0.5), (1, 1, 1), 1.0)
    shapes = []
    for _ in range(int(rng.integers(50, 400))):
        shapes.append(SphereData(tuple(float(v) for v in rng.uniform(-10, 10, 3)), float(10.0 ** rng.uniform(-2, 0.7)), m, bool(rng.random() < 0.1)))
    shapes += shapes[:20]                                                # exact duplicates with later ids
    shapes.append(SphereData((0.0, 0.0, 0.0), 60.0, m, True))          # environment: linear list
    for _ in range(int(rng.integers(50, 300))):
        a = rng.uniform(-10, 10, 3)
        b = a + rng.standard_normal(3) * 10.0 ** rng.uniform(-3, 0.5)
        c = a + rng.standard_normal(3) * 10.0 ** rng.uniform(-3, 0.5)
        shapes.append(TriangleData(tuple(map(float, a)), tuple(map(float, b)), tuple(map(float, c)), m))
    shapes.append(TriangleData((1.0, 1.0, 1.0), (1.0, 1.0, 1.0), (2.0, 2.0, 2.0), m))   # zero area
    shapes.append(PlaneData((0.0, -9.0, 0.0), (0.0, 1.0, 0.0), m))
    shapes = [shapes[i] for i in rng.permutation(len(shapes))]
    base = Hp.deterministic_scene()
    flat = SceneData("fuzzbvh", base.output_settings, (0, 0, 0), shapes, base.camera_settings, base.camera_data).flatten()
    o, d = Hp.random_rays(rng, 200_000, extent=12.0)
    d[:3000, int(rng.integers(0, 3))] = 0.0
    d[3000:4000] = np.eye(3)[rng.integers(0, 3, 1000)] * rng.choice([-1.0, 1.0], (1000, 1))
    res = {}
    try:
        for mode in (1, 2):
            gpu_ctx.set_accel_mode(mode)
            gpu_ctx.set_scene(flat, JobConfiguration(1))
            res[mode] = gpu_ctx.trace_rays(o, d)
    finally:
        gpu_ctx.set_accel_mode(0)
    assert np.array_equal(res[1][0], res[2][0])
    assert np.array_equal(res[1][1].view(np.uint64), res[2][1].view(np.uint64))
    ho, to = O.trace_rays(flat, o[:20_000], d[:20_000])
    assert np.array_equal(res[2][0][:20_000], ho)
    assert np.array_equal(res[2][1][:20_000].view(np.uint64), to.view(np.uint64))


@pytest.mark.parametrize("seed", range(400, 408))
def test_random_scenes_with_triangles_boxes_and_rectangles_render_like_the_oracle(gpu_ctx, seed):
    """Random mixed scenes — spheres, planes, loose triangles, boxes, rectangles, a small mesh, all four material
    kinds — rendered through the BVH by every kernel that takes triangles, and by the linear-scan direct kernel."""
    from flux_b200 import BoxData, MeshData, RectangleData
    rng = np.random.default_rng(seed)
    sd0 = random_scene(seed, allow_glossy=True)
    shapes = list(sd0.shapes)
    for _ in range(int(rng.integers(3, 30))):
        a = rng.uniform(-5, 5, 3)
        shapes.append(TriangleData(tuple(map(float, a)), tuple(map(float, a + rng.standard_normal(3) * 10.0 ** rng.uniform(-1, 0.5))),
                                   tuple(map(float, a + rng.standard_normal(3) * 10.0 ** rng.uniform(-1, 0.5))), _material(rng, True)))
    for _ in range(int(rng.integers(1, 4))):
        lo = rng.uniform(-4, 3, 3)
        shapes.append(BoxData(tuple(map(float, lo)), tuple(map(float, lo + rng.uniform(0.2, 2.0, 3))), _material(rng, True)))
    for _ in range(int(rng.integers(1, 3))):
        c = rng.uniform(-4, 4, 3)
        shapes.append(RectangleData(tuple(map(float, c)), tuple(map(float, rng.standard_normal(3) * 2)), tuple(map(float, rng.standard_normal(3) * 2)),
                                    _material(rng, True)))
    verts = rng.uniform(-3, 3, (12, 3))
    faces = np.array([rng.choice(12, 3, replace=False) for _ in range(10)])
    shapes.append(MeshData(verts, faces, _material(rng, True)))
    shapes = [shapes[i] for i in rng.permutation(len(shapes))]
    sd = SceneData(f"fuzzmix{seed}", sd0.output_settings, sd0.background, shapes, sd0.camera_settings, sd0.camera_data)
    cfg = JobConfiguration(16, 5, 50)
    flat = sd.flatten()
    ss = Hp.oracle_samples(seed, cfg, 20, 14)
    ref = O.render_rows(flat, cfg, ss, 0, 13)
    try:
        for accel, modes in ((2, (1, 2, 4)), (1, (1,))):
            gpu_ctx.set_accel_mode(accel)
            Hp.upload(gpu_ctx, flat, cfg, ss)
            for mode in modes:
                gpu_ctx.set_kernel_mode(mode)
                img = gpu_ctx.render_rows(0, 13, 20)
                assert Hp.rel_err(img, ref) <= 1e-6, (accel, mode)
    finally:
        gpu_ctx.set_kernel_mode(0)
        gpu_ctx.set_accel_mode(0)


@pytest.mark.parametrize("seed", range(500, 540))
def test_primary_ray_mask_never_changes_a_pixel(gpu_ctx, seed):
    """render_wave2.cu classifies, for warps of camera rays, only the boxes its per-pixel bundle test lets through
    (primary_may_hit).  The instrumented instantiation never uses that mask (it counts every box test, and its counts
    and pixels are held to the oracle above), so the two must agree bit for bit — on cameras chosen to stress the
    bundle bound: wide lenses, short and long focal distances, strong zoom, the eye inside or next to spheres, spheres
    straddling the lens plane and behind it, huge and tiny radii."""
    rng = np.random.default_rng(seed)
    sd = random_scene(seed, allow_glossy=bool(seed & 1), width=24, height=16)
    shapes = list(sd.shapes)
    eye = np.array(sd.camera_settings.eye)
    for _ in range(int(rng.integers(2, 10))):     # spheres around the eye: some contain it, some touch the lens disc
        c = eye + rng.standard_normal(3) * 10.0 ** rng.uniform(-1.5, 0.5)
        shapes.append(SphereData(tuple(map(float, c)), float(10.0 ** rng.uniform(-2.0, 0.5)), _material(rng, bool(seed & 1)),
                                 bool(rng.random() < 0.2)))
    cam = CameraData(float(10.0 ** rng.uniform(-2.0, -0.3)), 500.0, float(10.0 ** rng.uniform(-0.5, 1.5)),
                     float(rng.choice([0.0, 0.01, 0.3, 2.0])))
    sd = SceneData(sd.scene_name, sd.output_settings, sd.background, shapes, sd.camera_settings, cam)
    cfg = JobConfiguration(16, 4, 50)
    gpu_ctx.set_kernel_mode(4)
    try:
        gpu_ctx.set_scene(sd.flatten(), cfg)
        gpu_ctx.generate_samples(seed, 24)
        plain = gpu_ctx.render_rows(0, 15, 24)
        gpu_ctx.enable_counters(True)
        counted = gpu_ctx.render_rows(0, 15, 24)
    finally:
        gpu_ctx.enable_counters(False)
        gpu_ctx.set_kernel_mode(0)
    assert np.array_equal(plain.view(np.uint64), counted.view(np.uint64))
