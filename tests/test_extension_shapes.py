"""Rectangle and Box (EXTENSION shapes; the reference's TODO.md:2 asks for a "Quad (for area
light)"): their decomposition into triangles on the host, and renders of a closed room lit by a rectangle against the
oracle through every kernel that takes triangles."""
import numpy as np
import pytest

from flux_b200 import (BoxData, CameraData, CameraSettings, Emissive, GlossyReflective, JobConfiguration, Matte,
                       OutputSettings, RectangleData, Reflective, SceneData, SphereData)
from oracle import oracle_py as O
from tests import helpers as Hp

GREY = Matte((0.7, 0.7, 0.7), (1, 1, 1), 1.0)


def _normal(tri):
    a, b, c = (np.asarray(v, np.float64) for v in tri)
    return np.cross(b - a, c - a)


def test_rectangle_is_two_triangles_wound_like_edge_a_cross_edge_b():
    r = RectangleData((1.0, 2.0, 3.0), (2.0, 0.0, 0.5), (0.0, 1.5, 0.25), GREY)
    tris = r.triangles()
    want = np.cross(r.edge_a, r.edge_b)
    assert len(tris) == 2
    for t in tris:
        n = _normal(t)
        assert np.allclose(np.cross(n, want), 0.0) and np.dot(n, want) > 0           # same direction
    assert np.isclose(sum(np.linalg.norm(_normal(t)) / 2 for t in tris), np.linalg.norm(want))   # areas add up
    corners = {tuple(np.asarray(v)) for t in tris for v in t}
    c, a, b = (np.asarray(x) for x in (r.corner, r.edge_a, r.edge_b))
    assert corners == {tuple(c), tuple(c + a), tuple(c + a + b), tuple(c + b)}


def test_box_is_twelve_outward_triangles_covering_its_surface():
    b = BoxData((-1.0, 0.5, 2.0), (0.5, 2.5, 2.75), GREY)
    centre = (np.asarray(b.min) + np.asarray(b.max)) / 2
    tris = [t for r in b.rectangles() for t in r.triangles()]
    assert len(tris) == 12
    area = 0.0
    for t in tris:
        n = _normal(t)
        centroid = np.mean([np.asarray(v) for v in t], axis=0)
        assert np.dot(n, centroid - centre) > 0                                         # outward
        assert np.count_nonzero(n) == 1                                                 # axis-aligned face
        area += np.linalg.norm(n) / 2
    dx, dy, dz = np.asarray(b.max) - np.asarray(b.min)
    assert np.isclose(area, 2 * (dx * dy + dy * dz + dx * dz))
    flat = SceneData("b", OutputSettings(4, 4, 0.5), (0, 0, 0), [SphereData((0, 0, 0), 1.0, GREY, False), b],
                     CameraSettings((0, 0, -5), (0, 0, 0), (0, 1, 0)), CameraData(1.0, 500.0, 10.0, 0.0)).flatten()
    s = flat.struct
    assert (s.n_spheres, s.n_triangles) == (1, 12)
    assert list(np.ctypeslib.as_array(s.tri_shape_id, shape=(12,))) == list(range(1, 13))   # every triangle its own id


def room_scene(width=40, height=30):
    """A closed room (inward-facing rectangles), a rectangle light under the ceiling, two boxes (one mirror-like,
    one glossy) and a sphere.  Everything a path can hit after the camera ray is a triangle or the sphere."""
    white, red, green = Matte((0.75, 0.75, 0.75), (1, 1, 1), 1.0), Matte((0.75, 0.2, 0.2), (1, 1, 1), 1.0), Matte((0.2, 0.7, 0.25), (1, 1, 1), 1.0)
    shapes = [
        RectangleData((-3, 0, -3), (0, 0, 6), (6, 0, 0), white),        # floor, normal +y
        RectangleData((-3, 5, -3), (6, 0, 0), (0, 0, 6), white),        # ceiling, normal -y
        RectangleData((-3, 0, 3), (0, 5, 0), (6, 0, 0), white),         # back wall, normal -z
        RectangleData((-3, 0, -3), (0, 5, 0), (0, 0, 6), red),          # left wall, normal +x
        RectangleData((3, 0, -3), (0, 0, 6), (0, 5, 0), green),         # right wall, normal -x
        RectangleData((-1, 4.99, -1), (2, 0, 0), (0, 0, 2), Emissive((1.0, 0.95, 0.85), 9.0)),   # light, facing down
        BoxData((-2.0, 0.0, 0.2), (-0.6, 2.6, 1.6), Reflective(0.85, (0.95, 0.95, 0.95))),
        BoxData((0.6, 0.0, -0.9), (1.9, 1.2, 0.4), GlossyReflective(0.6, (0.9, 0.8, 0.6), 50.0)),
        SphereData((1.25, 1.8, -0.25), 0.6, white, False),
    ]
    return SceneData("room", OutputSettings(width, height, 0.5), (0.0, 0.0, 0.0), shapes,
                     CameraSettings((0.0, 2.5, -8.5), (0.0, 2.3, 0.0), (0.0, 1.0, 0.0)),
                     CameraData(3.2 * width / 800.0, 500.0, 8.0, 0.02))


def test_room_scene_oracle_is_lit_only_through_the_rectangle():
    """Host-side sanity of the scene itself (CPU oracle): it is closed (no path escapes to the black background
    undetected: miss count 0 beyond the open front) and the only emitter is the rectangle."""
    sd = room_scene(16, 12)
    cfg = JobConfiguration(3, 5, 50)
    ss = Hp.oracle_samples(4, cfg, 16, 12)
    img, cn = O.render_rows(sd.flatten(), cfg, ss, 0, 11, counters=True)
    assert cn["hit_tri"] > 0 and cn["emissive"] > 0 and cn["hit_plane"] == 0
    assert np.isfinite(img).all() and img.max() > 0.05


@pytest.mark.gpu
@pytest.mark.parametrize("accel", [1, 2])
def test_room_scene_matches_the_oracle_through_every_kernel(gpu_ctx, accel):
    sd = room_scene()
    cfg = JobConfiguration(16, 5, 50)      # 256 spp: direct, regeneration and (with the BVH) the wavefront kernel
    flat = sd.flatten()
    ss = Hp.oracle_samples(6, cfg, 40, 30)
    ref, cn_o = O.render_rows(flat, cfg, ss, 0, 29, counters=True)
    try:
        gpu_ctx.set_accel_mode(accel)
        Hp.upload(gpu_ctx, flat, cfg, ss)
        modes = (1, 2, 4) if accel == 2 else (1,)      # the linear-scan wavefront / regeneration kernels take no triangles
        for mode in modes:
            gpu_ctx.set_kernel_mode(mode)
            gpu_ctx.enable_counters(True)
            gpu_ctx.reset_counters()
            img = gpu_ctx.render_rows(0, 29, 40)
            cn = gpu_ctx.counters()
            gpu_ctx.enable_counters(False)
            assert Hp.rel_err(img, ref) <= 1e-6, (accel, mode)
            for k in ("samples", "segments", "hit_tri", "hit_sphere", "emissive", "matte", "specular", "glossy", "miss", "depth_cut"):
                assert abs(cn[k] - cn_o[k]) <= max(2, 1e-6 * cn_o[k]), (accel, mode, k, cn[k], cn_o[k])
    finally:
        gpu_ctx.set_kernel_mode(0)
        gpu_ctx.set_accel_mode(0)
        gpu_ctx.enable_counters(False)


@pytest.mark.gpu
def test_rectangle_diagonal_and_box_edges_are_watertight(gpu_ctx):
    """Rays aimed exactly at a rectangle's diagonal, at box edges and corners: the shared edges of the triangles are
    inclusive on both sides (0 <= u, v >= 0, u + v <= 1), so no ray slips through, and the GPU names the triangle the
    oracle names (lowest shape id on exact ties)."""
    sd = room_scene()
    flat = sd.flatten()
    gpu_ctx.set_scene(flat, JobConfiguration(1))
    rng = np.random.default_rng(2)
    n = 4000
    o = np.tile(np.array([0.3, 2.0, -6.0]), (n, 1))
    targets = np.empty((n, 3))
    s = rng.random(n)
    targets[:1000] = np.array([-3, 0, -3]) + s[:1000, None] * np.array([6, 0, 6])            # floor diagonal
    targets[1000:2000] = np.array([-2.0, 0.0, 0.2]) + s[1000:2000, None] * np.array([0, 2.6, 0])   # a vertical box edge
    targets[2000:3000] = np.array([0.6, 1.2, -0.9]) + s[2000:3000, None] * np.array([1.3, 0, 0])   # a top box edge
    corners = np.array([[x, y, z] for x in (-2.0, -0.6) for y in (0.0, 2.6) for z in (0.2, 1.6)])
    targets[3000:] = corners[rng.integers(0, 8, 1000)]
    d = targets - o
    hit, t = gpu_ctx.trace_rays(o, d)
    ho, to = O.trace_rays(flat, o, d)
    assert np.array_equal(hit, ho) and np.array_equal(t.view(np.uint64), to.view(np.uint64))
    assert (hit >= 0).all()                                                                     # nothing slips through
