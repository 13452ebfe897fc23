"""Known-answer vectors for the CPU oracle (the reference ships no tests, so these
hand-derivable cases are what pins the restatement, next to the demo.png statistical check)."""
import math

import numpy as np
import pytest

from oracle import oracle_py as O

T_MIN = 0.0005  # fluxcore/src/constants.rs:4


# ---- Sphere::hit, shapes.rs:171-217 ------------------------------------------------------------
def test_sphere_head_on():
    h = O.sphere_hit((0, 0, 0), 1.0, False, (0, 0, -5), (0, 0, 1))
    assert h["t"] == 4.0
    assert np.array_equal(h["normal"], [0, 0, -1])
    assert np.array_equal(h["point"], [0, 0, -1])


def test_sphere_invert_flips_normal():
    h = O.sphere_hit((0, 0, 0), 1.0, True, (0, 0, -5), (0, 0, 1))
    assert h["t"] == 4.0
    assert np.array_equal(h["normal"], [0, 0, 1])


def test_sphere_origin_inside_takes_second_root():
    h = O.sphere_hit((0, 0, 0), 1.0, False, (0, 0, 0), (0, 0, 1))
    assert h["t"] == 1.0  # t1 = -1 <= T_MIN -> t2 = 1 (shapes.rs:199-212)
    assert np.array_equal(h["normal"], [0, 0, 1])


def test_sphere_behind_and_miss():
    assert O.sphere_hit((0, 0, 0), 1.0, False, (0, 0, 5), (0, 0, 1)) is None
    assert O.sphere_hit((0, 0, 0), 1.0, False, (0, 2, -5), (0, 0, 1)) is None


def test_sphere_unnormalised_direction_scales_t():
    h = O.sphere_hit((0, 0, 0), 1.0, False, (0, 0, -5), (0, 0, 2))
    assert h["t"] == 2.0  # a = 4, b = -40, c = 24, disc = 1216... t = (40 - sqrt(1600-384))/8
    assert np.array_equal(h["point"], [0, 0, -1])


def test_sphere_radius_scaling_of_normal():
    h = O.sphere_hit((1, 2, 3), 2.0, False, (1, 2, -7), (0, 0, 1))
    assert h["t"] == 8.0
    assert np.array_equal(h["normal"], [0, 0, -1])  # (temp + t d) / r


def test_sphere_tangent_axis_parallel_ray_rejected_by_bbox():
    # grazing ray along z at x = 1: slab test has tx_min = tx_max -> t0 < t1 can still hold on other axes;
    # disc = 0 exactly: the quadratic would hit at t = 5, and the bbox admits it (t0=4 < t1=6)
    h = O.sphere_hit((0, 0, 0), 1.0, False, (1, 0, -5), (0, 0, 1))
    assert h is not None and h["t"] == 5.0


# ---- BoundingBox::hit, shapes.rs:98-133 incl. IEEE corners ---------------------------------------
def test_bbox_basic():
    assert O.bbox_hit((-1, -1, -1), (1, 1, 1), (0, 0, -5), (0, 0, 1))
    assert not O.bbox_hit((-1, -1, -1), (1, 1, 1), (0, 0, 5), (0, 0, 1))   # behind: t1 < T_MIN
    assert not O.bbox_hit((-1, -1, -1), (1, 1, 1), (3, 0, -5), (0, 0, 1))  # outside the x slab


def test_bbox_zero_direction_component_on_slab_face_is_nan():
    # origin exactly on the x = 1 face, dx = 0: (c1x - ox) * inf = 0 * inf = NaN. The private
    # max/min (shapes.rs:90-96) return the second argument when the comparison is false, so the
    # NaN in tx_max poisons t1 = min(NaN, ..) -> comparison false -> returns the inner min; the
    # outcome is whatever the reference's expression order gives; we pin it.
    r = O.bbox_hit((-1, -1, -1), (1, 1, 1), (1, 0, -5), (0, 0, 1))
    # tx_min = (-1-1)*inf = -inf, tx_max = (1-1)*inf = NaN; t0 = max(-inf, max(ty_min, tz_min)) = 4
    # t1 = min(NaN, min(ty_max, tz_max)): NaN < 6 false -> 6.  4 < 6 and 6 > T_MIN -> hit
    assert r is True


def test_bbox_negative_zero_direction():
    # dx = -0.0: a = 1/-0.0 = -inf; `a >= 0.0` false -> swapped corners (shapes.rs:108-112)
    # tx_min = (c1x-ox)*(-inf), tx_max = (c0x-ox)*(-inf); inside the slab: -inf, +inf -> hit
    assert O.bbox_hit((-1, -1, -1), (1, 1, 1), (0, 0, -5), (-0.0, 0, 1))
    # outside the slab: ox = 3: tx_min = (1-3)*-inf = +inf -> t0 = inf -> miss
    assert not O.bbox_hit((-1, -1, -1), (1, 1, 1), (3, 0, -5), (-0.0, 0, 1))


# ---- Plane::hit, shapes.rs:135-152 -----------------------------------------------------------------
def test_plane_from_above():
    h = O.plane_hit((0, 0, 0), (0, 1, 0), (0, 1, 0), (0, -1, 0))
    assert h["t"] == 1.0
    assert np.array_equal(h["normal"], [0, 1, 0])  # as given: not normalised, not flipped
    assert np.array_equal(h["point"], [0, 0, 0])


def test_plane_normal_not_normalised_or_flipped():
    h = O.plane_hit((0, 0, 0), (0, -2, 0), (0, 1, 0), (0, -1, 0))
    assert h["t"] == 1.0 and np.array_equal(h["normal"], [0, -2, 0])


def test_plane_parallel_ray():
    # d.n = 0: t = (p-o).n / 0 = -inf (origin above) -> miss; +inf (origin below) -> a valid hit at t = inf
    assert O.plane_hit((0, 0, 0), (0, 1, 0), (0, 1, 0), (1, 0, 0)) is None
    h = O.plane_hit((0, 0, 0), (0, 1, 0), (0, -1, 0), (1, 0, 0))
    assert h is not None and math.isinf(h["t"]) and h["t"] > 0
    # origin in the plane: 0/0 = NaN -> `t > T_MIN` false -> miss
    assert O.plane_hit((0, 0, 0), (0, 1, 0), (0, 0, 0), (1, 0, 0)) is None


def test_plane_t_min_threshold():
    assert O.plane_hit((0, 0, 0), (0, 1, 0), (0, T_MIN, 0), (0, -1, 0)) is None  # t == T_MIN is not > T_MIN
    assert O.plane_hit((0, 0, 0), (0, 1, 0), (0, 2 * T_MIN, 0), (0, -1, 0)) is not None


# ---- Scene::hit tie rule, scene.rs:156-160 + common.rs:17-23 ------------------------------------
def _scene(shapes, w=8, h=8):
    from flux_b200 import CameraData, CameraSettings, OutputSettings, SceneData
    return SceneData("t", OutputSettings(w, h, 0.5), (0.1, 0.2, 0.3), shapes,
                     CameraSettings((0, 0, -9.0), (0, 0, 0), (0, 1, 0)), CameraData(1.0, 500.0, 10.0, 0.0)).flatten()


def test_tie_goes_to_earlier_shape():
    from flux_b200 import Emissive, Matte, PlaneData, SphereData
    m0, m1 = Matte((1, 0, 0), (1, 1, 1), 1.0), Emissive((0, 1, 0), 1.0)
    flat = _scene([SphereData((0, 0, 0), 1.0, m0, False), SphereData((0, 0, 0), 1.0, m1, False)])
    h = O.hit_record(flat, (0, 0, -5), (0, 0, 1))
    assert h["shape_id"] == 0 and h["t"] == 4.0 and h["material"] == 0
    # sphere and plane at the same distance, plane first in YAML order
    flat = _scene([PlaneData((0, 0, -1), (0, 0, -1), m1), SphereData((0, 0, 0), 1.0, m0, False)])
    h = O.hit_record(flat, (0, 0, -5), (0, 0, 1))
    assert h["shape_id"] == 0 and h["t"] == 4.0
    flat = _scene([SphereData((0, 0, 0), 1.0, m0, False), PlaneData((0, 0, -1), (0, 0, -1), m1)])
    h = O.hit_record(flat, (0, 0, -5), (0, 0, 1))
    assert h["shape_id"] == 0 and np.array_equal(h["normal"], [0, 0, -1])


def test_closest_wins_regardless_of_order():
    from flux_b200 import Matte, SphereData
    m = Matte((1, 1, 1), (1, 1, 1), 1.0)
    flat = _scene([SphereData((0, 0, 5), 1.0, m, False), SphereData((0, 0, 0), 1.0, m, False)])
    assert O.hit_record(flat, (0, 0, -5), (0, 0, 1))["shape_id"] == 1
    assert O.hit_record(flat, (0, 0, -5), (0, 1, 0)) is None


# ---- triangle (EXTENSION; semantics defined in DESIGN.md) ---------------------------------------------
def test_triangle_known_answers():
    v0, v1, v2 = (0, 0, 0), (1, 0, 0), (0, 1, 0)
    h = O.triangle_hit(v0, v1, v2, (0.25, 0.25, -2), (0, 0, 1))
    assert h["t"] == 2.0 and np.array_equal(h["normal"], [0, 0, 1]) and np.array_equal(h["point"], [0.25, 0.25, 0])
    h = O.triangle_hit(v0, v1, v2, (0.25, 0.25, 2), (0, 0, -1))  # two-sided, normal as wound
    assert h["t"] == 2.0 and np.array_equal(h["normal"], [0, 0, 1])
    assert O.triangle_hit(v0, v1, v2, (0.75, 0.75, -2), (0, 0, 1)) is None  # u+v > 1
    assert O.triangle_hit(v0, v1, v2, (-0.1, 0.2, -2), (0, 0, 1)) is None
    assert O.triangle_hit(v0, v1, v2, (0.25, 0.25, -2), (1, 0, 0)) is None  # parallel: det == 0
    assert O.triangle_hit(v0, v1, v2, (0.0, 0.0, -2), (0, 0, 1)) is not None  # vertex: u = v = 0 inclusive


# ---- CameraBasis::new, scene.rs:29-34; primary ray, trace.rs:44-51,72-80 -------------------------------
def test_camera_basis_axis_aligned():
    u, v, w = O.camera_basis((0, 0, -9), (0, 0, 0), (0, 1, 0))
    assert np.array_equal(w, [0, 0, -1])
    assert np.array_equal(u, [-1, 0, 0])  # up x w
    assert np.array_equal(v, [0, 1, 0])   # w x u


def test_primary_ray_pinhole_centre():
    from flux_b200 import Matte, SphereData
    flat = _scene([SphereData((0, 0, 0), 1.0, Matte((1, 1, 1), (1, 1, 1), 1.0), False)], w=8, h=8)
    # pixel sample (0,0) of pixel (row 4, col 4): u = 0.5*((4-4)+0) = 0, v = 0.5*((8-4)-4+0) = 0
    o, d = O.primary_ray(flat, 4, 4, 0.0, 0.0, 0.0, 0.0)
    assert np.array_equal(o, [0, 0, -9])
    assert np.array_equal(d, [0, 0, 1])  # -focal*w normalised
    # one pixel to the right on the image is -x in world space here (u = up x w = -x)
    o, d = O.primary_ray(flat, 4, 5, 0.0, 0.0, 0.0, 0.0)
    f = 10.0 / 500.0
    e = np.array([-(0.5 * f), 0.0, 10.0])
    assert np.allclose(d, e / np.linalg.norm(e), rtol=0, atol=1e-16)
    # lens sample moves the origin and bends the direction toward the same focal point
    flat.struct.lens_radius = 0.5
    o, d = O.primary_ray(flat, 4, 4, 0.0, 0.0, 1.0, 0.0)
    assert np.array_equal(o, [-0.5, 0, -9])
    focal_pt = o + d * (10.0 / d[2])
    assert np.allclose(focal_pt, [0, 0, 1], atol=1e-14)


# ---- BRDFs, brdf.rs:20-78 -----------------------------------------------------------------------
def test_lambertian_known_answer():
    n = np.array([0.0, 1.0, 0.0])
    wi, pdf, f = O.lambertian_sample_f(n, (0, 0, 1), (0.5, 0.25, 1.0), 0.8)
    assert np.array_equal(wi, n)  # hemisphere pole maps onto the normal
    assert pdf == 1.0 * (1.0 / math.pi)
    assert np.array_equal(f, [0.5 * 0.8 * (1 / math.pi), 0.25 * 0.8 * (1 / math.pi), 1.0 * 0.8 * (1 / math.pi)])
    # general sample: wi unit length, in the upper hemisphere, pdf = n.wi / pi
    wi, pdf, f = O.lambertian_sample_f(n, (0.6, 0.0, 0.8), (1, 1, 1), 1.0)
    assert abs(np.linalg.norm(wi) - 1) < 1e-15 and wi[1] > 0
    assert pdf == ((n[0] * wi[0] + n[1] * wi[1]) + n[2] * wi[2]) * (1 / math.pi)
    # basis: v = normalize((0.0034,1,0.0071) x n), u = v x n  -> for n = +y: v = (-0.0071,0,0.0034)/|.|
    vv = np.array([-0.0071, 0.0, 0.0034]); vv /= math.sqrt(0.0071 ** 2 + 0.0034 ** 2)
    uu = np.cross(vv, n)
    assert np.allclose(wi, (0.6 * uu + 0.8 * n) / np.linalg.norm(0.6 * uu + 0.8 * n), atol=1e-15)


def test_perfect_specular_known_answer():
    n = np.array([0.0, 1.0, 0.0])
    d = np.array([1.0, -1.0, 0.0]) / math.sqrt(2)
    wi, pdf, f = O.specular_sample_f(n, -d, (0.9, 0.8, 0.7), 0.5)
    assert np.allclose(wi, np.array([1.0, 1.0, 0.0]) / math.sqrt(2), atol=1e-16)
    assert pdf == ((n[0] * wi[0] + n[1] * wi[1]) + n[2] * wi[2])
    assert np.array_equal(f, [0.45, 0.4, 0.35])


def test_glossy_known_answer_pole_and_flip():
    n = np.array([0.0, 1.0, 0.0])
    d = np.array([1.0, -1.0, 0.0]) / math.sqrt(2)
    r = np.array([1.0, 1.0, 0.0]) / math.sqrt(2)
    # sample y = 0 -> cos_theta = 1 -> hemisphere pole -> wi = r, lobe = (r.r)^e ~ 1
    wi, pdf, f, flipped = O.glossy_sample_f(n, -d, (0.3, 0.0), (1, 1, 1), 0.5, 100.0)
    assert not flipped and np.allclose(wi, r, atol=1e-15)
    assert abs(pdf - wi[1]) < 1e-13 and np.allclose(f, 0.5, atol=1e-13)
    # grazing reflection + wide lobe: some samples fall below the surface and are mirrored (brdf.rs:67-71)
    dg = np.array([1.0, -0.05, 0.0]); dg /= np.linalg.norm(dg)
    nflip = 0
    for k in range(64):
        wi, pdf, f, flipped = O.glossy_sample_f(n, -dg, ((k + 0.5) / 64, 0.9), (1, 1, 1), 1.0, 1.0)
        nflip += flipped
        assert wi[1] >= -1e-15  # always above the surface after the flip
    assert 0 < nflip < 64


# ---- samplers/src/lib.rs:133-182 -------------------------------------------------------------
def test_to_unit_hemi_known_answers():
    assert np.allclose(O.to_unit_hemi(0.0, 0.0, 0.0), [0, 0, 1], atol=1e-16)   # y=0: pole
    h = O.to_unit_hemi(0.0, 1.0, 0.0)                                             # y=1: horizon, phi=0
    assert np.allclose(h, [1, 0, 0], atol=1e-16)
    h = O.to_unit_hemi(0.25, 0.5, 0.0)                                            # e=0: cos_theta = 1-y (uniform)
    assert abs(h[2] - 0.5) < 2e-16 and abs(h[1] - math.sqrt(0.75)) < 1e-15 and abs(h[0]) < 1e-16  # normalize() rounds
    h = O.to_unit_hemi(0.5, 0.5, 3.0)                                             # cos_theta = 0.5^(1/4)
    assert abs(h[2] - 0.5 ** 0.25) < 1e-15 and h[0] < 0


def test_to_poisson_disc_quadrants():
    # centre and the four branch regions of the Shirley concentric map (lib.rs:152-172)
    assert np.allclose(O.to_poisson_disc(0.5, 0.5), [0, 0], atol=0)      # spy == 0 guard: phi = 0, r = 0
    assert np.allclose(O.to_poisson_disc(1.0, 0.5), [1, 0], atol=1e-16)   # region 1: r = spx
    assert np.allclose(O.to_poisson_disc(0.5, 1.0), [0, 1], atol=1e-15)   # region 2: r = spy, phi = pi/2
    assert np.allclose(O.to_poisson_disc(0.0, 0.5), [-1, 0], atol=1e-15)  # region 3: r = -spx, phi = pi
    assert np.allclose(O.to_poisson_disc(0.5, 0.0), [0, -1], atol=1e-15)  # region 4: r = -spy, phi = 3pi/2
    rng = np.random.default_rng(0)
    for x, y in rng.uniform(0, 1, (200, 2)):
        d = O.to_poisson_disc(x, y)
        assert d[0] ** 2 + d[1] ** 2 <= 1.0 + 1e-15
        # radius is the sup-norm of the centred square point
        assert abs(math.hypot(*d) - max(abs(2 * x - 1), abs(2 * y - 1))) < 1e-15


# ---- multi-jittered grids, lib.rs:46-126 --------------------------------------------------------
@pytest.mark.parametrize("root", [1, 2, 3, 8, 16])
@pytest.mark.parametrize("correlated", [False, True])
def test_multi_jittered_stratification(root, correlated):
    g = O.mj_grid(99, root, 3, 2, correlated)
    n = root * root
    assert g.shape == (n, 2) and (g >= 0).all() and (g < 1).all()
    # n-rooks: exactly one sample in each of the n fine columns and n fine rows
    assert sorted(np.floor(g[:, 0] * n).astype(int)) == list(range(n))
    assert sorted(np.floor(g[:, 1] * n).astype(int)) == list(range(n))
    cells = np.floor(g[:, 0] * root).astype(int) * root + np.floor(g[:, 1] * root).astype(int)
    if correlated:
        # one shared x- and one shared y-permutation (lib.rs:78-82): sample (i, j) lands in coarse cell
        # (pix(i), piy(j)), a bijection -> exactly one sample per coarse cell
        assert sorted(cells) == list(range(n))
    # (uncorrelated: sample (i, j) lands in coarse cell (pix_j(i), piy_i(j)) (lib.rs:92-126), which need
    #  not cover every cell — a property of the reference's shuffle that is reproduced, not fixed)
    # fine offsets: x keeps sub-column root-1-j, y keeps sub-row root-1-i (lib.rs:56-59)
    k = np.arange(n)
    assert np.array_equal(np.floor(g[:, 0] * n).astype(int) % root, root - 1 - k % root)
    assert np.array_equal(np.floor(g[:, 1] * n).astype(int) % root, root - 1 - k // root)


def test_sample_sets_layout_and_determinism():
    a = O.generate_samples(5, 4, 3, 7)
    b = O.generate_samples(5, 4, 3, 7)
    c = O.generate_samples(6, 4, 3, 7)
    assert a.pixel.shape == (7, 16, 2) and a.disc.shape == (7, 16, 2) and a.hemi.shape == (7, 3, 16, 3)
    assert np.array_equal(a.pixel, b.pixel) and np.array_equal(a.hemi, b.hemi) and not np.array_equal(a.pixel, c.pixel)
    assert np.allclose(np.linalg.norm(a.hemi, axis=-1), 1.0, atol=1e-15) and (a.hemi[..., 2] >= 0).all()
    assert (np.linalg.norm(a.disc, axis=-1) <= 1 + 1e-15).all()
    assert not np.array_equal(a.pixel[0], a.pixel[1])  # sets differ
    assert not np.array_equal(a.hemi[0, 0], a.hemi[0, 1])  # depths differ


def test_set_index_rows_are_permutations():
    idx = O.generate_set_index(3, 20, 50, 50)
    for r in range(20):
        assert sorted(idx[r]) == list(range(50))  # shuffle_indices, sampling.rs:35-40
    assert not np.array_equal(idx[0], idx[1])
    idx = O.generate_set_index(3, 4, 50, 7)  # fewer sets than columns: perm[col % num_sets]
    assert idx.max() == 6 and np.array_equal(idx[:, :7], idx[:, 7:14])


# ---- Color::max_to_one, color.rs:35-44; Image::write quantisation, image.rs:50-52 ------------------
def test_max_to_one():
    assert np.array_equal(O.max_to_one((0.2, 0.5, 1.0)), [0.2, 0.5, 1.0])
    assert np.array_equal(O.max_to_one((2.0, 1.0, 0.5)), [2.0 * 0.5, 1.0 * 0.5, 0.5 * 0.5])
    out = O.max_to_one((float("nan"), 4.0, 1.0))  # NaN compares false: mx1 = g = 4 -> scaled, NaN stays
    assert math.isnan(out[0]) and out[1] == 1.0 and out[2] == 0.25
    out = O.max_to_one((0.5, 0.25, float("nan")))  # mx1 > NaN false -> mx2 = NaN -> not > 1 -> untouched
    assert out[0] == 0.5 and math.isnan(out[2])


def test_ppm_quantisation():
    q = O.ppm_quantize([0.0, 1.0, 0.5, float("nan"), -0.1, 2.0, 1e-6])
    assert list(q) == [0, 65535, 32767, 0, 0, 65535, 0]
