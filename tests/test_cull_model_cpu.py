"""CPU models of the two places where the render kernels decide a bounding-box test WITHOUT running the reference's
f64 slab arithmetic (BoundingBox::hit, shapes.rs:98-133) — and property tests that such a decision never differs
from it.  On the GPU the same claim is checked end to end (bbox_pass counters equal to the oracle's,
tests/test_gpu_parity.py); here the arithmetic of the device functions is restated in numpy and driven with
adversarial inputs the renders rarely produce: rays aimed at box faces, edges and corners to within 1e-12 .. 1e-4,
far planes sitting on T_MIN, huge and tiny coordinates.

  * cull_boxes / lin_cull_ray (render_wave2.cu, flux_cull.cuh): FP32 classification "certainly hit / certainly
    missed / undecided" with the error bound E = 1.01 * 2^-20 * max_k |1/d_k| (cmax + |o_k|).  The model rounds every
    step to float32 like the device; the device's rcp.approx.f32 (1 ulp) and single-rounded fmaf are modelled by a
    correctly rounded reciprocal and a double-rounded fma, and that difference is covered by ALSO demanding the
    property at a QUARTER of the device's bound.
  * primary_may_hit (render_wave2.cu): a sphere left out of a pixel's primary mask must fail the box test for every
    camera ray of that pixel (any pixel sample in [0,1]^2, any lens sample in the unit disc).
"""
import numpy as np
import pytest

T_MIN = 0.0005
F32 = np.float32


def exact_bbox_hit(c, r, o, d):
    """BoundingBox::hit (shapes.rs:98-133) on corners center -+ r (Sphere::new, shapes.rs:154-161), vectorised."""
    with np.errstate(all="ignore"):
        c0, c1 = c - r[:, None], c + r[:, None]
        a = 1.0 / d
        pos = a >= 0.0
        tmin = np.where(pos, (c0 - o) * a, (c1 - o) * a)
        tmax = np.where(pos, (c1 - o) * a, (c0 - o) * a)
        mx = lambda p, q: np.where(p > q, p, q)      # noqa: E731  shapes.rs:94-96
        mn = lambda p, q: np.where(p < q, p, q)      # noqa: E731  shapes.rs:90-92
        t0 = mx(tmin[:, 0], mx(tmin[:, 1], tmin[:, 2]))
        t1 = mn(tmax[:, 0], mn(tmax[:, 1], tmax[:, 2]))
        return (t0 < t1) & (t1 > T_MIN)


def fma32(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(F32)


def classify_f32(c, r, o, d, cmax, bound_scale=1.0):
    """cull_boxes + lin_cull_ray: returns (ok, fail) masks.  c, r: sphere (f64, rounded to f32 as flux_set_scene does);
    cmax: the scene bound RenderParams::cull_cmax (f32)."""
    with np.errstate(all="ignore"):
        cf, rf = c.astype(F32), r.astype(F32)
        of, df = o.astype(F32), d.astype(F32)
        ia = (F32(1.0) / df).astype(F32)
        no = (-(of * ia)).astype(F32)
        aa = np.abs(ia)
        e = (aa * (cmax[:, None] + np.abs(of))).astype(F32)
        E = (np.maximum(np.maximum(e[:, 0], e[:, 1]), e[:, 2]) * F32(1.01 * 9.5367431640625e-07)).astype(F32)
        sane = (E > F32(1e-30)) & (E < F32(1e30))
        e2 = np.where(sane, (F32(2.0) * E * F32(bound_scale) + F32(4e-9) * F32(bound_scale)).astype(F32), F32(np.nan))
        tc = fma32(cf, ia, no)
        near = fma32(-rf[:, None] * np.ones_like(aa), aa, tc)
        far = fma32(rf[:, None] * np.ones_like(aa), aa, tc)
        tn = np.fmax(np.fmax(near[:, 0], near[:, 1]), np.fmax(near[:, 2], F32(T_MIN)))
        tf = np.fmin(np.fmin(far[:, 0], far[:, 1]), far[:, 2])
        sgap = (tf - tn).astype(F32)
        return sgap > e2, sgap < -e2


def host_cull_entry(c, r):
    """flux_set_scene (api.cu): which spheres get an f32 entry at all, and the per-scene bound contribution."""
    ok = np.isfinite(c).all(axis=1) & np.isfinite(r) & (r >= 0.0) & (np.abs(c) < 1e30).all(axis=1) & (r < 1e30)
    up = F32(np.inf)
    with np.errstate(all="ignore"):
        cm = np.nextafter((np.abs(c) + r[:, None]).astype(F32), up).max(axis=1)
    return ok, np.nextafter(cm, up)


def adversarial_pairs(rng, n, scale):
    """(sphere, ray) pairs whose ray passes within a relative 1e-12 .. 1e-4 of a face, an edge or a corner of the box,
    or whose box's far plane sits within that of T_MIN."""
    c = rng.uniform(-scale, scale, (n, 3))
    r = np.abs(rng.normal(0, 0.3 * scale, n)) + 1e-3 * scale
    kind = rng.integers(0, 4, n)
    s = rng.choice([-1.0, 1.0], (n, 3))
    free = rng.uniform(-1, 1, (n, 3))
    nfree = np.where(kind[:, None] == 0, 2, np.where(kind[:, None] == 1, 1, 0))       # face, edge, corner (kind 3: T_MIN)
    pick = np.argsort(rng.random((n, 3)), axis=1)
    m = pick < nfree
    q = c + r[:, None] * np.where(m, free, s)                       # a point of the box surface
    o = q + rng.normal(0, 1.0, (n, 3)) * rng.choice([0.1, 1.0, 10.0], (n, 1)) * scale
    eps = 10.0 ** rng.uniform(-12, -4, (n, 1)) * rng.choice([-1.0, 1.0], (n, 1))
    target = q + eps * r[:, None] * s                                # just outside / just inside
    d = target - o
    d /= np.linalg.norm(d, axis=1, keepdims=True) * rng.choice([1.0, 1.0, 0.37, 12.0], (n, 1))   # not always unit (bounces are)
    # kind 3: origin such that the box is left at t = T_MIN (1 +- 1e-9 ..), along d
    k3 = kind == 3
    t_exit = T_MIN * (1.0 + 10.0 ** rng.uniform(-12, -5, n) * rng.choice([-1.0, 1.0], n))
    o[k3] = q[k3] - d[k3] * t_exit[k3, None]
    return c, r, o, d


@pytest.mark.parametrize("scale", [1.0, 100.0, 1e-2, 1e4])
def test_fp32_classification_never_contradicts_the_f64_slab_test(scale):
    rng = np.random.default_rng(int(scale * 1000) % 9973 + 1)
    n = 400_000
    c, r, o, d = adversarial_pairs(rng, n, scale)
    # plus plain random rays, axis-parallel rays (1/d = inf: nothing may be decided) and rays from inside
    m = n // 4
    o[:m] = rng.uniform(-3 * scale, 3 * scale, (m, 3))
    d[:m] = rng.normal(0, 1, (m, 3))
    d[: m // 4, 0] = 0.0
    d[m // 4: m // 2, 1] = -0.0
    o[m // 2: m // 2 + m // 8] = c[m // 2: m // 2 + m // 8] + 0.3 * r[m // 2: m // 2 + m // 8, None]
    ok_entry, cm = host_cull_entry(c, r)
    assert ok_entry.all()
    # the scene bound is a maximum over spheres: the tightest (hardest) case is the sphere's own, a looser one is random
    for cmax in (cm, np.maximum(cm, F32(3.0 * scale))):
        exact = exact_bbox_hit(c, r, o, d)
        for bound_scale in (1.0, 0.25):
            ok, fail = classify_f32(c, r, o, d, cmax.astype(F32), bound_scale)
            assert not np.any(ok & fail)
            wrong = (ok & ~exact) | (fail & exact)
            assert not wrong.any(), f"{wrong.sum()} contradictions at bound scale {bound_scale}: first {np.flatnonzero(wrong)[:3]}"
        ok, fail = classify_f32(c, r, o, d, cmax.astype(F32))
        undecided = ~(ok | fail)
        # ... and it is worth having: random rays are decided in FP32 almost always, the adversarial ones often not
        assert undecided[m // 2 + m // 8: m].mean() < 0.02 or scale != 1.0
        assert undecided[: m // 2].all()      # a zero direction component decides nothing


def primary_may_hit(cam, colf, rowf, c, r):
    """render_wave2.cu primary_may_hit, vectorised over spheres."""
    f = cam["focal"]
    if not f > 1e-9:
        return np.ones(len(r), bool)
    k = cam["aps"] * cam["factor"]
    q0u, q1u, q0v, q1v = k * colf, k * (colf + 1.0), k * rowf, k * (rowf + 1.0)
    qcu, qcv = 0.5 * (q0u + q1u), 0.5 * (q0v + q1v)
    rF = 0.5 * np.sqrt((q1u - q0u) ** 2 + (q1v - q0v) ** 2) * (1.0 + 1e-9)
    R = abs(cam["lens_radius"]) * (1.0 + 1e-9)
    rel = c - cam["eye"]
    cu, cv, cs = rel @ cam["u"], rel @ cam["v"], -(rel @ cam["w"])
    rho = 1.7320508075688774 * np.abs(r) * (1.0 + 1e-9)
    behind = cs + rho < 0.0
    s0 = np.where(cs > 0.0, cs, 0.0)
    a = s0 / f
    du, dv = cu - qcu * a, cv - qcv * a
    D = np.sqrt(du * du + dv * dv) - (R * np.abs(1.0 - a) + rF * a)
    L = (np.sqrt(qcu * qcu + qcv * qcv) + R + rF) / f
    scale = np.abs(cu) + np.abs(cv) + np.abs(cs) + rho + R + 1.0
    excluded = D - L * rho > rho * (1.0 + 1e-6) + 1e-9 * scale
    return ~(behind | excluded)


def _unit(v):
    return v / np.sqrt(v @ v)


@pytest.mark.parametrize("seed", range(6))
def test_primary_mask_never_drops_a_box_a_camera_ray_of_the_pixel_passes(seed):
    rng = np.random.default_rng(100 + seed)
    eye = rng.uniform(-5, 5, 3)
    w = _unit(eye - rng.uniform(-1, 1, 3))                      # CameraBasis::new, scene.rs:29-34
    u = _unit(np.cross(np.array([0.0, 1.0, 0.0]), w))
    v = np.cross(w, u)
    W, H = 64, 48
    focal = float(rng.choice([2.0, 12.0, 60.0]))
    vpd = float(rng.choice([50.0, 500.0]))
    cam = dict(eye=eye, u=u, v=v, w=w, focal=focal, factor=focal / vpd, aps=float(rng.choice([0.5, 1.0])) / float(rng.choice([1.0, 4.0])),
               lens_radius=float(rng.choice([0.0, 0.09, 0.6])))
    n_s = 4000
    total_excluded = 0
    for _ in range(12):
        col, row = int(rng.integers(0, W)), int(rng.integers(0, H))
        colf, rowf = float(col) - W * 0.5, float(H - row) - H * 0.5            # trace.rs:72-73
        # spheres scattered about the pixel's beam, many of them just outside it
        s = rng.uniform(0.05, 3.0, n_s) * focal
        k = cam["aps"] * cam["factor"]
        axis_u, axis_v = k * (colf + 0.5) * s / focal, k * (rowf + 0.5) * s / focal
        r = 10.0 ** rng.uniform(-2, 0.5, n_s)
        beam = cam["lens_radius"] * np.abs(1 - s / focal) + k * s / focal
        off = (beam + 1.7320508075688774 * r) * rng.uniform(0.6, 1.6, n_s)
        ang = rng.uniform(0, 2 * np.pi, n_s)
        c = eye + np.outer(axis_u + off * np.cos(ang), u) + np.outer(axis_v + off * np.sin(ang), v) - np.outer(s, w)
        may = primary_may_hit(cam, colf, rowf, c, r)
        ex = np.flatnonzero(~may)
        total_excluded += len(ex)
        if len(ex) == 0:
            continue
        # camera rays of the pixel as the reference makes them (trace.rs:44-51, 72-79): the corners and edges of the
        # sample square and of the lens disc are the extreme ones
        n_r = 300
        px, py = rng.random(n_r), rng.random(n_r)
        px[:40], py[:40] = rng.choice([0.0, 1.0], 40), rng.choice([0.0, 1.0], 40)
        la, lr = rng.uniform(0, 2 * np.pi, n_r), np.sqrt(rng.random(n_r))
        lr[:120] = 1.0
        lx, ly = lr * np.cos(la) * cam["lens_radius"], lr * np.sin(la) * cam["lens_radius"]
        uu, vv = cam["aps"] * (colf + px), cam["aps"] * (rowf + py)
        px2, py2 = uu * cam["factor"], vv * cam["factor"]
        dirs = np.outer(px2 - lx, u) + np.outer(py2 - ly, v) - focal * w
        dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
        orig = eye + np.outer(lx, u) + np.outer(ly, v)
        for j in ex[:: max(1, len(ex) // 400)]:
            hit = exact_bbox_hit(np.repeat(c[j:j + 1], n_r, 0), np.repeat(r[j:j + 1], n_r), orig, dirs)
            assert not hit.any(), f"sphere {j} left out of the mask of pixel ({col}, {row}) is hit by {hit.sum()} camera rays"
    assert total_excluded > 1000      # the mask does something on these scenes


# ---- the BVH's FP32 node test (flux_bvh.cuh: BvhTraversal::begin + BVH_CHILD) -------------------------------------------
def _dot(a, b):
    return (a[:, 0] * b[:, 0] + a[:, 1] * b[:, 1]) + a[:, 2] * b[:, 2]


def exact_sphere_t(c, r, o, d):
    """Sphere::hit distance (shapes.rs:171-217) or NaN, vectorised."""
    with np.errstate(all="ignore"):
        ok = exact_bbox_hit(c, r, o, d)
        temp = o - c
        a = _dot(d, d)
        b = 2.0 * _dot(temp, d)
        cc = _dot(temp, temp) - r * r
        disc = b * b - 4.0 * a * cc
        e = np.sqrt(np.where(disc < 0.0, np.nan, disc))
        t1, t2 = (-b - e) / (2.0 * a), (-b + e) / (2.0 * a)
        t = np.where(t1 > T_MIN, t1, np.where(t2 > T_MIN, t2, np.nan))
        return np.where(ok, t, np.nan)


def exact_tri_t(v0, e1, e2, o, d):
    """tri_t (flux_intersect.cuh / oracle tri_hit: the extension's definition) or NaN."""
    with np.errstate(all="ignore"):
        p = np.cross(d, e2)
        det = _dot(e1, p)
        s = o - v0
        inv = 1.0 / det
        u = _dot(s, p) * inv
        q = np.cross(s, e1)
        v = _dot(d, q) * inv
        t = _dot(e2, q) * inv
        ok = (det != 0.0) & (u >= 0.0) & (u <= 1.0) & (v >= 0.0) & (u + v <= 1.0) & (t > T_MIN)
        return np.where(ok, t, np.nan)


def round_out(lo, hi):
    """Builder::set_child: f64 box to f32, outward."""
    flo, fhi = lo.astype(F32), hi.astype(F32)
    flo = np.where(flo.astype(np.float64) > lo, np.nextafter(flo, F32(-np.inf)), flo)
    fhi = np.where(fhi.astype(np.float64) < hi, np.nextafter(fhi, F32(np.inf)), fhi)
    return flo, fhi


def node_test_f32(flo, fhi, o, d, ext, t_best):
    """True where the traversal would SKIP the child box [flo, fhi] for a ray whose best hit so far is t_best."""
    with np.errstate(all="ignore"):
        df, of = d.astype(F32), o.astype(F32)
        a = (F32(1.0) / df).astype(F32)
        a = np.where(np.abs(a) < F32(1e30), a, np.copysign(F32(1e30), df))
        noa = (-(of * a)).astype(F32)
        extf = np.nextafter(ext.astype(F32), F32(np.inf))[:, None]
        E = (np.abs(a) * (extf + np.abs(of)) * F32(1.01 * 9.5367431640625e-07)).astype(F32)
        sane = (df == df) & (np.abs(noa) < F32(1e37)) & (E < F32(1e37))
        pos = ~np.signbit(a)
        ia = np.where(sane, a, F32(0.0))
        nlo = np.where(sane, (noa - E).astype(F32), F32(-np.inf))
        nhi = np.where(sane, (noa + E).astype(F32), F32(np.inf))
        near, far = np.where(pos, flo, fhi), np.where(pos, fhi, flo)
        sn, sf = fma32(near, ia, nlo), fma32(far, ia, nhi)
        tn = np.fmax(np.fmax(sn[:, 0], sn[:, 1]), sn[:, 2])
        tf = np.fmin(np.fmin(sf[:, 0], sf[:, 1]), sf[:, 2])
        inv = 1.0 / d
        tscale = ext * np.minimum(np.abs(inv[:, 0]), np.minimum(np.abs(inv[:, 1]), np.abs(inv[:, 2])))
        t_prune = t_best + 1e-9 * (np.abs(t_best) + tscale)
        tp32 = t_prune.astype(F32)
        tp32 = np.where(tp32.astype(np.float64) < t_prune, np.nextafter(tp32, F32(np.inf)), tp32)   # __double2float_ru
        return (tn > tf) | (tf < F32(0.000499)) | (tn > tp32)


@pytest.mark.parametrize("scale", [1.0, 50.0, 1e3])
def test_bvh_node_test_never_skips_a_box_whose_primitive_yields_the_hit(scale):
    """A node may be skipped only if nothing below it can give a candidate at or before the best hit.  Hardest case: the
    child box is the primitive's own box (padded by 1e-7 of the scene extent and rounded outward, as the builder makes a
    leaf's box) and the best hit so far is this very primitive's distance (a tie the lower shape id must still win)."""
    rng = np.random.default_rng(int(scale) + 17)
    n = 300_000
    # spheres, rays through random points of the ball — many grazing — from near and far, some inside
    c = rng.uniform(-scale, scale, (n, 3))
    r = 10.0 ** rng.uniform(-3, 0, n) * scale * 0.1
    u = rng.normal(0, 1, (n, 3))
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    rad = np.where(rng.random(n) < 0.5, 1.0 - 10.0 ** rng.uniform(-12, -1, n), rng.random(n) ** (1 / 3))
    q = c + u * (r * rad)[:, None]
    o = q + rng.normal(0, 1, (n, 3)) * rng.choice([0.01, 1.0, 30.0], (n, 1)) * scale
    d = q - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    flat = rng.random(n) < 0.15                    # nearly axis-parallel: huge reciprocals on the other axes
    d[flat] *= np.array([1.0, 1e-9, 1e-7])
    d[flat] /= np.linalg.norm(d[flat], axis=1, keepdims=True)
    zero = rng.random(n) < 0.05
    d[zero, 2] = 0.0
    t = exact_sphere_t(c, r, o, d)
    ext = (np.abs(c) + r[:, None]).max(axis=1) * rng.choice([1.0, 1.0, 4.0], n)     # the scene holds at least this sphere
    pad = 1e-7 * ext
    flo, fhi = round_out(c - r[:, None] - pad[:, None], c + r[:, None] + pad[:, None])
    hit = ~np.isnan(t)
    assert hit.sum() > n // 3
    skip = node_test_f32(flo, fhi, o, d, ext, np.where(hit, t, np.inf))
    assert not (skip & hit).any(), f"{(skip & hit).sum()} sphere hits lost behind their own node box"
    # ancestors: any box that contains this one (rounding is monotone, so the slabs only widen)
    grow = (10.0 ** rng.uniform(-9, 1, (n, 3)) * r[:, None]).astype(F32)
    skip = node_test_f32(flo - grow * rng.integers(0, 2, (n, 3)).astype(F32), fhi + grow * rng.integers(0, 2, (n, 3)).astype(F32),
                         o, d, ext, np.where(hit, t, np.inf))
    assert not (skip & hit).any(), f"{(skip & hit).sum()} sphere hits lost behind an ancestor's box"
    # ... and the test is worth having: rays that miss the box by a clear margin are skipped
    far_off = c + u * (r * 4.0)[:, None] + 3.0 * r[:, None]
    d2 = far_off - o
    d2 /= np.linalg.norm(d2, axis=1, keepdims=True)
    miss = ~exact_bbox_hit(c, r * 1.5, o, d2)
    assert node_test_f32(flo, fhi, o, d2, ext, np.full(n, np.inf))[miss].mean() > 0.95

    # triangles, rays through random points of the face — many on its edges
    v0 = rng.uniform(-scale, scale, (n, 3))
    e1, e2 = rng.normal(0, 0.05 * scale, (n, 3)), rng.normal(0, 0.05 * scale, (n, 3))
    e1[: n // 10, 1] = 0.0
    e2[: n // 10, 1] = 0.0                                   # axis-aligned faces: a box of zero thickness
    b1, b2 = rng.random(n), rng.random(n)
    fold = b1 + b2 > 1.0
    b1, b2 = np.where(fold, 1.0 - b1, b1), np.where(fold, 1.0 - b2, b2)
    b2 = np.where(rng.random(n) < 0.3, 10.0 ** rng.uniform(-14, -6, n), b2)           # along an edge
    q = v0 + e1 * b1[:, None] + e2 * b2[:, None]
    o = q + rng.normal(0, 1, (n, 3)) * rng.choice([0.01, 1.0, 30.0], (n, 1)) * scale
    d = q - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    t = exact_tri_t(v0, e1, e2, o, d)
    v1, v2 = v0 + e1, v0 + e2
    lo, hi = np.minimum(np.minimum(v0, v1), v2), np.maximum(np.maximum(v0, v1), v2)
    ext = np.abs(np.concatenate([lo, hi], axis=1)).max(axis=1) * rng.choice([1.0, 1.0, 4.0], n)
    pad = 1e-7 * ext
    flo, fhi = round_out(lo - pad[:, None], hi + pad[:, None])
    hit = ~np.isnan(t)
    assert hit.sum() > n // 3
    skip = node_test_f32(flo, fhi, o, d, ext, np.where(hit, t, np.inf))
    assert not (skip & hit).any(), f"{(skip & hit).sum()} triangle hits lost behind their own node box"
