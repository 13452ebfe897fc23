"""second_opinion.py — a SECOND, independent CPU restatement of the reference's render path, in scalar Python.

TEST INFRASTRUCTURE ONLY (like everything under oracle/): imported by tests/ alone, never by flux_b200/, bench.py's
timed regions or the product path.

Why it exists.  The reference has no tests and cannot be built here (Rust), so `oracle/flux_oracle.cpp` — what every
GPU parity test compares against — is pinned only by hand-derived vectors and by `demo.png` statistically ("parity
unpinned").  What CAN be ruled out is a transcription slip in that oracle: this file was written a
second time straight from the Rust sources (not from the C++), as close to their text as Python allows — recursion
where the reference recurses, the operators in the order the reference writes them — and `tests/
test_second_opinion.py` demands that the two restatements agree BIT FOR BIT on whole renders.  Python floats are IEEE
binary64, every operation below is a single correctly rounded +, -, *, / or sqrt (no contraction is possible), and
`math.sin / cos / pow` are the C library's, the same ones the C++ oracle links.

nalgebra 0.16 semantics relied on (its source is not under /root/reference): `dot` of 3-vectors is
`(a0*b0 + a1*b1) + a2*b2`; `cross` is the component formula; `normalize` is `self / self.norm()` with
`norm = sqrt(dot(self, self))`; vector +, -, scalar *, / are component-wise.

Small images only: a pixel sample costs about 50 microseconds here.
"""
from __future__ import annotations

import math

from flux_b200.scene import (Emissive, GlossyReflective, Matte, PlaneData, Reflective, SceneData, SphereData)

T_MIN = 0.0005                 # fluxcore/src/constants.rs:4
INV_PI = 1.0 / math.pi         # fluxcore/src/constants.rs:5


# ---- nalgebra Vector3 / Point3 as tuples --------------------------------------------------------------------------
def dot(a, b):
    return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]


def cross(a, b):
    return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])


def add(a, b):
    return (a[0] + b[0], a[1] + b[1], a[2] + b[2])


def sub(a, b):
    return (a[0] - b[0], a[1] - b[1], a[2] - b[2])


def scale(a, s):
    return (a[0] * s, a[1] * s, a[2] * s)


def _div(a, b):
    """IEEE a / b, where Python raises ZeroDivisionError."""
    if b == 0.0:
        if a == 0.0 or a != a:
            return math.nan
        return math.copysign(math.inf, a) * math.copysign(1.0, b)
    return a / b


def divide(a, s):
    return (_div(a[0], s), _div(a[1], s), _div(a[2], s))


def neg(a):
    return (-a[0], -a[1], -a[2])


def normalize(a):
    return divide(a, math.sqrt(dot(a, a)))


# ---- Color (fluxcore/src/color.rs:46-105) -----------------------------------------------------------------------------
def c_mul(a, b):
    return (a[0] * b[0], a[1] * b[1], a[2] * b[2])


BLACK = (0.0, 0.0, 0.0)


class Hit:
    """common.rs:6-13"""
    __slots__ = ("local_hit_point", "normal", "material", "distance", "ray_origin", "ray_direction", "depth", "shape")


def min_(a, b):     # shapes.rs:90-92
    return a if a < b else b


def max_(a, b):     # shapes.rs:94-96
    return a if a > b else b


def bbox_hit(corner0, corner1, o, d):
    """BoundingBox::hit, shapes.rs:98-133.  1.0 / 0.0 is +inf in Rust, an exception in Python: spelled out."""
    def inv(x):
        return _div(1.0, x)
    a = inv(d[0])
    if a >= 0.0:
        tx_min, tx_max = (corner0[0] - o[0]) * a, (corner1[0] - o[0]) * a
    else:
        tx_min, tx_max = (corner1[0] - o[0]) * a, (corner0[0] - o[0]) * a
    b = inv(d[1])
    if b >= 0.0:
        ty_min, ty_max = (corner0[1] - o[1]) * b, (corner1[1] - o[1]) * b
    else:
        ty_min, ty_max = (corner1[1] - o[1]) * b, (corner0[1] - o[1]) * b
    c = inv(d[2])
    if c >= 0.0:
        tz_min, tz_max = (corner0[2] - o[2]) * c, (corner1[2] - o[2]) * c
    else:
        tz_min, tz_max = (corner1[2] - o[2]) * c, (corner0[2] - o[2]) * c
    t0 = max_(tx_min, max_(ty_min, tz_min))
    t1 = min_(tx_max, min_(ty_max, tz_max))
    return t0 < t1 and t1 > T_MIN


class Sphere:
    def __init__(self, data: SphereData, index: int):
        # Sphere::new, shapes.rs:154-169
        self.center = tuple(float(x) for x in data.center)
        self.radius = float(data.radius)
        self.invert = bool(data.invert)
        self.material = data.material
        delta = (self.radius, self.radius, self.radius)
        self.corner0 = sub(self.center, delta)
        self.corner1 = add(self.center, delta)
        self.index = index

    def hit(self, o, d, depth):
        """shapes.rs:171-217"""
        if not bbox_hit(self.corner0, self.corner1, o, d):
            return None
        temp = sub(o, self.center)
        a = dot(d, d)
        b = 2.0 * dot(temp, d)
        c = dot(temp, temp) - self.radius * self.radius
        disc = b * b - 4.0 * a * c
        invert_val = -1.0 if self.invert else 1.0
        if disc < 0.0:
            return None
        e = math.sqrt(disc)
        denom = 2.0 * a
        t = _div(-b - e, denom)
        if not t > T_MIN:
            t = _div(-b + e, denom)
            if not t > T_MIN:
                return None
        h = Hit()
        h.ray_origin, h.ray_direction, h.depth, h.distance = o, d, depth, t
        h.normal = divide(scale(add(temp, scale(d, t)), invert_val), self.radius)
        h.local_hit_point = add(o, scale(d, t))
        h.material = self.material
        h.shape = self.index
        return h


class Plane:
    def __init__(self, data: PlaneData, index: int):
        self.point = tuple(float(x) for x in data.point)
        self.normal = tuple(float(x) for x in data.normal)
        self.material = data.material
        self.index = index

    def hit(self, o, d, depth):
        """shapes.rs:135-152"""
        den = dot(d, self.normal)
        num = dot(sub(self.point, o), self.normal)
        t = _div(num, den)
        if not t > T_MIN:
            return None
        h = Hit()
        h.ray_origin, h.ray_direction, h.depth, h.distance = o, d, depth, t
        h.normal = self.normal
        h.local_hit_point = add(o, scale(d, t))
        h.material = self.material
        h.shape = self.index
        return h


def to_unit_hemi(px, py, e):
    """samplers/src/lib.rs:133-142"""
    cos_phi = math.cos(2.0 * math.pi * px)
    sin_phi = math.sin(2.0 * math.pi * px)
    cos_theta = math.pow(1.0 - py, 1.0 / (e + 1.0))
    sin_theta = math.sqrt(1.0 - cos_theta * cos_theta)
    return normalize((sin_theta * cos_phi, sin_theta * sin_phi, cos_theta))


def to_poisson_disc(px, py):
    """samplers/src/lib.rs:144-182 (Shirley's concentric map, with the reference's `spy != 0.0` guard)"""
    spx = 2.0 * px - 1.0
    spy = 2.0 * py - 1.0
    if spx > -spy:
        if spx > spy:
            r = spx
            phi = _div(spy, spx)
        else:
            r = spy
            phi = 2.0 - _div(spx, spy)
    else:
        if spx < spy:
            r = -spx
            phi = 4.0 + _div(spy, spx)
        else:
            r = -spy
            if spy != 0.0:
                phi = 6.0 - _div(spx, spy)
            else:
                phi = 0.0
    phi *= math.pi / 4.0
    return (r * math.cos(phi), r * math.sin(phi))


class Scene:
    """Scene::from_data, scene.rs:128-154: the shapes in the order of the scene file."""

    def __init__(self, sd: SceneData, max_trace_depth: int):
        self.shapes = []
        for i, sh in enumerate(sd.shapes):
            if isinstance(sh, SphereData):
                self.shapes.append(Sphere(sh, i))
            elif isinstance(sh, PlaneData):
                self.shapes.append(Plane(sh, i))
            else:
                raise TypeError("the reference has spheres and planes only (scene.rs:71-74)")
        self.background = tuple(float(x) for x in sd.background)
        self.max_trace_depth = max_trace_depth
        # CameraBasis::new, scene.rs:29-34
        s = sd.camera_settings
        self.eye = tuple(float(x) for x in s.eye)
        self.w = normalize(sub(self.eye, tuple(float(x) for x in s.look_at)))
        self.u = normalize(cross(tuple(float(x) for x in s.up), self.w))
        self.v = cross(self.w, self.u)

    def hit(self, o, d, depth):
        """scene.rs:156-160: filter_map + min_by(Hit::compare).  `min_by` keeps its accumulator unless the
        comparison says Greater, and Hit::compare (common.rs:17-23) says Less whenever acc.distance <= other.distance."""
        best = None
        for shape in self.shapes:
            h = shape.hit(o, d, depth)
            if h is None:
                continue
            if best is None or not (best.distance <= h.distance):
                best = h
        return best

    def shade(self, o, d, depth, samples, set_index, sample_index):
        """scene.rs:162-172"""
        if depth > self.max_trace_depth:
            return BLACK
        h = self.hit(o, d, depth)
        if h is None:
            return self.background
        return path_shade(h.material, self, h, samples, set_index, sample_index)


def lambertian_sample_f(cd, kd, hit, hemi_sample):
    """brdf.rs:20-30"""
    w = hit.normal
    v = normalize(cross((0.0034, 1.0, 0.0071), w))
    u = cross(v, w)
    wi = normalize(add(add(scale(u, hemi_sample[0]), scale(v, hemi_sample[1])), scale(w, hemi_sample[2])))
    pdf = dot(hit.normal, wi) * INV_PI
    return wi, pdf, scale(scale(cd, kd), INV_PI)


def perfect_specular_sample_f(cr, kr, hit, wo):
    """brdf.rs:39-45"""
    ndotwo = dot(hit.normal, wo)
    wi = add(neg(wo), scale(scale(hit.normal, ndotwo), 2.0))
    pdf = dot(hit.normal, wi)
    return wi, pdf, scale(cr, kr)


def glossy_specular_sample_f(cs, ks, exp, hit, wo, pixel_sample):
    """brdf.rs:55-78"""
    ndotwo = dot(hit.normal, wo)
    r = add(neg(wo), scale(scale(hit.normal, ndotwo), 2.0))
    w = r
    u = normalize(cross((0.00424, 1.0, 0.00764), w))
    v = cross(u, w)
    hs = to_unit_hemi(pixel_sample[0], pixel_sample[1], exp)
    wi0 = add(add(scale(u, hs[0]), scale(v, hs[1])), scale(w, hs[2]))
    if dot(hit.normal, wi0) < 0.0:
        wi = add(sub(scale(u, -hs[0]), scale(v, hs[1])), scale(w, hs[2]))
    else:
        wi = wi0
    phong_lobe = _powf(dot(r, wi), exp)
    pdf = phong_lobe * dot(hit.normal, wi)
    return wi, pdf, scale(scale(cs, ks), phong_lobe)


def _powf(x, y):
    """f64::powf = C pow(): NaN for a negative base with a non-integer exponent, where Python raises."""
    try:
        return math.pow(x, y)
    except (ValueError, OverflowError):
        if x < 0.0 and y != math.floor(y):
            return math.nan
        return math.inf


def path_shade(material, scene, hit, samples, set_index, sample_index):
    pixel_sets, _disc_sets, hemi_sets = samples
    if isinstance(material, Emissive):
        # materials.rs:42-49
        if dot(scale(hit.normal, -1.0), hit.ray_direction) > 0.0:
            return scale(tuple(float(x) for x in material.color), float(material.power))
        return BLACK
    if isinstance(material, Matte):
        # materials.rs:19-33 (the ambient BRDF and `wo` are never used)
        hemi_sample = hemi_sets[set_index][hit.depth - 1][sample_index]
        cd = tuple(float(x) for x in material.diffuse_color)
        wi, pdf, f = lambertian_sample_f(cd, float(material.diffuse_coefficient), hit, hemi_sample)
        ndotwi = dot(hit.normal, wi)
        child = scene.shade(hit.local_hit_point, wi, hit.depth + 1, samples, set_index, sample_index)
        return scale(c_mul(f, child), _div(ndotwi, pdf))
    if isinstance(material, (Reflective, GlossyReflective)):
        # materials.rs:57-71 with the BRDF material_from_data gives it (scene.rs:87-122)
        wo = scale(hit.ray_direction, -1.0)
        sq_sample = pixel_sets[set_index][sample_index]
        c = tuple(float(x) for x in material.reflect_color)
        if isinstance(material, Reflective):
            wi, pdf, fr = perfect_specular_sample_f(c, float(material.reflect_amount), hit, wo)
        else:
            wi, pdf, fr = glossy_specular_sample_f(c, float(material.reflect_amount), float(material.reflect_exponent),
                                                   hit, wo, sq_sample)
        child = scene.shade(hit.local_hit_point, wi, hit.depth + 1, samples, set_index, sample_index)
        return scale(c_mul(fr, child), _div(dot(hit.normal, wi), pdf))
    raise TypeError(f"not a MaterialData: {material!r}")


def ray_direction(scene, cam, px, py, lx, ly):
    """Camera::ray_direction, trace.rs:44-51"""
    factor = cam.focal_distance / cam.view_plane_distance
    px2 = px * factor
    py2 = py * factor
    return normalize(sub(add(scale(scene.u, px2 - lx), scale(scene.v, py2 - ly)), scale(scene.w, cam.focal_distance)))


def render_rows(sd: SceneData, sample_root: int, max_trace_depth: int, pixel_sets, disc_sets, hemi_sets, set_index, rows):
    """Camera::render, trace.rs:53-97, for the given rows.  Sample sets in the reference layout
    (pixel_sets[set][i] = (x, y), disc_sets[set][i], hemi_sets[set][depth - 1][i]); `set_index[row][col]` stands for the
    per-row shuffle (sampling.rs:35-40), which the reference draws from an unseeded generator.  Returns
    rows x width x (r, g, b) as nested lists."""
    scene = Scene(sd, max_trace_depth)
    cam = sd.camera_data
    img_h, img_w = sd.output_settings.image_height, sd.output_settings.image_width
    half_img_h = float(img_h) * 0.5
    half_img_w = float(img_w) * 0.5
    pixel_denom = 1.0 / float(sample_root * sample_root)
    adjusted_pixel_size = sd.output_settings.pixel_size / cam.zoom_factor
    # plain Python floats: numpy scalars would work too, but slower and with warnings instead of IEEE silence
    P = [[(float(p[0]), float(p[1])) for p in s] for s in pixel_sets]
    D = [[(float(p[0]), float(p[1])) for p in s] for s in disc_sets]
    Hm = [[[(float(p[0]), float(p[1]), float(p[2])) for p in lvl] for lvl in s] for s in hemi_sets]
    samples = (P, D, Hm)
    out = []
    for row in rows:
        row_pixels = []
        for col in range(img_w):
            color = BLACK
            si = int(set_index[row][col])
            pixel_samples = P[si % len(P)]
            disc_samples = D[si % len(D)]
            for index, point in enumerate(pixel_samples):
                u = adjusted_pixel_size * (float(col) - half_img_w + point[0])
                v = adjusted_pixel_size * (float(img_h - row) - half_img_h + point[1])
                lens_sample = disc_samples[index]
                lpx = lens_sample[0] * cam.lens_radius
                lpy = lens_sample[1] * cam.lens_radius
                direction = ray_direction(scene, cam, u, v, lpx, lpy)
                origin = add(add(scene.eye, scale(scene.u, lpx)), scale(scene.v, lpy))
                color = add(color, scene.shade(origin, direction, 1, samples, si, index))
            color = scale(color, pixel_denom)
            # Color::max_to_one, color.rs:35-44
            mx1 = color[0] if color[0] > color[1] else color[1]
            mx2 = mx1 if mx1 > color[2] else color[2]
            if mx2 > 1.0:
                i = 1.0 / mx2
                color = scale(color, i)
            row_pixels.append(color)
        out.append(row_pixels)
    return out


def trace(sd: SceneData, origins, dirs):
    """Scene::hit on explicit rays: (shape index or -1, distance) per ray."""
    scene = Scene(sd, 1)
    res = []
    for o, d in zip(origins, dirs):
        h = scene.hit(tuple(float(x) for x in o), tuple(float(x) for x in d), 1)
        res.append((-1, math.inf) if h is None else (h.shape, h.distance))
    return res


# ---- sample-set structure (samplers/src/lib.rs:46-126), on this repository's counter-based random source -------------------------
# The reference draws from one unseeded IsaacRng in sequence; this repository replaces the SOURCE of the random numbers
# by keyed counters (explicit, seeded sample sets) and keeps the STRUCTURE.  What is restated a second time
# here is that structure — the base grid, shuffle_y / shuffle_x and the transposes between them, written as the reference
# writes them instead of the closed form `out[i*root+j] = {base[pix_j(i)][j].x, base[i][piy_i(j)].y}` the C++ uses — fed
# from the same keyed source (mix64 / stream_key / rnd below restate oracle/flux_oracle.cpp's, not the reference's).
_M = (1 << 64) - 1
P_JITTER, P_PERM_Y, P_PERM_X = 0, 1, 2


def _mix64(z):
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M
    return z ^ (z >> 31)


def _stream_key(seed, a, b, c, d):
    k = _mix64((seed + 0x9E3779B97F4A7C15) & _M)
    for x in (a, b, c, d):
        k = _mix64((k + x) & _M)
    return k


def _rnd(key, ctr):
    return _mix64((key + (ctr + 1) * 0x9E3779B97F4A7C15) & _M)


def _u01(x):
    return float(x >> 11) * (1.0 / 9007199254740992.0)


def _shuffled_identity(n, key):
    """Rng::shuffle (rand 0.5.5) of 0..n: for i = n-1 down to 1: swap(i, gen_range(0, i+1))"""
    v = list(range(n))
    for i in range(n - 1, 0, -1):
        j = (_rnd(key, i) * (i + 1)) >> 64
        v[i], v[j] = v[j], v[i]
    return v


def _transpose(m):
    return [list(col) for col in zip(*m)]


def _shuffle_y(idxs, vals):     # lib.rs:92-108
    return [(sample[0], vals[idx][1]) for idx, sample in zip(idxs, vals)]


def _shuffle_x(idxs, vals):     # lib.rs:110-126
    return [(vals[idx][0], sample[1]) for idx, sample in zip(idxs, vals)]


def mj_grid(seed, root, set_, grid, correlated):
    """grid_multi_jittered (lib.rs:64-73) / grid_correlated_multi_jittered (lib.rs:75-90): root*root (x, y) tuples."""
    r2 = float(root * root)
    r_float = float(root)
    kj = _stream_key(seed, set_, grid, P_JITTER, 0)
    rng_range = [(float(v), float(root - 1 - v)) for v in range(root)]
    samples = []                # grid_multi_jittered_base, lib.rs:46-62
    for i, (big_row, little_col) in enumerate(rng_range):
        row = []
        for j, (big_col, little_row) in enumerate(rng_range):
            c = 2 * (i * root + j)
            a, b = _u01(_rnd(kj, c)), _u01(_rnd(kj, c + 1))
            row.append(((big_row / r_float) + (little_row + a) / r2, (big_col / r_float) + (little_col + b) / r2))
        samples.append(row)

    def idx(kind, line):
        return _shuffled_identity(root, _stream_key(seed, set_, grid, kind, 0 if correlated else line))

    y_shuffled = [_shuffle_y(idx(P_PERM_Y, line), vec) for line, vec in enumerate(samples)]
    x_shuffled = _transpose([_shuffle_x(idx(P_PERM_X, line), v) for line, v in enumerate(_transpose(y_shuffled))])
    return [p for row in x_shuffled for p in row]     # concat_vec
