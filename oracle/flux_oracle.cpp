// flux_oracle.cpp — CPU restatement of the reference's per-pixel render loop.
//
// TEST INFRASTRUCTURE ONLY.  This file is the parity checker and the CPU
// baseline for jtdaugherty/flux's hot path.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load it.  The product (libfluxb200.so) never links, imports or calls it.
//
// PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures
// and cannot be built here (no Rust toolchain, crates not vendored), so this
// restatement is pinned only by (i) hand-derived known-answer vectors in
// tests/test_oracle_known_answers.py and (ii) statistically, by the RMSE of a
// converged demo2 render against the reference's own demo.png.
// (iii) oracle/second_opinion.py, a second restatement written from the Rust
// sources independently of this one, agrees with it bit for bit on whole
// renders (tests/test_second_opinion.py): no transcription slip, still no pin.
//
// Third-party arithmetic that is not under /root/reference and is restated
// from its published behaviour: nalgebra 0.16.10 (Cargo.lock) Vector3/Point3
// `dot` = (a0*b0 + a1*b1) + a2*b2, `cross` = standard component formula,
// `normalize` = component-wise division by sqrt(dot(v,v)), scalar ops
// element-wise; rand 0.5.5 IsaacRng/thread_rng (NOT reproduced: the reference
// is unseeded, so sample sets and row permutations are explicit inputs here
// and come from the counter-based PRNG documented in DESIGN.md).
//
// Build: see oracle/Makefile (-O3 -ffp-contract=off -fno-fast-math: rustc never
// contracts a*b+c into an FMA, so neither may we).
//
// Every function cites the reference file:line it follows.

#include "../include/fluxb200.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// fluxcore/src/constants.rs:4-5
constexpr double T_MIN = 0.0005;
constexpr double PI = 3.14159265358979323846264338327950288;  // std::f64::consts::PI
constexpr double INV_PI = 1.0 / PI;

// ---- nalgebra 0.16.10 Vector3<f64> semantics -------------------------------
struct V3 {
    double x, y, z;
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline V3 operator/(V3 a, double s) { return {a.x / s, a.y / s, a.z / s}; }
inline V3 neg(V3 a) { return {-a.x, -a.y, -a.z}; }
inline double dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline V3 cross(V3 a, V3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline V3 normalize(V3 a) { return a / std::sqrt(dot(a, a)); }
inline V3 ld3(const double *p) { return {p[0], p[1], p[2]}; }
inline void st3(double *p, V3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }

// fluxcore/src/color.rs:8-12,46-105
struct Color {
    double r, g, b;
};
inline Color operator*(Color a, Color b) { return {a.r * b.r, a.g * b.g, a.b * b.b}; }
inline Color operator*(Color a, double s) { return {a.r * s, a.g * s, a.b * s}; }
inline Color black() { return {0.0, 0.0, 0.0}; }

// fluxcore/src/color.rs:35-44
inline void max_to_one(Color &c) {
    double mx1 = c.r > c.g ? c.r : c.g;
    double mx2 = mx1 > c.b ? mx1 : c.b;
    if (mx2 > 1.0) {
        double i = 1.0 / mx2;
        c.r *= i;
        c.g *= i;
        c.b *= i;
    }
}

// fluxcore/src/common.rs:26-30
struct Ray {
    V3 origin, direction;
};

// fluxcore/src/common.rs:7-14 (material reference → material index + shape id)
struct Hit {
    V3 local_hit_point;
    V3 normal;
    uint32_t material;
    int32_t shape_id;
    double distance;
    Ray ray;
    uint32_t depth;
};

struct Counters {
    uint64_t c[sizeof(flux_counters) / sizeof(uint64_t)] = {};
    flux_counters *as() { return reinterpret_cast<flux_counters *>(c); }
};

// fluxcore/src/shapes.rs:90-96 (private min/max: second argument on NaN)
inline double mn(double a, double b) { return a < b ? a : b; }
inline double mx(double a, double b) { return a > b ? a : b; }

// BoundingBox::hit, fluxcore/src/shapes.rs:98-133
inline bool bbox_hit(V3 corner0, V3 corner1, const Ray &r) {
    double ox = r.origin.x, oy = r.origin.y, oz = r.origin.z;
    double dx = r.direction.x, dy = r.direction.y, dz = r.direction.z;
    double tx_min, tx_max, ty_min, ty_max, tz_min, tz_max;
    double a = 1.0 / dx;
    if (a >= 0.0) {
        tx_min = (corner0.x - ox) * a;
        tx_max = (corner1.x - ox) * a;
    } else {
        tx_min = (corner1.x - ox) * a;
        tx_max = (corner0.x - ox) * a;
    }
    double b = 1.0 / dy;
    if (b >= 0.0) {
        ty_min = (corner0.y - oy) * b;
        ty_max = (corner1.y - oy) * b;
    } else {
        ty_min = (corner1.y - oy) * b;
        ty_max = (corner0.y - oy) * b;
    }
    double c = 1.0 / dz;
    if (c >= 0.0) {
        tz_min = (corner0.z - oz) * c;
        tz_max = (corner1.z - oz) * c;
    } else {
        tz_min = (corner1.z - oz) * c;
        tz_max = (corner0.z - oz) * c;
    }
    double t0 = mx(tx_min, mx(ty_min, tz_min));
    double t1 = mn(tx_max, mn(ty_max, tz_max));
    return t0 < t1 && t1 > T_MIN;
}

struct SphereS {
    V3 center;
    double radius;
    bool invert;
    uint32_t material;
    int32_t shape_id;
    V3 corner0, corner1;  // Sphere::new, shapes.rs:154-169
};
struct PlaneS {
    V3 point, normal;
    uint32_t material;
    int32_t shape_id;
};
// EXTENSION (not in the reference; DESIGN.md "extensions"): two-sided
// Moller-Trumbore triangle, geometric normal as wound, never flipped.
struct TriS {
    V3 v0, v1, v2;
    uint32_t material;
    int32_t shape_id;
};

// Sphere::hit, fluxcore/src/shapes.rs:171-217
inline bool sphere_hit(const SphereS &s, const Ray &r, uint32_t depth, Hit &h, Counters *cn) {
    if (cn) cn->as()->bbox_tests++;
    if (!bbox_hit(s.corner0, s.corner1, r)) return false;
    if (cn) cn->as()->bbox_pass++;
    V3 temp = r.origin - s.center;
    double a = dot(r.direction, r.direction);
    double b = 2.0 * dot(temp, r.direction);
    double c = dot(temp, temp) - s.radius * s.radius;
    double disc = b * b - 4.0 * a * c;
    double invert_val = s.invert ? -1.0 : 1.0;
    if (disc < 0.0) return false;
    if (cn) cn->as()->disc_nonneg++;
    double e = std::sqrt(disc);
    double denom = 2.0 * a;
    double t = (-b - e) / denom;
    if (!(t > T_MIN)) {
        if (cn) cn->as()->t2_evals++;
        t = (-b + e) / denom;
        if (!(t > T_MIN)) return false;
    }
    h.ray = r;
    h.distance = t;
    h.depth = depth;
    h.normal = ((temp + t * r.direction) * invert_val) / s.radius;
    h.local_hit_point = r.origin + t * r.direction;
    h.material = s.material;
    h.shape_id = s.shape_id;
    return true;
}

// Plane::hit, fluxcore/src/shapes.rs:135-152
inline bool plane_hit(const PlaneS &p, const Ray &r, uint32_t depth, Hit &h, Counters *cn) {
    if (cn) cn->as()->plane_tests++;
    double t = dot(p.point - r.origin, p.normal) / dot(r.direction, p.normal);
    if (t > T_MIN) {
        h.ray = r;
        h.depth = depth;
        h.distance = t;
        h.normal = p.normal;
        h.local_hit_point = r.origin + t * r.direction;
        h.material = p.material;
        h.shape_id = p.shape_id;
        return true;
    }
    return false;
}

// EXTENSION: triangle (semantics defined by this repo, DESIGN.md "extensions").
inline bool tri_hit(const TriS &tr, const Ray &r, uint32_t depth, Hit &h, Counters *cn) {
    if (cn) cn->as()->tri_tests++;
    V3 e1 = tr.v1 - tr.v0;
    V3 e2 = tr.v2 - tr.v0;
    V3 p = cross(r.direction, e2);
    double det = dot(e1, p);
    if (det == 0.0) return false;
    double inv = 1.0 / det;
    V3 s = r.origin - tr.v0;
    double u = dot(s, p) * inv;
    if (!(u >= 0.0 && u <= 1.0)) return false;
    V3 q = cross(s, e1);
    double v = dot(r.direction, q) * inv;
    if (!(v >= 0.0 && u + v <= 1.0)) return false;
    double t = dot(e2, q) * inv;
    if (!(t > T_MIN)) return false;
    h.ray = r;
    h.depth = depth;
    h.distance = t;
    h.normal = normalize(cross(e1, e2));
    h.local_hit_point = r.origin + t * r.direction;
    h.material = tr.material;
    h.shape_id = tr.shape_id;
    return true;
}

struct ShapeRef {
    int32_t shape_id;
    uint8_t kind;  // 0 sphere, 1 plane, 2 triangle
    uint32_t index;
};

// fluxcore/src/sampling.rs:5-10
struct SampleSets {
    uint32_t root = 0, n = 0, max_depth = 0, num_sets = 0;
    const double *pixel = nullptr;  // [set][i][2]
    const double *disc = nullptr;   // [set][i][2]
    const double *hemi = nullptr;   // [set][depth][i][3]
};

// fluxcore/src/scene.rs:76-85 + Camera (trace.rs:14-23)
struct Scene {
    uint32_t W = 0, H = 0;
    double pixel_size = 0;
    Color background{};
    V3 eye{}, u{}, v{}, w{};
    double zoom_factor = 0, view_plane_distance = 0, focal_distance = 0, lens_radius = 0;
    uint32_t max_trace_depth = 0;
    std::vector<flux_material> materials;
    std::vector<SphereS> spheres;
    std::vector<PlaneS> planes;
    std::vector<TriS> tris;
    std::vector<ShapeRef> order;  // YAML order
};

// CameraBasis::new, fluxcore/src/scene.rs:29-34
inline void camera_basis(V3 eye, V3 look_at, V3 up, V3 &u, V3 &v, V3 &w) {
    w = normalize(eye - look_at);
    u = normalize(cross(up, w));
    v = cross(w, u);
}

// Scene::from_data, fluxcore/src/scene.rs:128-154
bool build_scene(const flux_scene_flat *f, uint32_t max_trace_depth, Scene &s) {
    if (!f) return false;
    s.W = f->image_width;
    s.H = f->image_height;
    s.pixel_size = f->pixel_size;
    s.background = {f->background[0], f->background[1], f->background[2]};
    s.eye = ld3(f->eye);
    camera_basis(s.eye, ld3(f->look_at), ld3(f->up), s.u, s.v, s.w);
    s.zoom_factor = f->zoom_factor;
    s.view_plane_distance = f->view_plane_distance;
    s.focal_distance = f->focal_distance;
    s.lens_radius = f->lens_radius;
    s.max_trace_depth = max_trace_depth;
    s.materials.assign(f->materials, f->materials + f->n_materials);
    for (uint32_t i = 0; i < f->n_spheres; i++) {
        SphereS sp;
        sp.center = ld3(f->sphere_center + 3 * i);
        sp.radius = f->sphere_radius[i];
        sp.invert = f->sphere_invert[i] != 0;
        sp.material = f->sphere_material[i];
        sp.shape_id = (int32_t)f->sphere_shape_id[i];
        V3 delta = {sp.radius, sp.radius, sp.radius};  // shapes.rs:156-158
        sp.corner0 = sp.center - delta;
        sp.corner1 = sp.center + delta;
        if (sp.material >= f->n_materials) return false;
        s.order.push_back({sp.shape_id, 0, (uint32_t)s.spheres.size()});
        s.spheres.push_back(sp);
    }
    for (uint32_t i = 0; i < f->n_planes; i++) {
        PlaneS p;
        p.point = ld3(f->plane_point + 3 * i);
        p.normal = ld3(f->plane_normal + 3 * i);
        p.material = f->plane_material[i];
        p.shape_id = (int32_t)f->plane_shape_id[i];
        if (p.material >= f->n_materials) return false;
        s.order.push_back({p.shape_id, 1, (uint32_t)s.planes.size()});
        s.planes.push_back(p);
    }
    for (uint32_t i = 0; i < f->n_triangles; i++) {
        TriS t;
        t.v0 = ld3(f->tri_v0 + 3 * i);
        t.v1 = ld3(f->tri_v1 + 3 * i);
        t.v2 = ld3(f->tri_v2 + 3 * i);
        t.material = f->tri_material[i];
        t.shape_id = (int32_t)f->tri_shape_id[i];
        if (t.material >= f->n_materials) return false;
        s.order.push_back({t.shape_id, 2, (uint32_t)s.tris.size()});
        s.tris.push_back(t);
    }
    std::stable_sort(s.order.begin(), s.order.end(),
                     [](const ShapeRef &a, const ShapeRef &b) { return a.shape_id < b.shape_id; });
    return true;
}

// Scene::hit, fluxcore/src/scene.rs:156-160 with Hit::compare, common.rs:17-23:
// Iterator::min_by keeps the accumulated element unless compare(acc, cand) is
// Greater, i.e. unless !(acc.distance <= cand.distance): ties keep the earlier shape.
inline bool scene_hit(const Scene &s, const Ray &r, uint32_t depth, Hit &best, Counters *cn) {
    bool have = false;
    Hit cand;
    for (const ShapeRef &sr : s.order) {
        bool ok;
        if (sr.kind == 0)
            ok = sphere_hit(s.spheres[sr.index], r, depth, cand, cn);
        else if (sr.kind == 1)
            ok = plane_hit(s.planes[sr.index], r, depth, cand, cn);
        else
            ok = tri_hit(s.tris[sr.index], r, depth, cand, cn);
        if (!ok) continue;
        if (cn) cn->as()->candidates++;
        if (!have) {
            best = cand;
            have = true;
        } else if (!(best.distance <= cand.distance)) {
            best = cand;
        }
    }
    return have;
}

// to_unit_hemi, samplers/src/lib.rs:133-142
inline V3 to_unit_hemi(double px, double py, double e) {
    double cos_phi = std::cos(2.0 * PI * px);
    double sin_phi = std::sin(2.0 * PI * px);
    double cos_theta = std::pow(1.0 - py, 1.0 / (e + 1.0));
    double sin_theta = std::sqrt(1.0 - cos_theta * cos_theta);
    double pu = sin_theta * cos_phi;
    double pv = sin_theta * sin_phi;
    double pw = cos_theta;
    return normalize(V3{pu, pv, pw});
}

// to_poisson_disc (one point), samplers/src/lib.rs:144-182
inline void to_disc(double px, double py, double &ox, double &oy) {
    double spx = 2.0 * px - 1.0;
    double spy = 2.0 * py - 1.0;
    double phi, r;
    if (spx > -spy) {
        if (spx > spy) {
            r = spx;
            phi = spy / spx;
        } else {
            r = spy;
            phi = 2.0 - spx / spy;
        }
    } else {
        if (spx < spy) {
            r = -spx;
            phi = 4.0 + spy / spx;
        } else {
            r = -spy;
            if (spy != 0.0)
                phi = 6.0 - spx / spy;
            else
                phi = 0.0;
        }
    }
    phi *= PI / 4.0;
    ox = r * std::cos(phi);
    oy = r * std::sin(phi);
}

// Lambertian::sample_f, fluxcore/src/brdf.rs:20-30
inline void lambertian_sample_f(V3 normal, V3 hemi, Color cd, double kd, V3 &wi, double &pdf,
                                Color &f) {
    V3 w = normal;
    V3 v = normalize(cross(V3{0.0034, 1.0, 0.0071}, w));
    V3 u = cross(v, w);
    wi = normalize((hemi.x * u + hemi.y * v) + hemi.z * w);
    pdf = dot(normal, wi) * INV_PI;
    f = cd * kd * INV_PI;
}

// PerfectSpecular::sample_f, fluxcore/src/brdf.rs:39-45
inline void specular_sample_f(V3 normal, V3 wo, Color cr, double kr, V3 &wi, double &pdf,
                              Color &f) {
    double ndotwo = dot(normal, wo);
    wi = neg(wo) + normal * ndotwo * 2.0;
    pdf = dot(normal, wi);
    f = cr * kr;
}

// GlossySpecular::sample_f, fluxcore/src/brdf.rs:55-78
inline void glossy_sample_f(V3 normal, V3 wo, double sqx, double sqy, Color cs, double ks,
                            double ex, V3 &wi, double &pdf, Color &f, bool &flipped) {
    double ndotwo = dot(normal, wo);
    V3 r = neg(wo) + normal * ndotwo * 2.0;
    V3 w = r;
    V3 u = normalize(cross(V3{0.00424, 1.0, 0.00764}, w));
    V3 v = cross(u, w);
    V3 hs = to_unit_hemi(sqx, sqy, ex);
    V3 wi0 = (u * hs.x + v * hs.y) + w * hs.z;
    flipped = dot(normal, wi0) < 0.0;
    if (flipped)
        wi = (u * -hs.x - v * hs.y) + w * hs.z;
    else
        wi = wi0;
    double phong_lobe = std::pow(dot(r, wi), ex);
    pdf = phong_lobe * dot(normal, wi);
    f = cs * ks * phong_lobe;
}

Color shade(const Scene &s, const Ray &r, uint32_t depth, const SampleSets &ss, uint32_t set_index,
            uint32_t sample_index, Counters *cn);

// Material::path_shade, fluxcore/src/materials.rs:19-71
Color path_shade(const Scene &s, const Hit &hit, const SampleSets &ss, uint32_t set_index,
                 uint32_t sample_index, Counters *cn) {
    const flux_material &m = s.materials[hit.material];
    Color mc = {m.color[0], m.color[1], m.color[2]};
    switch (m.kind) {
    case FLUX_MAT_EMISSIVE: {  // materials.rs:42-49
        if (cn) cn->as()->emissive++;
        if (dot(hit.normal * -1.0, hit.ray.direction) > 0.0) return mc * m.k;
        return black();
    }
    case FLUX_MAT_MATTE: {  // materials.rs:19-33
        if (cn) cn->as()->matte++;
        const double *hp =
            ss.hemi + (((size_t)set_index * ss.max_depth + (hit.depth - 1)) * ss.n + sample_index) * 3;
        V3 wi;
        double pdf;
        Color f;
        lambertian_sample_f(hit.normal, ld3(hp), mc, m.k, wi, pdf, f);
        double ndotwi = dot(hit.normal, wi);
        Ray refl{hit.local_hit_point, wi};
        return f * shade(s, refl, hit.depth + 1, ss, set_index, sample_index, cn) * (ndotwi / pdf);
    }
    case FLUX_MAT_REFLECTIVE:
    case FLUX_MAT_GLOSSY: {  // materials.rs:57-71
        V3 wo = hit.ray.direction * -1.0;
        const double *sq = ss.pixel + ((size_t)set_index * ss.n + sample_index) * 2;
        V3 wi;
        double pdf;
        Color fr;
        if (m.kind == FLUX_MAT_REFLECTIVE) {
            if (cn) cn->as()->specular++;
            specular_sample_f(hit.normal, wo, mc, m.k, wi, pdf, fr);
        } else {
            bool flipped;
            if (cn) cn->as()->glossy++;
            glossy_sample_f(hit.normal, wo, sq[0], sq[1], mc, m.k, m.exp, wi, pdf, fr, flipped);
            if (cn && flipped) cn->as()->glossy_flip++;
        }
        Ray refl{hit.local_hit_point, wi};
        return fr * shade(s, refl, hit.depth + 1, ss, set_index, sample_index, cn) *
               (dot(hit.normal, wi) / pdf);
    }
    }
    return black();
}

// Scene::shade, fluxcore/src/scene.rs:162-172
Color shade(const Scene &s, const Ray &r, uint32_t depth, const SampleSets &ss, uint32_t set_index,
            uint32_t sample_index, Counters *cn) {
    if (depth > s.max_trace_depth) {
        if (cn) cn->as()->depth_cut++;
        return black();
    }
    if (cn) cn->as()->segments++;
    Hit h;
    if (!scene_hit(s, r, depth, h, cn)) {
        if (cn) cn->as()->miss++;
        return s.background;
    }
    if (cn) {
        const ShapeRef *sr = nullptr;
        for (const ShapeRef &x : s.order)
            if (x.shape_id == h.shape_id) {
                sr = &x;
                break;
            }
        if (sr) {
            if (sr->kind == 0) cn->as()->hit_sphere++;
            else if (sr->kind == 1) cn->as()->hit_plane++;
            else cn->as()->hit_tri++;
        }
    }
    return path_shade(s, h, ss, set_index, sample_index, cn);
}

// Camera::ray_direction, fluxcore/src/trace.rs:44-51
inline V3 ray_direction(const Scene &s, double px, double py, double lx, double ly) {
    double factor = s.focal_distance / s.view_plane_distance;
    double px2 = px * factor;
    double py2 = py * factor;
    return normalize(((px2 - lx) * s.u + (py2 - ly) * s.v) - s.focal_distance * s.w);
}

// Body of the per-sample loop, fluxcore/src/trace.rs:71-80
inline Ray primary_ray(const Scene &s, uint32_t row, uint32_t col, double spx, double spy,
                       double ldx, double ldy) {
    double half_img_h = (double)s.H * 0.5;
    double half_img_w = (double)s.W * 0.5;
    double adjusted_pixel_size = s.pixel_size / s.zoom_factor;
    double u = adjusted_pixel_size * (((double)col - half_img_w) + spx);
    double v = adjusted_pixel_size * (((double)(s.H - row) - half_img_h) + spy);
    double lpx = ldx * s.lens_radius;
    double lpy = ldy * s.lens_radius;
    Ray r;
    r.direction = ray_direction(s, u, v, lpx, lpy);
    r.origin = (s.eye + lpx * s.u) + lpy * s.v;
    return r;
}

// One pixel of Camera::render, fluxcore/src/trace.rs:66-87
inline Color render_pixel(const Scene &s, const SampleSets &ss, uint32_t row, uint32_t col,
                          uint32_t set, Counters *cn) {
    double pixel_denom = 1.0 / (double)((uint64_t)ss.root * ss.root);
    Color color = black();
    const double *ps = ss.pixel + (size_t)set * ss.n * 2;
    const double *ds = ss.disc + (size_t)set * ss.n * 2;
    for (uint32_t index = 0; index < ss.n; index++) {
        Ray r = primary_ray(s, row, col, ps[2 * index], ps[2 * index + 1], ds[2 * index],
                            ds[2 * index + 1]);
        if (cn) cn->as()->samples++;
        Color c = shade(s, r, 1, ss, set, index, cn);
        color.r += c.r;
        color.g += c.g;
        color.b += c.b;
    }
    color.r *= pixel_denom;
    color.g *= pixel_denom;
    color.b *= pixel_denom;
    max_to_one(color);
    return color;
}

// ---- counter-based PRNG (this repo's; the reference's IsaacRng is unseeded) --
inline uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
inline uint64_t stream_key(uint64_t seed, uint64_t a, uint64_t b, uint64_t c, uint64_t d) {
    uint64_t k = mix64(seed + 0x9E3779B97F4A7C15ull);
    k = mix64(k + a);
    k = mix64(k + b);
    k = mix64(k + c);
    k = mix64(k + d);
    return k;
}
inline uint64_t rnd(uint64_t key, uint64_t ctr) {
    return mix64(key + (ctr + 1) * 0x9E3779B97F4A7C15ull);
}
inline double u01(uint64_t x) { return (double)(x >> 11) * (1.0 / 9007199254740992.0); }
inline uint32_t below(uint64_t x, uint32_t n) {
    return (uint32_t)(((unsigned __int128)x * (unsigned __int128)n) >> 64);
}
// Rng::shuffle (rand 0.5.5): for i = len-1 down to 1: swap(i, gen_range(0, i+1))
template <class T> inline void fisher_yates(T *v, uint32_t len, uint64_t key) {
    for (uint32_t i = len; i >= 2;) {
        i -= 1;
        uint32_t j = below(rnd(key, i), i + 1);
        T tmp = v[i];
        v[i] = v[j];
        v[j] = tmp;
    }
}
enum { P_JITTER = 0, P_PERM_Y = 1, P_PERM_X = 2, P_ROW = 3 };
enum { G_PIXEL = 0, G_DISC = 1, G_HEMI0 = 2 };

// grid_multi_jittered_base, samplers/src/lib.rs:46-62; out: base[i][j] {x,y}
void mj_base(uint32_t root, uint64_t seed, uint32_t set, uint32_t grid, std::vector<double> &base) {
    double r2 = (double)((uint64_t)root * root);
    double r_float = (double)root;
    uint64_t kj = stream_key(seed, set, grid, P_JITTER, 0);
    base.resize((size_t)root * root * 2);
    for (uint32_t i = 0; i < root; i++) {      // (big_row, little_col) = (i, root-1-i)
        for (uint32_t j = 0; j < root; j++) {  // (big_col, little_row) = (j, root-1-j)
            uint64_t c = 2ull * ((uint64_t)i * root + j);
            double a = u01(rnd(kj, c));
            double b = u01(rnd(kj, c + 1));
            double big_row = (double)i, little_col = (double)(root - 1 - i);
            double big_col = (double)j, little_row = (double)(root - 1 - j);
            base[((size_t)i * root + j) * 2 + 0] = (big_row / r_float) + (little_row + a) / r2;
            base[((size_t)i * root + j) * 2 + 1] = (big_col / r_float) + (little_col + b) / r2;
        }
    }
}

// grid_multi_jittered (correlated=false, lib.rs:64-73) and
// grid_correlated_multi_jittered (correlated=true, lib.rs:75-90):
// out[i*root+j] = { x: base[pix_j(i)][j].x, y: base[i][piy_i(j)].y }  (shuffle_y/x, lib.rs:92-126)
void mj_grid(uint32_t root, uint64_t seed, uint32_t set, uint32_t grid, bool correlated,
             double *out) {
    std::vector<double> base;
    mj_base(root, seed, set, grid, base);
    std::vector<uint32_t> py((size_t)root * root), pxm((size_t)root * root);
    for (uint32_t line = 0; line < root; line++) {
        uint32_t *y = &py[(size_t)line * root];
        uint32_t *x = &pxm[(size_t)line * root];
        for (uint32_t k = 0; k < root; k++) y[k] = x[k] = k;
        uint32_t l = correlated ? 0 : line;
        fisher_yates(y, root, stream_key(seed, set, grid, P_PERM_Y, l));
        fisher_yates(x, root, stream_key(seed, set, grid, P_PERM_X, l));
    }
    for (uint32_t i = 0; i < root; i++)
        for (uint32_t j = 0; j < root; j++) {
            uint32_t xi = pxm[(size_t)j * root + i];  // pix_j(i)
            uint32_t yj = py[(size_t)i * root + j];   // piy_i(j)
            out[((size_t)i * root + j) * 2 + 0] = base[((size_t)xi * root + j) * 2 + 0];
            out[((size_t)i * root + j) * 2 + 1] = base[((size_t)i * root + yj) * 2 + 1];
        }
}

}  // namespace

// =============================== C interface ==================================
extern "C" {

const char *oracle_version(void) { return "flux-oracle 1 (CPU restatement; parity unpinned)"; }

// Camera::render over [row_start, row_end] (trace.rs:53-97) with rows in parallel
// like rows.par_iter() (trace.rs:63).  num_threads <= 0: all cores.
int oracle_render_row_list(const flux_scene_flat *scene, const flux_job_config *cfg,
                           uint32_t num_sets, const double *pixel_xy, const double *disc_xy,
                           const double *hemi_xyz, const uint32_t *set_index, const uint32_t *rows,
                           uint32_t n_rows, double *out_rgb, flux_counters *counters,
                           int num_threads) {
    Scene s;
    if (!cfg || !build_scene(scene, cfg->max_trace_depth, s)) return FLUX_ERR_INVALID;
    SampleSets ss;
    ss.root = cfg->sample_root;
    ss.n = cfg->sample_root * cfg->sample_root;
    ss.max_depth = cfg->max_trace_depth;
    ss.num_sets = num_sets;
    ss.pixel = pixel_xy;
    ss.disc = disc_xy;
    ss.hemi = hemi_xyz;
    for (uint32_t k = 0; k < n_rows; k++)
        if (rows[k] >= s.H) return FLUX_ERR_INVALID;
#ifdef _OPENMP
    int nt = num_threads > 0 ? num_threads : omp_get_max_threads();
#else
    int nt = 1;
#endif
    std::vector<Counters> cns(counters ? nt : 0);
    const int64_t total = (int64_t)n_rows;
    const uint32_t W = s.W;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nt)
    for (int64_t k = 0; k < total; k++) {
#ifdef _OPENMP
        Counters *cn = counters ? &cns[omp_get_thread_num()] : nullptr;
#else
        Counters *cn = counters ? &cns[0] : nullptr;
#endif
        uint32_t row = rows[k];
        for (uint32_t col = 0; col < W; col++) {
            uint32_t set = set_index[(size_t)row * W + col] % num_sets;  // trace.rs:68-69
            Color c = render_pixel(s, ss, row, col, set, cn);
            double *o = out_rgb + ((size_t)k * W + col) * 3;
            o[0] = c.r;
            o[1] = c.g;
            o[2] = c.b;
        }
    }
    if (counters) {
        uint64_t *dst = reinterpret_cast<uint64_t *>(counters);
        for (auto &c : cns)
            for (size_t i = 0; i < sizeof(flux_counters) / sizeof(uint64_t); i++) dst[i] += c.c[i];
    }
    return FLUX_OK;
}

int oracle_render_rows(const flux_scene_flat *scene, const flux_job_config *cfg, uint32_t num_sets,
                       const double *pixel_xy, const double *disc_xy, const double *hemi_xyz,
                       const uint32_t *set_index, uint32_t row_start, uint32_t row_end_inclusive,
                       double *out_rgb, flux_counters *counters, int num_threads) {
    if (row_end_inclusive < row_start) return FLUX_ERR_INVALID;
    std::vector<uint32_t> rows;
    for (uint32_t r = row_start; r <= row_end_inclusive; r++) rows.push_back(r);
    return oracle_render_row_list(scene, cfg, num_sets, pixel_xy, disc_xy, hemi_xyz, set_index,
                                  rows.data(), (uint32_t)rows.size(), out_rgb, counters,
                                  num_threads);
}

// Scene::hit on explicit rays (scene.rs:156-160)
int oracle_trace_rays(const flux_scene_flat *scene, uint64_t n, const double *origin_xyz,
                      const double *dir_xyz, int32_t *hit_shape_id, double *t, int num_threads) {
    Scene s;
    if (!build_scene(scene, 1, s)) return FLUX_ERR_INVALID;
#ifdef _OPENMP
    int nt = num_threads > 0 ? num_threads : omp_get_max_threads();
#else
    int nt = 1;
#endif
    (void)nt;
#pragma omp parallel for schedule(static) num_threads(nt)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        Ray r{ld3(origin_xyz + 3 * i), ld3(dir_xyz + 3 * i)};
        Hit h;
        if (scene_hit(s, r, 1, h, nullptr)) {
            hit_shape_id[i] = h.shape_id;
            t[i] = h.distance;
        } else {
            hit_shape_id[i] = -1;
            t[i] = INFINITY;
        }
    }
    return FLUX_OK;
}

// Full hit record for one ray (unit tests): returns shape id or -1.
int oracle_hit_record(const flux_scene_flat *scene, const double *o, const double *d, double *t,
                      double *normal, double *point, uint32_t *material) {
    Scene s;
    if (!build_scene(scene, 1, s)) return -2;
    Ray r{ld3(o), ld3(d)};
    Hit h{};
    if (!scene_hit(s, r, 1, h, nullptr)) return -1;
    *t = h.distance;
    st3(normal, h.normal);
    st3(point, h.local_hit_point);
    *material = h.material;
    return h.shape_id;
}

// MasterSampleSets::new, fluxcore/src/sampling.rs:13-33
void oracle_generate_samples(uint64_t seed, uint32_t root, uint32_t max_depth, uint32_t num_sets,
                             double *pixel_xy, double *disc_xy, double *hemi_xyz) {
    const size_t n = (size_t)root * root;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t s = 0; s < (int64_t)num_sets; s++) {
        std::vector<double> g(n * 2);
        // pixel_sets: grid_correlated_multi_jittered (sampling.rs:16-17)
        mj_grid(root, seed, (uint32_t)s, G_PIXEL, true, pixel_xy + (size_t)s * n * 2);
        // disc_sets: to_poisson_disc(grid_correlated_multi_jittered) (sampling.rs:19-21)
        mj_grid(root, seed, (uint32_t)s, G_DISC, true, g.data());
        for (size_t i = 0; i < n; i++)
            to_disc(g[2 * i], g[2 * i + 1], disc_xy[((size_t)s * n + i) * 2],
                    disc_xy[((size_t)s * n + i) * 2 + 1]);
        // hemi_sets: to_hemisphere(grid_multi_jittered, 0.0) per depth (sampling.rs:23-29)
        for (uint32_t d = 0; d < max_depth; d++) {
            mj_grid(root, seed, (uint32_t)s, G_HEMI0 + d, false, g.data());
            for (size_t i = 0; i < n; i++) {
                V3 h = to_unit_hemi(g[2 * i], g[2 * i + 1], 0.0);
                st3(hemi_xyz + (((size_t)s * max_depth + d) * n + i) * 3, h);
            }
        }
    }
}

// Raw unit-square grids for stratification tests.
void oracle_mj_grid(uint64_t seed, uint32_t root, uint32_t set, uint32_t grid, int correlated,
                    double *out_xy) {
    mj_grid(root, seed, set, grid, correlated != 0, out_xy);
}

// shuffle_indices per row, fluxcore/src/sampling.rs:35-40 → idx[row][col] (trace.rs:64,68-69)
void oracle_generate_set_index(uint64_t seed, uint32_t image_height, uint32_t image_width,
                               uint32_t num_sets, uint32_t *idx) {
#pragma omp parallel for schedule(static)
    for (int64_t row = 0; row < (int64_t)image_height; row++) {
        std::vector<uint32_t> perm(num_sets);
        for (uint32_t k = 0; k < num_sets; k++) perm[k] = k;
        fisher_yates(perm.data(), num_sets, stream_key(seed, 0xFFFFFFFFull, (uint64_t)row, P_ROW, 0));
        for (uint32_t col = 0; col < image_width; col++)
            idx[(size_t)row * image_width + col] = perm[col % num_sets];
    }
}

// ---- unit-level entry points (tests/test_oracle_known_answers.py) ------------
void oracle_camera_basis(const double *eye, const double *look_at, const double *up, double *u,
                         double *v, double *w) {
    V3 U, V, W;
    camera_basis(ld3(eye), ld3(look_at), ld3(up), U, V, W);
    st3(u, U);
    st3(v, V);
    st3(w, W);
}

int oracle_bbox_hit(const double *c0, const double *c1, const double *o, const double *d) {
    Ray r{ld3(o), ld3(d)};
    return bbox_hit(ld3(c0), ld3(c1), r) ? 1 : 0;
}

int oracle_sphere_hit(const double *center, double radius, int invert, const double *o,
                      const double *d, double *t, double *normal, double *point) {
    SphereS s;
    s.center = ld3(center);
    s.radius = radius;
    s.invert = invert != 0;
    s.material = 0;
    s.shape_id = 0;
    V3 delta = {radius, radius, radius};
    s.corner0 = s.center - delta;
    s.corner1 = s.center + delta;
    Ray r{ld3(o), ld3(d)};
    Hit h;
    if (!sphere_hit(s, r, 1, h, nullptr)) return 0;
    *t = h.distance;
    st3(normal, h.normal);
    st3(point, h.local_hit_point);
    return 1;
}

int oracle_plane_hit(const double *p, const double *n, const double *o, const double *d, double *t,
                     double *normal, double *point) {
    PlaneS pl{ld3(p), ld3(n), 0, 0};
    Ray r{ld3(o), ld3(d)};
    Hit h;
    if (!plane_hit(pl, r, 1, h, nullptr)) return 0;
    *t = h.distance;
    st3(normal, h.normal);
    st3(point, h.local_hit_point);
    return 1;
}

int oracle_triangle_hit(const double *v0, const double *v1, const double *v2, const double *o,
                        const double *d, double *t, double *normal, double *point) {
    TriS tr{ld3(v0), ld3(v1), ld3(v2), 0, 0};
    Ray r{ld3(o), ld3(d)};
    Hit h;
    if (!tri_hit(tr, r, 1, h, nullptr)) return 0;
    *t = h.distance;
    st3(normal, h.normal);
    st3(point, h.local_hit_point);
    return 1;
}

void oracle_lambertian_sample_f(const double *normal, const double *hemi, const double *cd,
                                double kd, double *wi, double *pdf, double *f) {
    V3 WI;
    Color F;
    lambertian_sample_f(ld3(normal), ld3(hemi), Color{cd[0], cd[1], cd[2]}, kd, WI, *pdf, F);
    st3(wi, WI);
    f[0] = F.r;
    f[1] = F.g;
    f[2] = F.b;
}

void oracle_specular_sample_f(const double *normal, const double *wo, const double *cr, double kr,
                              double *wi, double *pdf, double *f) {
    V3 WI;
    Color F;
    specular_sample_f(ld3(normal), ld3(wo), Color{cr[0], cr[1], cr[2]}, kr, WI, *pdf, F);
    st3(wi, WI);
    f[0] = F.r;
    f[1] = F.g;
    f[2] = F.b;
}

int oracle_glossy_sample_f(const double *normal, const double *wo, const double *sq,
                           const double *cs, double ks, double ex, double *wi, double *pdf,
                           double *f) {
    V3 WI;
    Color F;
    bool flipped;
    glossy_sample_f(ld3(normal), ld3(wo), sq[0], sq[1], Color{cs[0], cs[1], cs[2]}, ks, ex, WI,
                    *pdf, F, flipped);
    st3(wi, WI);
    f[0] = F.r;
    f[1] = F.g;
    f[2] = F.b;
    return flipped ? 1 : 0;
}

void oracle_to_unit_hemi(double px, double py, double e, double *out) {
    st3(out, to_unit_hemi(px, py, e));
}

void oracle_to_poisson_disc(double px, double py, double *out) { to_disc(px, py, out[0], out[1]); }

void oracle_max_to_one(double *rgb) {
    Color c{rgb[0], rgb[1], rgb[2]};
    max_to_one(c);
    rgb[0] = c.r;
    rgb[1] = c.g;
    rgb[2] = c.b;
}

// Primary ray of (row, col) for one pixel/disc sample (trace.rs:72-80).
int oracle_primary_ray(const flux_scene_flat *scene, uint32_t row, uint32_t col, double spx,
                       double spy, double ldx, double ldy, double *o, double *d) {
    Scene s;
    if (!build_scene(scene, 1, s)) return FLUX_ERR_INVALID;
    Ray r = primary_ray(s, row, col, spx, spy, ldx, ldy);
    st3(o, r.origin);
    st3(d, r.direction);
    return FLUX_OK;
}

// Image::write, fluxcore/src/image.rs:42-60: (c * 65535.99) as u16 is a
// saturating, truncating cast (NaN -> 0) since Rust 1.45.
void oracle_ppm_quantize(const double *rgb, uint64_t n, uint16_t *out) {
    for (uint64_t i = 0; i < n; i++) {
        double v = rgb[i] * 65535.99;
        uint16_t q;
        if (!(v == v)) q = 0;
        else if (v <= 0.0) q = 0;
        else if (v >= 65535.0) q = 65535;
        else q = (uint16_t)v;
        out[i] = q;
    }
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
