// flux_intersect.cuh — closest hit over spheres / planes / triangles (linear scan).
//
// Device restatement of Scene::hit (fluxcore/src/scene.rs:156-160) with
// Hit::compare (common.rs:17-23), Sphere::hit + BoundingBox::hit
// (shapes.rs:98-133,171-217) and Plane::hit (shapes.rs:135-152).  Ray-invariant
// terms (1/d, sign tests, d.d, 2a, 4a) are hoisted out of the shape loop: they
// are the same IEEE values the reference recomputes per shape.  The full hit
// record (normal, point) is built once for the closest hit only.
#pragma once
#include "flux_scene.cuh"

struct RayCtx {
    V3 o, d;
    double ia, ib, ic;    // 1/dx, 1/dy, 1/dz       shapes.rs:107,114,121
    bool pa, pb, pc;      // ia >= 0.0 ...          shapes.rs:108,115,122
    double A;             // d.d                    shapes.rs:177
    double A2, A4;        // 2.0*a, 4.0*a           shapes.rs:187,180
};

__device__ __forceinline__ RayCtx make_ray(V3 o, V3 d) {
    RayCtx r;
    r.o = o;
    r.d = d;
    r.ia = 1.0 / d.x;
    r.ib = 1.0 / d.y;
    r.ic = 1.0 / d.z;
    r.pa = r.ia >= 0.0;
    r.pb = r.ib >= 0.0;
    r.pc = r.ic >= 0.0;
    r.A = dot3(d, d);
    r.A2 = 2.0 * r.A;
    r.A4 = 4.0 * r.A;
    return r;
}

struct HitRef {
    double t;
    uint32_t shape_id;   // 0xFFFFFFFF = none
    uint32_t kind;       // KIND_*
    uint32_t index;      // index inside the per-kind arrays
};

// candidate c replaces best iff !(best.t <= c.t) when c comes later in shape
// order (common.rs:17-23 + Iterator::min_by); shapes of different kinds are
// scanned kind by kind here, so order is restored through shape_id.
__device__ __forceinline__ void consider(HitRef &best, double t, uint32_t shape_id, uint32_t kind, uint32_t index) {
    bool take = (best.shape_id == 0xFFFFFFFFu) || (t < best.t) || (t == best.t && shape_id < best.shape_id);
    if (take) {
        best.t = t;
        best.shape_id = shape_id;
        best.kind = kind;
        best.index = index;
    }
}

// BoundingBox::hit, shapes.rs:98-133, on hoisted reciprocals.
__device__ __forceinline__ bool bbox_hit(const RayCtx &r, double c0x, double c0y, double c0z, double c1x,
                                         double c1y, double c1z) {
    double tx_min = ((r.pa ? c0x : c1x) - r.o.x) * r.ia;
    double tx_max = ((r.pa ? c1x : c0x) - r.o.x) * r.ia;
    double ty_min = ((r.pb ? c0y : c1y) - r.o.y) * r.ib;
    double ty_max = ((r.pb ? c1y : c0y) - r.o.y) * r.ib;
    double tz_min = ((r.pc ? c0z : c1z) - r.o.z) * r.ic;
    double tz_max = ((r.pc ? c1z : c0z) - r.o.z) * r.ic;
    double t0 = ref_max(tx_min, ref_max(ty_min, tz_min));
    double t1 = ref_min(tx_max, ref_min(ty_max, tz_max));
    return t0 < t1 && t1 > FLUX_T_MIN;
}

// Sphere::hit distance, shapes.rs:171-217 (returns false for None).
template <bool COUNT>
__device__ __forceinline__ bool sphere_t(const RayCtx &r, const double *__restrict__ sph, uint32_t n, uint32_t i,
                                         double &t_out, unsigned long long *cn) {
    if (COUNT) cn[CN_BBOX_TESTS]++;
    if (!bbox_hit(r, __ldg(sph + SPH_C0X * n + i), __ldg(sph + SPH_C0Y * n + i), __ldg(sph + SPH_C0Z * n + i),
                  __ldg(sph + SPH_C1X * n + i), __ldg(sph + SPH_C1Y * n + i), __ldg(sph + SPH_C1Z * n + i)))
        return false;
    if (COUNT) cn[CN_BBOX_PASS]++;
    V3 c = mk3(__ldg(sph + SPH_CX * n + i), __ldg(sph + SPH_CY * n + i), __ldg(sph + SPH_CZ * n + i));
    V3 temp = r.o - c;
    double b = 2.0 * dot3(temp, r.d);
    double cc = dot3(temp, temp) - __ldg(sph + SPH_RR * n + i);
    double disc = b * b - r.A4 * cc;
    if (disc < 0.0) return false;
    if (COUNT) cn[CN_DISC_NONNEG]++;
    double e = sqrt(disc);
    double t = (-b - e) / r.A2;
    if (!(t > FLUX_T_MIN)) {
        if (COUNT) cn[CN_T2]++;
        t = (-b + e) / r.A2;
        if (!(t > FLUX_T_MIN)) return false;
    }
    t_out = t;
    return true;
}

// Plane::hit distance, shapes.rs:135-152.
__device__ __forceinline__ bool plane_t(const RayCtx &r, const double *__restrict__ pln, uint32_t n, uint32_t i,
                                        double &t_out) {
    V3 p = mk3(__ldg(pln + PLN_PX * n + i), __ldg(pln + PLN_PY * n + i), __ldg(pln + PLN_PZ * n + i));
    V3 nn = mk3(__ldg(pln + PLN_NX * n + i), __ldg(pln + PLN_NY * n + i), __ldg(pln + PLN_NZ * n + i));
    double t = dot3(p - r.o, nn) / dot3(r.d, nn);
    t_out = t;
    return t > FLUX_T_MIN;
}

// EXTENSION: two-sided Moller-Trumbore, same op order as oracle tri_hit.
__device__ __forceinline__ bool tri_t(const RayCtx &r, const double *__restrict__ tri, uint32_t n, uint32_t i,
                                      double &t_out) {
    V3 v0 = mk3(__ldg(tri + TRI_V0X * n + i), __ldg(tri + TRI_V0Y * n + i), __ldg(tri + TRI_V0Z * n + i));
    V3 e1 = mk3(__ldg(tri + TRI_E1X * n + i), __ldg(tri + TRI_E1Y * n + i), __ldg(tri + TRI_E1Z * n + i));
    V3 e2 = mk3(__ldg(tri + TRI_E2X * n + i), __ldg(tri + TRI_E2Y * n + i), __ldg(tri + TRI_E2Z * n + i));
    V3 p = cross3(r.d, e2);
    double det = dot3(e1, p);
    if (det == 0.0) return false;
    double inv = 1.0 / det;
    V3 s = r.o - v0;
    double u = dot3(s, p) * inv;
    if (!(u >= 0.0 && u <= 1.0)) return false;
    V3 q = cross3(s, e1);
    double v = dot3(r.d, q) * inv;
    if (!(v >= 0.0 && u + v <= 1.0)) return false;
    double t = dot3(e2, q) * inv;
    if (!(t > FLUX_T_MIN)) return false;
    t_out = t;
    return true;
}

// Scene::hit by linear scan over all shapes.
template <bool COUNT>
__device__ __forceinline__ HitRef closest_hit_linear(const DevScene &sc, const RayCtx &r, unsigned long long *cn) {
    HitRef best;
    best.t = 0.0;
    best.shape_id = 0xFFFFFFFFu;
    best.kind = 0;
    best.index = 0;
    for (uint32_t i = 0; i < sc.n_spheres; i++) {
        double t;
        if (sphere_t<COUNT>(r, sc.sph, sc.n_spheres, i, t, cn)) {
            if (COUNT) cn[CN_CANDIDATES]++;
            consider(best, t, __ldg(sc.sph_meta + i), KIND_SPHERE, i);
        }
    }
    for (uint32_t i = 0; i < sc.n_planes; i++) {
        double t;
        if (COUNT) cn[CN_PLANE_TESTS]++;
        if (plane_t(r, sc.pln, sc.n_planes, i, t)) {
            if (COUNT) cn[CN_CANDIDATES]++;
            consider(best, t, __ldg(sc.pln_meta + i), KIND_PLANE, i);
        }
    }
    for (uint32_t i = 0; i < sc.n_tris; i++) {
        double t;
        if (COUNT) cn[CN_TRI_TESTS]++;
        if (tri_t(r, sc.tri, sc.n_tris, i, t)) {
            if (COUNT) cn[CN_CANDIDATES]++;
            consider(best, t, __ldg(sc.tri_meta + i), KIND_TRI, i);
        }
    }
    return best;
}

// Hit record of the closest hit: normal, local_hit_point, material
// (shapes.rs:140-147, 191-198).
struct HitRec {
    V3 normal, point;
    uint32_t material;
};

__device__ __forceinline__ HitRec build_hit(const DevScene &sc, const RayCtx &r, const HitRef &h) {
    HitRec out;
    out.point = r.o + h.t * r.d;
    if (h.kind == KIND_SPHERE) {
        uint32_t n = sc.n_spheres, i = h.index;
        V3 c = mk3(__ldg(sc.sph + SPH_CX * n + i), __ldg(sc.sph + SPH_CY * n + i), __ldg(sc.sph + SPH_CZ * n + i));
        V3 temp = r.o - c;
        out.normal = ((temp + h.t * r.d) * __ldg(sc.sph + SPH_INV * n + i)) / __ldg(sc.sph + SPH_R * n + i);
        out.material = __ldg(sc.sph_meta + n + i);
    } else if (h.kind == KIND_PLANE) {
        uint32_t n = sc.n_planes, i = h.index;
        out.normal = mk3(__ldg(sc.pln + PLN_NX * n + i), __ldg(sc.pln + PLN_NY * n + i), __ldg(sc.pln + PLN_NZ * n + i));
        out.material = __ldg(sc.pln_meta + n + i);
    } else {
        uint32_t n = sc.n_tris, i = h.index;
        V3 e1 = mk3(__ldg(sc.tri + TRI_E1X * n + i), __ldg(sc.tri + TRI_E1Y * n + i), __ldg(sc.tri + TRI_E1Z * n + i));
        V3 e2 = mk3(__ldg(sc.tri + TRI_E2X * n + i), __ldg(sc.tri + TRI_E2Y * n + i), __ldg(sc.tri + TRI_E2Z * n + i));
        out.normal = normalize3(cross3(e1, e2));
        out.material = __ldg(sc.tri_meta + n + i);
    }
    return out;
}
