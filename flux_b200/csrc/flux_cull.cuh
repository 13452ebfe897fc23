// flux_cull.cuh — Scene::hit by linear scan (scene.rs:156-160) with the conservative FP32 box classification of
// render_wave2.cu, for the direct kernel (render.cu: low sample counts — BASELINE config 1 — and the progressive passes).
//
// Round 1's direct kernel ran BoundingBox::hit (shapes.rs:98-133) in f64 for every sphere of every segment: on the
// shipped scenes that is more than half of the path's FP64 operations (515 of 919 per sample on demo2),
// and ncu had the kernel at 0.11 – 0.13 of the FP64 roofline.  Here every sphere box is first classified in FP32 from
// the kernel-parameter constant bank (RenderParams::cull) exactly as in render_wave2.cu — certainly hit / certainly
// missed / too close to call, with the rigorous bound E derived there — and only the undecided boxes run the exact f64
// test.  The set of passing boxes, and with it hit ids, distances and event counters, is exactly the reference's.
// The quadratics of the passing spheres run in shape order on the SoA scene in global memory (L1-resident: a few
// spheres), both roots sharing one refined reciprocal of 2a (flux_math.cuh).
//
// (The classification code is a copy of render_wave2.cu's cull_boxes, not a shared function: the wavefront kernel sits
// on its register and instruction-cache limits and is not to be perturbed by a refactoring.)
#pragma once
#include "flux_intersect.cuh"

struct LinCullRay {
    float iax, iay, iaz, nox, noy, noz;   // 1/d and -(o * 1/d), rounded to f32
    float aax, aay, aaz;                  // |1/d|
    float e2;                             // 2E + slack, NaN when nothing may be decided in FP32
};

__device__ __forceinline__ LinCullRay lin_cull_ray(const RenderParams &p, V3 o, V3 d) {
    LinCullRay c;
    asm("rcp.approx.f32 %0, %1;" : "=f"(c.iax) : "f"((float)d.x));   // max relative error 2^-23: part of E
    asm("rcp.approx.f32 %0, %1;" : "=f"(c.iay) : "f"((float)d.y));
    asm("rcp.approx.f32 %0, %1;" : "=f"(c.iaz) : "f"((float)d.z));
    const float ofx = (float)o.x, ofy = (float)o.y, ofz = (float)o.z;
    c.nox = -(ofx * c.iax); c.noy = -(ofy * c.iay); c.noz = -(ofz * c.iaz);
    c.aax = fabsf(c.iax); c.aay = fabsf(c.iay); c.aaz = fabsf(c.iaz);
    const float ex = c.aax * (p.cull_cmax + fabsf(ofx));
    const float ey = c.aay * (p.cull_cmax + fabsf(ofy));
    const float ez = c.aaz * (p.cull_cmax + fabsf(ofz));
    const float E = fmaxf(fmaxf(ex, ey), ez) * (1.01f * 9.5367431640625e-07f);   // 1.01 * 2^-20 max_k |1/d_k| (cmax + |o_k|)
    // outside a sane range (NaN, inf, denormal reciprocals) nothing is decided in FP32
    c.e2 = (E > 1e-30f && E < 1e30f) ? 2.0f * E + 4e-9f : __int_as_float(0x7fc00000);
    return c;
}

// Closest hit over spheres (FP32-classified boxes, at most FLUX_CULL_MAX of them) and planes; no triangles.
template <bool COUNT>
__device__ __forceinline__ HitRef closest_hit_linear_culled(const RenderParams &p, V3 o, V3 d, unsigned long long *cn) {
    const DevScene &sc = p.scene;
    const uint32_t ns = sc.n_spheres;
    const LinCullRay c = lin_cull_ray(p, o, d);
    const double A = dot3(d, d);
    const double A4 = 4.0 * A;                 // shapes.rs:180
    const RcpD rA2 = rcp_prepare(2.0 * A);     // shapes.rs:187
    HitRef best;
    best.t = 0.0;
    best.shape_id = 0xFFFFFFFFu;
    best.kind = 0;
    best.index = 0;
#pragma unroll 1
    for (uint32_t base = 0; base < ns; base += 32u) {
        const uint32_t nsb = ns - base < 32u ? ns - base : 32u;
        uint32_t okm = 0u, failm = 0u;
#pragma unroll 1
        for (uint32_t j = 0; j < nsb; j++) {
            const uint32_t i = base + j;
            const float r = p.cull[i][3];
            const float tcx = fmaf(p.cull[i][0], c.iax, c.nox);
            const float tcy = fmaf(p.cull[i][1], c.iay, c.noy);
            const float tcz = fmaf(p.cull[i][2], c.iaz, c.noz);
            const float tn = fmaxf(fmaxf(fmaf(-r, c.aax, tcx), fmaf(-r, c.aay, tcy)), fmaxf(fmaf(-r, c.aaz, tcz), (float)FLUX_T_MIN));
            const float tf = fminf(fminf(fmaf(r, c.aax, tcx), fmaf(r, c.aay, tcy)), fmaf(r, c.aaz, tcz));
            const float sgap = tf - tn;
            if (sgap > c.e2) okm |= 1u << j;
            if (sgap < -c.e2) failm |= 1u << j;
        }
        const uint32_t valid = nsb >= 32u ? ~0u : ((1u << nsb) - 1u);
        uint32_t mask = okm & valid;
        uint32_t unc = ~(okm | failm) & valid;
        if (unc) {   // rays grazing a box face, degenerate rays, invalid spheres: the exact test (shapes.rs:98-133)
            const RayCtx r = make_ray(o, d);
            while (unc) {
                const uint32_t j = (uint32_t)__ffs((int)unc) - 1u;
                unc &= unc - 1u;
                const uint32_t i = base + j;
                if (bbox_hit(r, __ldg(sc.sph + SPH_C0X * ns + i), __ldg(sc.sph + SPH_C0Y * ns + i), __ldg(sc.sph + SPH_C0Z * ns + i),
                             __ldg(sc.sph + SPH_C1X * ns + i), __ldg(sc.sph + SPH_C1Y * ns + i), __ldg(sc.sph + SPH_C1Z * ns + i)))
                    mask |= 1u << j;
            }
        }
        if (COUNT) {
            cn[CN_BBOX_TESTS] += nsb;
            cn[CN_BBOX_PASS] += __popc(mask);
        }
        while (mask) {   // quadratics of the passing spheres in shape order (shapes.rs:176-212)
            const uint32_t j = (uint32_t)__ffs((int)mask) - 1u;
            mask &= mask - 1u;
            const uint32_t i = base + j;
            const V3 temp = mk3(o.x - __ldg(sc.sph + SPH_CX * ns + i), o.y - __ldg(sc.sph + SPH_CY * ns + i), o.z - __ldg(sc.sph + SPH_CZ * ns + i));
            const double b = 2.0 * dot3(temp, d);
            const double cc = dot3(temp, temp) - __ldg(sc.sph + SPH_RR * ns + i);
            const double disc = b * b - A4 * cc;
            if (disc < 0.0) continue;
            if (COUNT) cn[CN_DISC_NONNEG]++;
            const double e = sqrt(disc);
            double t = div_by(-b - e, rA2);
            if (!(t > FLUX_T_MIN)) {
                if (COUNT) cn[CN_T2]++;
                t = div_by(-b + e, rA2);
                if (!(t > FLUX_T_MIN)) continue;
            }
            if (COUNT) cn[CN_CANDIDATES]++;
            // spheres arrive in shape order: a later one wins only if strictly closer (common.rs:17-23 + min_by)
            if (best.shape_id == 0xFFFFFFFFu || t < best.t) {
                best.t = t;
                best.shape_id = __ldg(sc.sph_meta + i);
                best.kind = KIND_SPHERE;
                best.index = i;
            }
        }
    }
    for (uint32_t i = 0; i < sc.n_planes; i++) {   // shapes.rs:137-139
        if (COUNT) cn[CN_PLANE_TESTS]++;
        const V3 pp = mk3(__ldg(sc.pln + PLN_PX * sc.n_planes + i), __ldg(sc.pln + PLN_PY * sc.n_planes + i), __ldg(sc.pln + PLN_PZ * sc.n_planes + i));
        const V3 pn = mk3(__ldg(sc.pln + PLN_NX * sc.n_planes + i), __ldg(sc.pln + PLN_NY * sc.n_planes + i), __ldg(sc.pln + PLN_NZ * sc.n_planes + i));
        const double t = dot3(pp - o, pn) / dot3(d, pn);
        if (!(t > FLUX_T_MIN)) continue;
        if (COUNT) cn[CN_CANDIDATES]++;
        consider(best, t, __ldg(sc.pln_meta + i), KIND_PLANE, i);
    }
    return best;
}
