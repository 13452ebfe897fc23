// render_wave2.cu — block-local wavefront render kernel, second generation.
//
// Same arithmetic as render.cu / render_regen.cu / render_wave.cu (Camera::render trace.rs:53-97 →
// Scene::shade scene.rs:162-172 → Material::path_shade materials.rs:19-71); it re-divides the work after
// what ncu showed on render_wave.cu (profiles/r1e_render_wave_full.txt: 169 warp instructions per sample,
// 29 % of them FP64, issue slots 62 % busy, barrier stalls on top):
//
//   * the slab test BoundingBox::hit (shapes.rs:98-133) was 26 % of all issue slots at 42 instructions per
//     test.  Here every sphere first gets a CONSERVATIVE FP32 slab test whose operands come from the
//     kernel-parameter constant bank (uniform loads, FP32 pipe): 9 FFMA + 3 min/max + 2 compares classify the
//     box as "certainly missed", "certainly hit" or "too close to call" with a rigorous error bound E; only
//     the last class (rays grazing a box face within ~1e-4, degenerate rays) runs the exact FP64 test.  The
//     pass/fail decisions — and therefore hit ids, distances and counters — are exactly the reference's;
//   * ray generation ran with 18 of 32 lanes (only the slots that had just died).  Terminated paths are now
//     one more "kind" of the sorted shading stage: one thread per ITEM unwinds the path's (f, w) stack,
//     adds the radiance to the slot's running sum and generates the slot's next camera ray, at full warps;
//   * the block-wide compaction of (slot, sphere) candidate pairs cost more issue slots (scan + publish +
//     fold + 3 barriers) than it saved in the quadratic; the slot owner now walks its own candidate bits.
//     One packed block scan and 3 barriers per iteration remain (was 3 scans and 8 barriers);
//   * divisions that share a divisor share one reciprocal refinement (flux_math.cuh rcp_prepare/div_by,
//     bit-identical to '/'): both roots by 2a, the sphere normal by r (refined once per CTA), normalize.
//
// A CTA owns one pixel and S = 256 path slots in shared memory.  Per iteration:
//   1 owner     (thread = slot) FP32 cull + exact tests of the uncertain boxes, quadratics of the passing
//               spheres in shape order (shapes.rs:176-212), planes (shapes.rs:137-139), closest hit
//               (scene.rs:156-160, common.rs:17-23), hit record, classification
//   2 bin       one packed prefix sum orders the slots by kind: matte | specular | glossy | terminated
//   3 item      (thread = item) Matte / PerfectSpecular / GlossySpecular shading (materials.rs:19-71,
//               brdf.rs:20-78) or unwind + accumulate + regenerate (materials.rs:31-32,69-70; trace.rs:72-82)
//
// Determinism: ranks come from prefix sums over slot order (no atomics); a slot's radiance sum is updated by
// exactly one thread per iteration, in iteration order; slot sums are combined by a fixed tree.  A pixel's
// value is therefore independent of grid size, scheduling and sharding, and equals
// render_wave.cu's bit for bit (same slot/sample assignment, same summation order).
#define FLUX_NORM_NOINLINE   // one copy of normalize3_dev per kernel: smaller hot instruction footprint (+1.4 % measured)
#include "flux_bvh.cuh"
#include "flux_kernels.cuh"
#include "flux_shade.cuh"

#ifndef WAVE2_S
#define WAVE2_S 256          // path slots per CTA == threads per CTA
#endif
#ifndef WAVE2_MIN_BLOCKS
#define WAVE2_MIN_BLOCKS (1024 / WAVE2_S)   // 64 registers per thread: 32 warps per SM
#endif
#define WAVE2_MAX_DEPTH 8    // deeper jobs use the other kernels
#ifndef WAVE2_OWNER_LIST
#define WAVE2_OWNER_LIST 1   // owner thread j traces the slot of item j of the iteration before: camera rays (the
                             // regenerated, "terminated" items, first in the item order) then sit in whole warps
#endif
#ifndef WAVE2_ORDER
#define WAVE2_ORDER 0         // item order after the terminated items: 0 specular | glossy | matte, 1 matte | specular | glossy
#endif
#ifndef WAVE2_BARRIER3
#define WAVE2_BARRIER3 1     // keep the barrier between the item stage and the next owner stage.  With the owner list it is not
                             // needed for correctness (see the loop end); measured in DESIGN.md §4.3
#endif
#ifndef WAVE2_PMASK
#define WAVE2_PMASK 1        // ... and those warps classify only the boxes a camera ray of the pixel can reach
#endif
#ifndef WAVE2_MIN_SPP
#define WAVE2_MIN_SPP 256    // measured r1 (demo2): 4.63 / 5.18 / 5.99 / 6.44 / 6.96 Gsamples/s at 256 / 529 / 1024 / 2025 /
                             // 16384 spp against 4.5-4.7 for the regeneration kernel; below 256 the per-pixel drain tail wins
#endif

namespace {

// sphere record in shared memory (doubles)
enum { V_C0X = 0, V_C1X, V_C0Y, V_C1Y, V_C0Z, V_C1Z, V_CX, V_CY, V_CZ, V_RR, V_RB, V_RY, V_INV, V_OK, V_SPH_STRIDE };
enum { V_PPX = 0, V_PPY, V_PPZ, V_PNX, V_PNY, V_PNZ, V_PLN_STRIDE };
// slot states
enum { ST_IDLE = 0,    // no path and no sample left
       ST_FRESH = 1,   // needs a camera ray, nothing to accumulate (start of a pixel)
       ST_ALIVE = 2 }; // carries a ray
// what the item stage does with a slot
enum { K_NONE = 0, K_MATTE = 1, K_SPEC = 2, K_GLOSSY = 3, K_TERM = 4 };
// how a path ended (payload of K_TERM)
enum { T_FRESH = 0, T_BLACK = 1, T_BACKGROUND = 2, T_EMIT = 3 };

// meta word: depth (0-7) | top (8-15) | material (16-27) | state (28-29) | term (30-31)
__device__ __forceinline__ uint32_t meta_pack(uint32_t depth, uint32_t top, uint32_t mat, uint32_t st, uint32_t term) {
    return depth | (top << 8) | (mat << 16) | (st << 28) | (term << 30);
}
__device__ __forceinline__ uint32_t meta_depth(uint32_t m) { return m & 0xFFu; }
__device__ __forceinline__ uint32_t meta_top(uint32_t m) { return (m >> 8) & 0xFFu; }
__device__ __forceinline__ uint32_t meta_mat(uint32_t m) { return (m >> 16) & 0xFFFu; }
__device__ __forceinline__ uint32_t meta_state(uint32_t m) { return (m >> 28) & 3u; }
__device__ __forceinline__ uint32_t meta_term(uint32_t m) { return (m >> 30) & 3u; }

struct W2Smem {
    double *sph, *pln;
    DevMaterial *mat;
    uint32_t *sph_id, *sph_mat, *pln_id, *pln_mat;
    double *ox, *oy, *oz, *dx, *dy, *dz, *nx, *ny, *nz;
    double *acc_r, *acc_g, *acc_b;
    double *stk_w, *stk_lobe;   // [depth][slot]
    uint8_t *stk_mat;           // [depth][slot]
    uint32_t *si, *meta, *list;
    unsigned long long *scr;    // [2][S/32] warp totals (double-buffered)
};

__host__ __device__ inline size_t w2_smem_bytes(uint32_t ns, uint32_t np, uint32_t nm, uint32_t max_depth) {
    size_t b = 0;
    b += (size_t)ns * V_SPH_STRIDE * 8 + (size_t)np * V_PLN_STRIDE * 8;
    b += (size_t)nm * sizeof(DevMaterial);
    b += (size_t)(2 * ns + 2 * np) * 4;
    b = (b + 15) / 16 * 16;
    b += (size_t)12 * WAVE2_S * 8;                     // ray, normal, radiance sums
    b += (size_t)2 * max_depth * WAVE2_S * 8;          // stack weight, lobe
    b += 2 * (WAVE2_S / 32) * 8;                       // scan scratch
    b += (size_t)3 * WAVE2_S * 4;                      // si, meta, list
    b += (size_t)max_depth * WAVE2_S;                  // stack material
    return (b + 15) / 16 * 16;
}

__device__ __forceinline__ W2Smem w2_carve(unsigned char *raw, uint32_t ns, uint32_t np, uint32_t nm, uint32_t max_depth) {
    W2Smem w;
    double *d = reinterpret_cast<double *>(raw);
    w.sph = d; d += (size_t)ns * V_SPH_STRIDE;
    w.pln = d; d += (size_t)np * V_PLN_STRIDE;
    w.mat = reinterpret_cast<DevMaterial *>(d);
    uint32_t *u = reinterpret_cast<uint32_t *>(w.mat + nm);
    w.sph_id = u; u += ns;
    w.sph_mat = u; u += ns;
    w.pln_id = u; u += np;
    w.pln_mat = u; u += np;
    size_t off = (size_t)(reinterpret_cast<unsigned char *>(u) - raw);
    off = (off + 15) / 16 * 16;
    d = reinterpret_cast<double *>(raw + off);
    w.ox = d; d += WAVE2_S; w.oy = d; d += WAVE2_S; w.oz = d; d += WAVE2_S;
    w.dx = d; d += WAVE2_S; w.dy = d; d += WAVE2_S; w.dz = d; d += WAVE2_S;
    w.nx = d; d += WAVE2_S; w.ny = d; d += WAVE2_S; w.nz = d; d += WAVE2_S;
    w.acc_r = d; d += WAVE2_S; w.acc_g = d; d += WAVE2_S; w.acc_b = d; d += WAVE2_S;
    w.stk_w = d; d += (size_t)max_depth * WAVE2_S;
    w.stk_lobe = d; d += (size_t)max_depth * WAVE2_S;
    w.scr = reinterpret_cast<unsigned long long *>(d); d += 2 * (WAVE2_S / 32);
    u = reinterpret_cast<uint32_t *>(d);
    w.si = u; u += WAVE2_S;
    w.meta = u; u += WAVE2_S;
    w.list = u; u += WAVE2_S;
    w.stk_mat = reinterpret_cast<uint8_t *>(u);
    return w;
}

// ---- conservative FP32 slab test against the constant-bank spheres ------------------------------------------
// A sphere's box is centre -+ r on every axis (Sphere::new, shapes.rs:156-161), so with tc_k = (m_k - o_k) / d_k the
// slab interval of axis k is tc_k -+ r |1/d_k|, whatever the sign of d_k.  In FP32, from f32-rounded m, r, 1/d and
// o/d:   tc_k = fma(m_k, ia_k, -(o_k ia_k));  near_k = fma(-r, |ia_k|, tc_k);  far_k = fma(r, |ia_k|, tc_k)
// with ia_k = rcp((float)d_k) (relative error <= 2^-22 including the rounding of d_k) and o_k ia_k one more f32
// product, each slab value differs from the reference's double (c - o) * (1/d) by at most
// (2^-22 + 2^-23)(|m| + |o| + r)|ia| + 2^-24 (|tc| + |near|)  <=  1.4 * 2^-21 (cmax + |o_k|) |ia_k|
// <= E = 1.01 * 2^-20 * max_k |ia_k| (cmax + |o_k|), cmax >= |m_k| + r.  With
//   s = min_k far_k - max(max_k near_k, T_MIN):
//   s >  2E (+ rounding slack)  =>  t0 < t1 and t1 > T_MIN in the reference: the box is hit
//   s < -2E (- rounding slack)  =>  t0 >= t1 or t1 < T_MIN in the reference: the box is missed
// anything else — including every NaN / inf, for which no comparison holds — is "uncertain" and takes the exact
// FP64 test.  Spheres with a negative or non-finite radius / centre carry NaN here and are always uncertain.
struct CullRay {
    float iax, iay, iaz, nox, noy, noz;   // 1/d and -(o * 1/d), rounded to f32
    float aax, aay, aaz;                  // |1/d|
    float e2;                             // 2E + slack
};

// MUFU.RCP: maximum relative error 2^-23 (PTX ISA, rcp.approx.f32); 1/0 = inf, denormal inputs behave as written
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Box classification of spheres [base, base + count): groups of 4 in a run-time loop, constant-bank operands through
// a uniform index.  (Fully unrolled with immediate constant offsets it was 8.6 % slower, r1: 150 more instructions
// in a hot loop that already fills the 32 KB L1.5 instruction cache; groups of 1 / 2 / 8: -1.3 % / 0 / -5 %.)
// Entries past the last sphere hold NaN (undecided) and are masked off by the caller.
#ifndef WAVE2_CULL_GROUP
#define WAVE2_CULL_GROUP 4
#endif
__device__ __forceinline__ void cull_boxes(const RenderParams &p, const CullRay &c, uint32_t base, uint32_t count,
                                           uint32_t &okm, uint32_t &failm) {
#pragma unroll 1
    for (uint32_t g = 0; g < count; g += WAVE2_CULL_GROUP) {
        uint32_t ok = 0, fail = 0;
#pragma unroll
        for (uint32_t j = 0; j < WAVE2_CULL_GROUP; j++) {
            const uint32_t i = base + g + j;
            const float r = p.cull[i][3];
            const float tcx = fmaf(p.cull[i][0], c.iax, c.nox);
            const float tcy = fmaf(p.cull[i][1], c.iay, c.noy);
            const float tcz = fmaf(p.cull[i][2], c.iaz, c.noz);
            const float tn = fmaxf(fmaxf(fmaf(-r, c.aax, tcx), fmaf(-r, c.aay, tcy)), fmaxf(fmaf(-r, c.aaz, tcz), (float)FLUX_T_MIN));
            const float tf = fminf(fminf(fmaf(r, c.aax, tcx), fmaf(r, c.aay, tcy)), fmaf(r, c.aaz, tcz));
            const float sgap = tf - tn;
            if (sgap > c.e2) ok |= 1u << j;
            if (sgap < -c.e2) fail |= 1u << j;
        }
        okm |= ok << g;
        failm |= fail << g;
    }
}

// The same classification for the spheres named by the bits of `m` only (uniform across the warp): the per-pixel
// primary-ray mask, see primary_mask() below.  One sphere per trip.
__device__ __forceinline__ void cull_boxes_masked(const RenderParams &p, const CullRay &c, uint32_t base, uint32_t m,
                                                  uint32_t &okm, uint32_t &failm) {
#pragma unroll 1
    while (m) {
        const uint32_t j = (uint32_t)__ffs((int)m) - 1u;
        m &= m - 1u;
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
}

// Per-pixel primary-ray mask.  Every camera ray of a pixel starts on the lens disc (radius R around the eye in the
// U,V plane, trace.rs:74-79) and passes through the pixel's rectangle on the focal plane (trace.rs:44-51, 72-73): in
// camera coordinates (U, V, -W) its point at axial distance s is  l (1 - s/f) + q (s/f)  with |l| <= R and q in that
// rectangle (centre qc, half diagonal rF), for s >= 0.  A sphere's box lies within rho = sqrt(3) |r| of its centre
// (cu, cv, cs).  With  D(s) = |(cu, cv) - qc s/f| - (R |1 - s/f| + rF s/f)  — Lipschitz in s with constant
// L = (|qc| + R + rF)/f — no ray of the pixel comes within rho of the centre at any s in [cs - rho, cs + rho] if
// D(max(cs, 0)) - L rho > rho, or if cs + rho < 0 (behind the lens).  Such a box fails BoundingBox::hit
// (shapes.rs:98-133) for every primary ray of the pixel, by a margin (1e-6 relative, 1e-9 absolute) that dwarfs the
// f64 rounding of the reference's slab arithmetic; it is left out of the mask.  Anything non-finite compares false and
// stays in.  Warps whose rays are all primary then classify 2-3 boxes instead of every sphere of the scene.
__device__ __forceinline__ bool primary_may_hit(const DevCamera &cam, double colf, double rowf, double cx, double cy, double cz,
                                                double r) {
    const double f = cam.focal;
    if (!(f > 1e-9)) return true;
    const double k = cam.aps * cam.factor;
    const double q0u = k * colf, q1u = k * (colf + 1.0), q0v = k * rowf, q1v = k * (rowf + 1.0);
    const double qcu = 0.5 * (q0u + q1u), qcv = 0.5 * (q0v + q1v);
    const double rF = 0.5 * sqrt((q1u - q0u) * (q1u - q0u) + (q1v - q0v) * (q1v - q0v)) * (1.0 + 1e-9);
    const double R = fabs(cam.lens_radius) * (1.0 + 1e-9);
    const V3 rel = mk3(cx - cam.eye.x, cy - cam.eye.y, cz - cam.eye.z);
    const double cu = dot3(rel, cam.u), cv = dot3(rel, cam.v), cs = -dot3(rel, cam.w);
    const double rho = 1.7320508075688774 * fabs(r) * (1.0 + 1e-9);
    if (cs + rho < 0.0) return false;
    const double s0 = cs > 0.0 ? cs : 0.0;
    const double a = s0 / f;
    const double du = cu - qcu * a, dv = cv - qcv * a;
    const double D = sqrt(du * du + dv * dv) - (R * fabs(1.0 - a) + rF * a);
    const double L = (sqrt(qcu * qcu + qcv * qcv) + R + rF) / f;
    const double scale = fabs(cu) + fabs(cv) + fabs(cs) + rho + R + 1.0;
    const bool excluded = D - L * rho > rho * (1.0 + 1e-6) + 1e-9 * scale;
    return !excluded;
}

// (A/B'd and rejected, r1: prefetch.global.L1 of the hemisphere / lobe sample at classification time and of the next
// window of camera samples — 6 % slower at both 4096 and 16384 spp; the loads' latency is already covered by the
// other resident CTAs.)

struct SphereScan {
    double A4;      // 4.0 * a, shapes.rs:180
    RcpD rA2;       // 2.0 * a, shapes.rs:187, with its refined reciprocal
    double best_t;
    uint32_t best_ref;
};

// Spheres [base, base + 32): FP32 classification of every box, exact BoundingBox::hit (shapes.rs:98-133) for the
// boxes FP32 could not decide, then the quadratics of the passing spheres in shape order (shapes.rs:176-212).
template <bool COUNT>
__device__ __forceinline__ void sphere_pass(const RenderParams &p, const double *sph, const CullRay &c, uint32_t base, uint32_t ns,
                                            V3 o, V3 d, SphereScan &sc, unsigned long long *cn, bool use_pm, uint32_t pm) {
    const uint32_t nsb = ns - base < 32u ? ns - base : 32u;   // spheres in this pass
    uint32_t okm = 0u, failm = 0u;
    uint32_t valid = nsb >= 32u ? ~0u : ((1u << nsb) - 1u);
    if (use_pm) {          // uniform: every tracing lane of the warp carries a camera ray of this pixel
        valid &= pm;
        cull_boxes_masked(p, c, base, valid, okm, failm);
    } else {
        cull_boxes(p, c, base, nsb, okm, failm);
    }
    uint32_t mask = okm & valid;
    uint32_t unc = ~(okm | failm) & valid;
    while (unc) {
        // the exact reciprocals (shapes.rs:107,114,121) are needed on this rare path only
        const double ia = 1.0 / d.x, ib = 1.0 / d.y, ic = 1.0 / d.z;
        const int sx = ia >= 0.0 ? 0 : 1, sy = ib >= 0.0 ? 0 : 1, sz = ic >= 0.0 ? 0 : 1;
        const uint32_t j = (uint32_t)__ffs((int)unc) - 1u;
        unc &= unc - 1u;
        const double *s = sph + (size_t)(base + j) * V_SPH_STRIDE;
        const double tx_min = (s[V_C0X + sx] - o.x) * ia, tx_max = (s[V_C1X - sx] - o.x) * ia;
        const double ty_min = (s[V_C0Y + sy] - o.y) * ib, ty_max = (s[V_C1Y - sy] - o.y) * ib;
        const double tz_min = (s[V_C0Z + sz] - o.z) * ic, tz_max = (s[V_C1Z - sz] - o.z) * ic;
        const double t0 = ref_max(tx_min, ref_max(ty_min, tz_min));
        const double t1 = ref_min(tx_max, ref_min(ty_max, tz_max));
        if (t0 < t1 && t1 > FLUX_T_MIN) mask |= 1u << j;
    }
    if (COUNT) {
        cn[CN_BBOX_TESTS] += nsb;
        cn[CN_BBOX_PASS] += __popc(mask);
    }
    while (mask) {
        const uint32_t j = (uint32_t)__ffs((int)mask) - 1u;
        mask &= mask - 1u;
        const double *s = sph + (size_t)(base + j) * V_SPH_STRIDE;
        const V3 temp = mk3(o.x - s[V_CX], o.y - s[V_CY], o.z - s[V_CZ]);
        const double b = 2.0 * dot3(temp, d);
        const double cc = dot3(temp, temp) - s[V_RR];
        const double disc = b * b - sc.A4 * cc;
        if (disc < 0.0) continue;
        if (COUNT) cn[CN_DISC_NONNEG]++;
        const double e = sqrt(disc);
        double t = div_by(-b - e, sc.rA2);
        if (!(t > FLUX_T_MIN)) {
            if (COUNT) cn[CN_T2]++;
            t = div_by(-b + e, sc.rA2);
            if (!(t > FLUX_T_MIN)) continue;
        }
        if (COUNT) cn[CN_CANDIDATES]++;
        // spheres arrive in shape order: a later one wins only if strictly closer (common.rs:17-23 + min_by)
        if (sc.best_ref == 0xFFFFFFFFu || t < sc.best_t) {
            sc.best_t = t;
            sc.best_ref = base + j;
        }
    }
}

// -DWAVE2_TIMING: per-warp cycle accounting of the stages and barrier waits (lane 0 of every warp, clock64), summed
// into the event-counter array and read by tools/wave2_timing.py: [0] owner stage, [1] wait at barrier 1, [2] bin,
// [3] wait at barrier 2, [4] item stage, [5] wait at barrier 3, [6] warp-iterations, [7..14] item-stage time and
// count by the kind of the warp's first lane (matte, glossy/specular, terminated, idle).  A debugging build only.
#ifdef WAVE2_TIMING
#define TMARK(k) { const long long t_now = clock64(); if (lane == 0) t_acc[k] += t_now - t_prev; t_prev = t_now; }
#else
#define TMARK(k)
#endif

// BVH = true: the owner stage finds the closest hit through the 4-wide BVH over the scene in global memory (meshes,
// more than 40 bounded shapes; flux_bvh.cuh, traversal stack in local memory) instead of classifying the spheres
// staged in shared memory; everything else — binning, sorted shading, regeneration — is the same code.
template <bool COUNT, bool BVH>
__global__ void __launch_bounds__(WAVE2_S, WAVE2_MIN_BLOCKS) render_wave2_kernel(const __grid_constant__ RenderParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t ns = BVH ? 0u : p.scene.n_spheres, np = BVH ? 0u : p.scene.n_planes, nm = p.scene.n_materials;
    const uint32_t max_depth = p.cam.max_depth;
    const W2Smem w = w2_carve(smem_raw, ns, np, nm, max_depth);
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    // ---- stage the scene (once per CTA) ----
    for (uint32_t k = tid; k < ns; k += WAVE2_S) {
        const double *g = p.scene.sph;
        double *r = w.sph + (size_t)k * V_SPH_STRIDE;
        r[V_C0X] = g[SPH_C0X * ns + k]; r[V_C1X] = g[SPH_C1X * ns + k];
        r[V_C0Y] = g[SPH_C0Y * ns + k]; r[V_C1Y] = g[SPH_C1Y * ns + k];
        r[V_C0Z] = g[SPH_C0Z * ns + k]; r[V_C1Z] = g[SPH_C1Z * ns + k];
        r[V_CX] = g[SPH_CX * ns + k]; r[V_CY] = g[SPH_CY * ns + k]; r[V_CZ] = g[SPH_CZ * ns + k];
        r[V_RR] = g[SPH_RR * ns + k];
        const RcpD rr = rcp_prepare(g[SPH_R * ns + k]);   // the normal divides by r (shapes.rs:195): refine 1/r once
        r[V_RB] = rr.b; r[V_RY] = rr.y; r[V_OK] = rr.ok ? 1.0 : 0.0;
        r[V_INV] = g[SPH_INV * ns + k];
        w.sph_id[k] = p.scene.sph_meta[k];
        w.sph_mat[k] = p.scene.sph_meta[ns + k];
    }
    for (uint32_t k = tid; k < np; k += WAVE2_S) {
        double *r = w.pln + (size_t)k * V_PLN_STRIDE;
        for (int f = 0; f < PLN_FIELDS; f++) r[f] = p.scene.pln[(size_t)f * np + k];
        w.pln_id[k] = p.scene.pln_meta[k];
        w.pln_mat[k] = p.scene.pln_meta[np + k];
    }
    for (uint32_t k = tid; k < nm; k += WAVE2_S) w.mat[k] = p.scene.materials[k];
    __syncthreads();

    const DevCamera &cam = p.cam;
    const uint32_t W = cam.W;
    const uint32_t npix = p.n_rows * W;
    const uint32_t n = p.ss.n;
    unsigned long long cn[COUNT ? CN_COUNT : 1];
    if (COUNT)
        for (int k = 0; k < CN_COUNT; k++) cn[k] = 0;
    const double pixel_denom = 1.0 / (double)((unsigned long long)p.ss.root * p.ss.root);  // trace.rs:59
    const float cull_scale = 1.01f * 9.5367431640625e-07f;   // 1.01 * 2^-20
    __shared__ uint32_t s_pixel;
    __shared__ uint32_t s_pm[FLUX_CULL_MAX / 32];   // per-pixel primary-ray mask (primary_may_hit)
    __shared__ double s_red[3][WAVE2_S / 32];
    uint32_t rot = 0;
#ifdef WAVE2_TIMING
    long long t_acc[15] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#endif

    for (;;) {
        if (tid == 0) s_pixel = atomicAdd(p.work_counter, 1u);
        __syncthreads();
        const uint32_t pixel = s_pixel;
        if (pixel >= npix) break;
        const uint32_t rk = pixel / W;
        const uint32_t col = pixel - rk * W;
        const uint32_t row = p.rows[rk];
        const uint32_t set = p.set_index[(size_t)row * W + col];
        const double2 *ps = p.ss.pixel + (size_t)set * n;
        const double2 *ds = p.ss.disc + (size_t)set * n;
        const double *hs = p.ss.hemi + (size_t)set * p.ss.max_depth * n * 3;
        const double colf = (double)col - cam.half_w;           // trace.rs:72
        const double rowf = (double)(cam.H - row) - cam.half_h; // trace.rs:73

        uint32_t next = 0;   // next unassigned sample index (identical in all threads)
        w.meta[tid] = meta_pack(0, 0, 0, ST_FRESH, T_FRESH);
        w.acc_r[tid] = 0.0; w.acc_g[tid] = 0.0; w.acc_b[tid] = 0.0;
        uint32_t n_prev = WAVE2_S;   // items of the iteration before = owner threads of this one
        if (WAVE2_OWNER_LIST) w.list[tid] = tid;
        if (WAVE2_PMASK && !BVH && !COUNT) {
            bool may = false;
            if (tid < ns) {
                const double *s = w.sph + (size_t)tid * V_SPH_STRIDE;
                // caller-supplied sample sets outside the unit square / disc void the bound: then every box stays in the mask
                may = p.primary_mask_ok ? primary_may_hit(cam, colf, rowf, s[V_CX], s[V_CY], s[V_CZ], s[V_RB]) : true;
            }
            const uint32_t b = __ballot_sync(0xffffffffu, may);
            if (lane == 0 && warp < FLUX_CULL_MAX / 32) s_pm[warp] = b;
        }
        __syncthreads();

        for (;;) {
            // ======================= 1: owner — closest hit and classification =======================
#ifdef WAVE2_TIMING
            long long t_prev = clock64();
#endif
            // thread -> slot: with WAVE2_OWNER_LIST the slot this thread shaded or regenerated as item `tid` of the
            // iteration before (read again from the list, which is rewritten only after barrier 1: carrying it in a
            // register across the iteration cost 60 bytes of spill and 10 % — r2k); else slot `tid`
            const bool own = !WAVE2_OWNER_LIST || tid < n_prev;
            const uint32_t sl = WAVE2_OWNER_LIST ? (own ? w.list[tid] : 0u) : tid;
            uint32_t m = own ? w.meta[sl] : meta_pack(0, 0, 0, ST_IDLE, 0);
            uint32_t kind = K_NONE;
            bool use_pm = false;
            if (WAVE2_PMASK && !BVH && !COUNT) {
                const bool tracing = meta_state(m) == ST_ALIVE && meta_depth(m) <= max_depth;
                use_pm = __all_sync(0xffffffffu, !tracing || meta_depth(m) == 1u);
            }
            if (meta_state(m) == ST_FRESH) {
                kind = K_TERM;   // payload T_FRESH: just generate the first ray
            } else if (meta_state(m) == ST_ALIVE) {
                const uint32_t depth = meta_depth(m), top = meta_top(m);
                if (depth > max_depth) {   // scene.rs:164-165
                    if (COUNT) cn[CN_DEPTH_CUT]++;
                    kind = K_TERM;
                    w.meta[sl] = meta_pack(depth, top, 0, ST_ALIVE, T_BLACK);
                } else {
                    if (COUNT) cn[CN_SEGMENTS]++;
                    const V3 o = mk3(w.ox[sl], w.oy[sl], w.oz[sl]);
                    const V3 d = mk3(w.dx[sl], w.dy[sl], w.dz[sl]);
                    bool found;
                    V3 normal = mk3(0.0, 0.0, 0.0), point = mk3(0.0, 0.0, 0.0);
                    uint32_t mi = 0;
                    if (BVH) {
                        uint2 lstack[BVH_STACK];   // local memory: shared memory is taken by the path rows
                        const RayCtx rc = make_ray(o, d);
                        const HitRef h = closest_hit_bvh<COUNT>(p.scene, rc, lstack, 1u, cn);
                        found = h.shape_id != 0xFFFFFFFFu;
                        if (found) {
                            const HitRec hrec = build_hit(p.scene, rc, h);
                            normal = hrec.normal;
                            point = hrec.point;
                            mi = hrec.material;
                            if (COUNT) cn[h.kind == KIND_SPHERE ? CN_HIT_SPHERE : (h.kind == KIND_PLANE ? CN_HIT_PLANE : CN_HIT_TRI)]++;
                        }
                    } else {
                        // ---- spheres: conservative FP32 box classification, exact test where undecided, quadratics ----
                        // 1/d in f32 from MUFU.RCP (the approximation error is part of E); the exact f64 reciprocals of
                        // BoundingBox::hit are formed only where a box needs the exact test
                        CullRay c;
                        c.iax = rcp_approx((float)d.x); c.iay = rcp_approx((float)d.y); c.iaz = rcp_approx((float)d.z);
                        const float ofx = (float)o.x, ofy = (float)o.y, ofz = (float)o.z;
                        c.nox = -(ofx * c.iax); c.noy = -(ofy * c.iay); c.noz = -(ofz * c.iaz);
                        c.aax = fabsf(c.iax); c.aay = fabsf(c.iay); c.aaz = fabsf(c.iaz);
                        {
                            const float ex = c.aax * (p.cull_cmax + fabsf(ofx));
                            const float ey = c.aay * (p.cull_cmax + fabsf(ofy));
                            const float ez = c.aaz * (p.cull_cmax + fabsf(ofz));
                            const float E = fmaxf(fmaxf(ex, ey), ez) * cull_scale;
                            // outside a sane range (NaN, inf, denormal reciprocals) nothing is decided in FP32
                            c.e2 = (E > 1e-30f && E < 1e30f) ? 2.0f * E + 4e-9f : __int_as_float(0x7fc00000);
                        }
                        const double A = dot3(d, d);
                        SphereScan sc;
                        sc.A4 = 4.0 * A;
                        sc.rA2 = rcp_prepare(2.0 * A);
                        sc.best_t = 0.0;
                        sc.best_ref = 0xFFFFFFFFu;  // sphere index; plane = 0x80000000 | index; none = 0xFFFFFFFF
    #pragma unroll 1
                        for (uint32_t base = 0; base < ns; base += 32) sphere_pass<COUNT>(p, w.sph, c, base, ns, o, d, sc, cn, use_pm, use_pm ? s_pm[base >> 5] : 0u);
                        double best_t = sc.best_t;
                        uint32_t best_ref = sc.best_ref;
                        // ---- planes (shapes.rs:137-139) ----
                        uint32_t best_id = best_ref == 0xFFFFFFFFu ? 0xFFFFFFFFu : w.sph_id[best_ref];
                        const double *pl = w.pln;
                        for (uint32_t i = 0; i < np; i++, pl += V_PLN_STRIDE) {
                            if (COUNT) cn[CN_PLANE_TESTS]++;
                            const V3 pn = mk3(pl[V_PNX], pl[V_PNY], pl[V_PNZ]);
                            const double t = div_full(dot3(mk3(pl[V_PPX] - o.x, pl[V_PPY] - o.y, pl[V_PPZ] - o.z), pn), dot3(d, pn));
                            if (!(t > FLUX_T_MIN)) continue;
                            if (COUNT) cn[CN_CANDIDATES]++;
                            const uint32_t id = w.pln_id[i];
                            if (best_id == 0xFFFFFFFFu || t < best_t || (t == best_t && id < best_id)) {
                                best_t = t;
                                best_id = id;
                                best_ref = 0x80000000u | i;
                            }
                        }
                        found = best_id != 0xFFFFFFFFu;
                        if (found) {
                            // hit record of the closest hit only (shapes.rs:140-147,191-198)
                            point = o + best_t * d;
                            if (best_ref & 0x80000000u) {
                                const uint32_t k = best_ref & 0x7FFFFFFFu;
                                const double *q = w.pln + (size_t)k * V_PLN_STRIDE;
                                normal = mk3(q[V_PNX], q[V_PNY], q[V_PNZ]);
                                mi = w.pln_mat[k];
                                if (COUNT) cn[CN_HIT_PLANE]++;
                            } else {
                                const double *s = w.sph + (size_t)best_ref * V_SPH_STRIDE;
                                const V3 temp = mk3(o.x - s[V_CX], o.y - s[V_CY], o.z - s[V_CZ]);
                                const V3 nn = (temp + best_t * d) * s[V_INV];
                                RcpD rr;
                                rr.b = s[V_RB]; rr.y = s[V_RY]; rr.ok = s[V_OK] != 0.0;
                                normal = V3{div_by(nn.x, rr), div_by(nn.y, rr), div_by(nn.z, rr)};
                                mi = w.sph_mat[best_ref];
                                if (COUNT) cn[CN_HIT_SPHERE]++;
                            }
                        }
                    }
                    if (!found) {  // scene.rs:168
                        if (COUNT) cn[CN_MISS]++;
                        kind = K_TERM;
                        w.meta[sl] = meta_pack(depth, top, 0, ST_ALIVE, T_BACKGROUND);
                    } else {
                        const uint32_t mk = w.mat[mi].kind;
                        if (mk == FLUX_MAT_EMISSIVE) {  // materials.rs:42-49
                            if (COUNT) cn[CN_EMISSIVE]++;
                            kind = K_TERM;
                            w.meta[sl] = meta_pack(depth, top, mi, ST_ALIVE, dot3(normal * -1.0, d) > 0.0 ? T_EMIT : T_BLACK);
                        } else {
                            kind = mk == FLUX_MAT_MATTE ? K_MATTE : (mk == FLUX_MAT_REFLECTIVE ? K_SPEC : K_GLOSSY);
                            w.nx[sl] = normal.x; w.ny[sl] = normal.y; w.nz[sl] = normal.z;
                            w.ox[sl] = point.x; w.oy[sl] = point.y; w.oz[sl] = point.z;  // child ray origin
                            w.meta[sl] = meta_pack(depth, top, mi, ST_ALIVE, 0);
                        }
                    }
                }
            }

            // ======================= 2: bin the slots by kind (one packed scan) =======================
            // per-warp counts travel as two packed words (matte | specular << 16, glossy | terminated << 16); after the
            // barrier every warp sums the 8 warp totals, and those of the warps before it, with four independent
            // REDUX instructions (the item stage was measured to wait on this dependent chain: tools/wave2_timing.py).
            // Item order: terminated | specular | glossy | matte.  The terminated items regenerate camera rays, and with
            // WAVE2_OWNER_LIST item j's slot is traced by owner thread j of the next iteration: the camera rays start at
            // thread 0, in whole warps, which is what the per-pixel primary mask needs.  (r1 had the terminated items
            // between glossy and matte so that the two then most expensive kinds never shared a warp; since the lobe is
            // no longer evaluated the three kinds cost 4.1 - 4.8 K cycles per warp and the order is free.)
            uint32_t n_matte, n_spec, n_gloss, n_term, pos;
            {
                const uint32_t bm = __ballot_sync(0xffffffffu, kind == K_MATTE);
                const uint32_t bs = __ballot_sync(0xffffffffu, kind == K_SPEC);
                const uint32_t bg = __ballot_sync(0xffffffffu, kind == K_GLOSSY);
                const uint32_t bt = __ballot_sync(0xffffffffu, kind == K_TERM);
                uint2 *scr = reinterpret_cast<uint2 *>(w.scr) + (WAVE2_S / 32) * (rot++ & 1u);
                if (lane == 0) scr[warp] = make_uint2(__popc(bm) | (__popc(bs) << 16), __popc(bg) | (__popc(bt) << 16));
                const uint32_t mine = kind == K_MATTE ? bm : (kind == K_SPEC ? bs : (kind == K_GLOSSY ? bg : bt));
                const uint32_t rank = __popc(mine & lt_mask);
                TMARK(0);
                __syncthreads();
                TMARK(1);
                const uint2 v = lane < WAVE2_S / 32 ? scr[lane] : make_uint2(0u, 0u);
                const uint32_t tx = __reduce_add_sync(0xffffffffu, v.x), ty = __reduce_add_sync(0xffffffffu, v.y);
                const uint32_t bx = __reduce_add_sync(0xffffffffu, lane < warp ? v.x : 0u);
                const uint32_t by = __reduce_add_sync(0xffffffffu, lane < warp ? v.y : 0u);
                n_matte = tx & 0xFFFFu; n_spec = tx >> 16; n_gloss = ty & 0xFFFFu; n_term = ty >> 16;
#if WAVE2_ORDER == 0   // terminated | specular | glossy | matte
                pos = rank + (kind == K_TERM ? (by >> 16)
                              : kind == K_SPEC ? n_term + (bx >> 16)
                              : kind == K_GLOSSY ? n_term + n_spec + (by & 0xFFFFu)
                                                 : n_term + n_spec + n_gloss + (bx & 0xFFFFu));
#else                  // terminated | matte | specular | glossy
                pos = rank + (kind == K_TERM ? (by >> 16)
                              : kind == K_MATTE ? n_term + (bx & 0xFFFFu)
                              : kind == K_SPEC ? n_term + n_matte + (bx >> 16)
                                               : n_term + n_matte + n_spec + (by & 0xFFFFu));
#endif
            }
            const uint32_t n_items = n_matte + n_spec + n_gloss + n_term;
            if (n_items == 0) break;   // every slot idle (uniform)
            if (kind != K_NONE) w.list[pos] = sl;
            if (WAVE2_OWNER_LIST) n_prev = n_items;
            TMARK(2);
            __syncthreads();
            TMARK(3);

            // ======================= 3: one thread per item, in kind order =======================
            if (tid < n_items) {
                const uint32_t sl = w.list[tid];
                const uint32_t sm = w.meta[sl];
                const uint32_t depth = meta_depth(sm), top = meta_top(sm), mi = meta_mat(sm);
#if WAVE2_ORDER == 0
                const uint32_t a_gloss = n_term + n_spec, a_matte = a_gloss + n_gloss;   // block starts (specular: n_term)
                const bool is_matte = tid >= a_matte, is_spec = tid < a_gloss;
#else
                const uint32_t a_matte = n_term, a_spec = n_term + n_matte, a_gloss = a_spec + n_spec;
                const bool is_matte = tid < a_spec, is_spec = tid < a_gloss;
#endif
                if (tid >= n_term) {
                    const V3 normal = mk3(w.nx[sl], w.ny[sl], w.nz[sl]);
                    const V3 dir = mk3(w.dx[sl], w.dy[sl], w.dz[sl]);
                    const uint32_t i = w.si[sl];
                    V3 wi;
                    double weight, lobe = 1.0;
                    if (is_matte) {  // materials.rs:19-33
                        if (COUNT) cn[CN_MATTE]++;
                        const double *hp = hs + ((size_t)(depth - 1) * n + i) * 3;
                        matte_sample(normal, mk3(hp[0], hp[1], hp[2]), wi, weight);
                    } else if (is_spec) {  // specular: materials.rs:57-71, brdf.rs:39-45
                        if (COUNT) cn[CN_SPECULAR]++;
                        specular_sample(normal, dir, wi, weight);
                    } else {  // materials.rs:57-71, brdf.rs:55-78
                        if (COUNT) cn[CN_GLOSSY]++;
                        const DevMaterial &gm = w.mat[mi];
                        bool flipped;
                        if (p.ss.ghemi) {  // lobe table: to_unit_hemi(pixel sample, exp) precomputed (same bits)
                            const double *gh = p.ss.ghemi + (((size_t)set * p.ss.gk + gm.gidx) * n + i) * 3;
                            glossy_sample_hs(normal, dir, mk3(gh[0], gh[1], gh[2]), gm.exp, wi, weight, lobe, flipped);
                        } else {
                            const double2 s = ps[i];
                            glossy_sample(normal, dir, s.x, s.y, gm.exp, gm.inv_e1, wi, weight, lobe, flipped);
                        }
                        if (COUNT && flipped) cn[CN_GLOSSY_FLIP]++;
                    }
                    w.stk_w[(size_t)top * WAVE2_S + sl] = weight;
                    w.stk_lobe[(size_t)top * WAVE2_S + sl] = lobe;
                    w.stk_mat[(size_t)top * WAVE2_S + sl] = (uint8_t)mi;
                    w.dx[sl] = wi.x; w.dy[sl] = wi.y; w.dz[sl] = wi.z;
                    w.meta[sl] = meta_pack(depth + 1, top + 1, 0, ST_ALIVE, 0);
                } else {
                    // ---- a path ended: unwind, accumulate, start the slot's next path ----
                    const uint32_t term = meta_term(sm);
                    if (term != T_FRESH) {
                        double Lr = 0.0, Lg = 0.0, Lb = 0.0;
                        if (term == T_BACKGROUND) { Lr = cam.bg[0]; Lg = cam.bg[1]; Lb = cam.bg[2]; }
                        else if (term == T_EMIT) { const DevMaterial &em = w.mat[mi]; Lr = em.c[0]; Lg = em.c[1]; Lb = em.c[2]; }
                        uint32_t tp = top;
                        while (tp > 0) {  // (f (*) L) * w, innermost first: materials.rs:31-32,69-70
                            tp--;
                            const DevMaterial &smat = w.mat[w.stk_mat[(size_t)tp * WAVE2_S + sl]];
                            const double lobe = w.stk_lobe[(size_t)tp * WAVE2_S + sl];
                            const double wt = w.stk_w[(size_t)tp * WAVE2_S + sl];
                            Lr = ((smat.c[0] * lobe) * Lr) * wt;
                            Lg = ((smat.c[1] * lobe) * Lg) * wt;
                            Lb = ((smat.c[2] * lobe) * Lb) * wt;
                        }
                        w.acc_r[sl] += Lr;  // trace.rs:82 (this slot's paths, in completion order)
                        w.acc_g[sl] += Lg;
                        w.acc_b[sl] += Lb;
                    }
                    // terminated slots take the pixel's next sample indices in slot order
                    const uint32_t i = next + tid;   // the terminated items are the first of the list
                    if (i < n) {
                        const double2 s = ps[i];
                        const double2 l = ds[i];
                        // trace.rs:72-80 + Camera::ray_direction trace.rs:44-51
                        const double u = cam.aps * (colf + s.x);
                        const double v = cam.aps * (rowf + s.y);
                        const double lpx = l.x * cam.lens_radius;
                        const double lpy = l.y * cam.lens_radius;
                        const double px2 = u * cam.factor;
                        const double py2 = v * cam.factor;
                        const V3 d = normalize3_dev(((px2 - lpx) * cam.u + (py2 - lpy) * cam.v) - cam.focal_w);
                        const V3 o = (cam.eye + lpx * cam.u) + lpy * cam.v;
                        w.ox[sl] = o.x; w.oy[sl] = o.y; w.oz[sl] = o.z;
                        w.dx[sl] = d.x; w.dy[sl] = d.y; w.dz[sl] = d.z;
                        w.si[sl] = i;
                        w.meta[sl] = meta_pack(1, 0, 0, ST_ALIVE, 0);
                        if (COUNT) cn[CN_SAMPLES]++;
                    } else {
                        w.meta[sl] = meta_pack(0, 0, 0, ST_IDLE, 0);
                    }
                }
            }
            next += n_term;
#ifdef WAVE2_TIMING
            {
                const long long t_now = clock64();
                if (lane == 0) {
                    const uint32_t a1 = n_term, a2 = a1 + n_spec + n_gloss;
                    const int kk = tid >= n_items ? 3 : (tid >= a2 ? 0 : (tid < a1 ? 2 : 1));
                    t_acc[7 + 2 * kk] += t_now - t_prev;
                    t_acc[8 + 2 * kk] += 1;
                }
            }
#endif
            TMARK(4);
            // With WAVE2_OWNER_LIST the thread that handled a slot as an item traces it as its owner: between the item
            // stage and the next owner stage nothing crosses threads (rows [sl], the stack column and the radiance sum
            // of a slot are touched by its one thread; `list` is rewritten only after barrier 1, which no thread passes
            // before all have left this item stage), so this barrier could go — and was measured slower without
            // (WAVE2_BARRIER3 above).
            if (!WAVE2_OWNER_LIST || WAVE2_BARRIER3) __syncthreads();
            TMARK(5);
#ifdef WAVE2_TIMING
            if (lane == 0) t_acc[6] += 1;
#endif
        }

        // ---- fixed-shape reduction of the 256 slot sums, then trace.rs:85-86 + color.rs:35-44 ----
        double ar = w.acc_r[tid], ag = w.acc_g[tid], ab = w.acc_b[tid];
#pragma unroll
        for (uint32_t o2 = 16; o2 > 0; o2 >>= 1) {
            ar += __shfl_xor_sync(0xffffffffu, ar, o2);
            ag += __shfl_xor_sync(0xffffffffu, ag, o2);
            ab += __shfl_xor_sync(0xffffffffu, ab, o2);
        }
        if (lane == 0) {
            s_red[0][warp] = ar;
            s_red[1][warp] = ag;
            s_red[2][warp] = ab;
        }
        __syncthreads();
        if (tid == 0) {
            double tr[WAVE2_S / 32], tg[WAVE2_S / 32], tb[WAVE2_S / 32];
            for (int k = 0; k < WAVE2_S / 32; k++) { tr[k] = s_red[0][k]; tg[k] = s_red[1][k]; tb[k] = s_red[2][k]; }
            for (int st = 1; st < WAVE2_S / 32; st <<= 1)   // fixed pairwise tree over the 8 warp sums
                for (int k = 0; k + st < WAVE2_S / 32; k += 2 * st) { tr[k] += tr[k + st]; tg[k] += tg[k + st]; tb[k] += tb[k + st]; }
            double r = tr[0] * pixel_denom, g = tg[0] * pixel_denom, b = tb[0] * pixel_denom;
            const double mx1 = r > g ? r : g;
            const double mx2 = mx1 > b ? mx1 : b;
            if (mx2 > 1.0) {
                const double inv = 1.0 / mx2;
                r *= inv;
                g *= inv;
                b *= inv;
            }
            double *out = p.out + (p.out_by_row ? (size_t)row * W + col : (size_t)pixel) * 3;
            out[0] = r;
            out[1] = g;
            out[2] = b;
        }
        __syncthreads();
    }
    if (COUNT) {
        for (int k = 0; k < CN_COUNT; k++)
            if (cn[k]) atomicAdd(p.counters + k, cn[k]);
    }
#ifdef WAVE2_TIMING
    if (lane == 0)
        for (int k = 0; k < 15; k++) atomicAdd(p.counters + k, (unsigned long long)t_acc[k]);
#endif
}

}  // namespace

// Applies when a CTA can own a pixel (spp >= 256; the per-pixel drain tail is ~2 % at 16384 spp), depth <= 8, at most
// 255 materials, and the scene either goes through the BVH or has only spheres (at most FLUX_CULL_MAX = 128) and planes.
static size_t w2_smem_for(const RenderParams &p) {
    return p.scene.use_bvh ? w2_smem_bytes(0, 0, p.scene.n_materials, p.cam.max_depth)
                           : w2_smem_bytes(p.scene.n_spheres, p.scene.n_planes, p.scene.n_materials, p.cam.max_depth);
}

bool wave2_kernel_applicable(const RenderParams &p) {
    if (p.ss.n < WAVE2_MIN_SPP || p.scene.n_materials > 255 || p.cam.max_depth < 1 || p.cam.max_depth > WAVE2_MAX_DEPTH) return false;
    if (!p.scene.use_bvh && (p.scene.n_tris != 0 || p.scene.n_spheres > FLUX_CULL_MAX)) return false;
    return w2_smem_for(p) <= (size_t)(226 * 1024) / 3 - 1024;   // at least 3 CTAs per SM
}

void launch_render_wave2(const RenderParams &p, bool count, int sm_count, cudaStream_t stream) {
    const size_t smem = w2_smem_for(p);
    const uint64_t npix = (uint64_t)p.n_rows * p.cam.W;
    const uint64_t cap = (uint64_t)sm_count * WAVE2_MIN_BLOCKS;
    const int blocks = (int)(npix < cap ? (npix ? npix : 1) : cap);
    auto go = [&](auto kern) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kern<<<blocks, WAVE2_S, smem, stream>>>(p);
    };
    const bool bvh = p.scene.use_bvh != 0;
    if (count) {
        if (bvh) go(render_wave2_kernel<true, true>);
        else go(render_wave2_kernel<true, false>);
    } else {
        if (bvh) go(render_wave2_kernel<false, true>);
        else go(render_wave2_kernel<false, false>);
    }
}
