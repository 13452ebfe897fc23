// render_regen.cu — render kernel for high sample counts: one warp per pixel with
// in-warp path regeneration, scene resident in shared memory.
//
// Same arithmetic as render.cu (Camera::render trace.rs:53-97 → Scene::shade
// scene.rs:162-172 → Material::path_shade materials.rs:19-71), different
// scheduling.  ncu on the v1 kernel (profiles/r1a_render_v1_full.txt) shows
// 15.75 of 32 lanes active per instruction: paths end after 1..max_depth
// segments (2.39 on demo2) and finished lanes idle until the longest path of
// their batch ends.  Here every lane is a persistent path slot: when its path
// ends it adds the radiance to a lane-private sum and immediately takes the
// pixel's next sample index (ballot + popc rank over the lanes that need one),
// so the closest-hit loop always runs with (almost) all lanes live.
//
// Determinism: the order in which a lane receives sample indices depends only
// on the lengths of the paths, i.e. on the inputs; lane sums are combined by a
// fixed xor-shuffle tree.  A pixel's value therefore does not depend on grid
// size, scheduling or sharding.  It does differ from the v1
// kernel's value in the last bits (different summation order); both are within
// 1e-13 of the oracle's sequential sum.
//
// The scene (spheres, planes, materials) is staged once per CTA into shared
// memory as AoS records; bounding-box corners are addressed through per-ray
// near/far offsets so the slab test needs no selects.
#include "flux_bvh.cuh"
#include "flux_cull.cuh"
#include "flux_intersect.cuh"
#include "flux_kernels.cuh"
#include "flux_shade.cuh"

// build-time tuning knobs (A/B'd on the GPU)
#ifndef REGEN_THREADS
#define REGEN_THREADS 256
#endif
#ifndef REGEN_MIN_BLOCKS
#define REGEN_MIN_BLOCKS 3
#endif
#ifndef REGEN_TWO_PHASE
#define REGEN_TWO_PHASE 1
#endif
#ifndef REGEN_TRI_ONLY
#define REGEN_TRI_ONLY 1   // mesh scenes whose tree holds no sphere run the triangle-only traversal instantiation
#endif
#ifndef REGEN_CULL32
#define REGEN_CULL32 1   // phase 1 classifies the sphere boxes in FP32 (conservative; exact f64 test where undecided), as the
                         // wavefront and the direct kernel do (flux_cull.cuh): round 1 ran 12 f64 slab products per sphere
#endif

namespace {

// shared-memory sphere record (doubles): slab corners interleaved per axis so that the
// near/far corner of axis k is rec[2k + sign_k] / rec[2k + 1 - sign_k]
// stride 13 (odd): lanes reading the same field of different spheres in phase 2 hit distinct 8-byte banks
enum { R_C0X = 0, R_C1X, R_C0Y, R_C1Y, R_C0Z, R_C1Z, R_CX, R_CY, R_CZ, R_RR, R_R, R_INV, R_SPH_PAD, R_SPH_STRIDE };
enum { R_PPX = 0, R_PPY, R_PPZ, R_PNX, R_PNY, R_PNZ, R_PLN_STRIDE };

struct Rgb {
    double r, g, b;
};

struct SmemScene {
    const double *sph;       // [ns][R_SPH_STRIDE]
    const double *pln;       // [np][R_PLN_STRIDE]
    const uint32_t *sph_id;  // [ns] shape id
    const uint32_t *sph_mat; // [ns]
    const uint32_t *pln_id;
    const uint32_t *pln_mat;
    const DevMaterial *mat;
    uint32_t ns, np;
};

__host__ __device__ __forceinline__ size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Scene::hit (scene.rs:156-160) over the shared-memory scene.  Spheres first, in shape order,
// then planes; ties across kinds are resolved by shape id (common.rs:17-23 + min_by).
//
// Two phases (same tests, same values as the reference's per-shape sequence, different grouping):
//   1. BoundingBox::hit (shapes.rs:98-133) for EVERY sphere, branch-free, into a 64-bit pass mask.
//      Straight-line FP64 with 12 independent products per sphere: no divergence, ILP for the pipe.
//   2. each lane walks ITS OWN set bits in increasing sphere order and runs the quadratic
//      (shapes.rs:176-212) for that sphere.  The warp iterates max-over-lanes(popcount) times
//      instead of once per sphere that any lane's box test passed (v1: the quadratic ran for 85 % of
//      all spheres per warp with 11 of 32 lanes active; profiles/r1b_render_regen_full.txt).
template <bool COUNT>
__device__ __forceinline__ void closest_hit_smem(const RenderParams &p, const SmemScene &sc, V3 o, V3 d, double &best_t, uint32_t &best_id,
                                                 uint32_t &best_ref, unsigned long long *cn) {
    const double A = dot3(d, d);
    const double A2 = 2.0 * A, A4 = 4.0 * A;
    best_t = 0.0;
    best_id = 0xFFFFFFFFu;
    best_ref = 0;
#if REGEN_TWO_PHASE && REGEN_CULL32
    // ---- phase 1: conservative FP32 classification of every sphere box (constant-bank operands), exact test where undecided ----
    unsigned long long mask = 0ull;
    {
        const LinCullRay c = lin_cull_ray(p, o, d);
#pragma unroll 1
        for (uint32_t base = 0; base < sc.ns; base += 32u) {
            const uint32_t nsb = sc.ns - base < 32u ? sc.ns - base : 32u;
            uint32_t okm = 0u, failm = 0u;
#pragma unroll 2
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
            uint32_t m32 = okm & valid;
            uint32_t unc = ~(okm | failm) & valid;
            if (unc) {   // the exact BoundingBox::hit (shapes.rs:98-133); its reciprocals are formed on this rare path only
                const double ia = 1.0 / d.x, ib = 1.0 / d.y, ic = 1.0 / d.z;
                const int sx = ia >= 0.0 ? 0 : 1, sy = ib >= 0.0 ? 0 : 1, sz = ic >= 0.0 ? 0 : 1;
                while (unc) {
                    const uint32_t j = (uint32_t)__ffs((int)unc) - 1u;
                    unc &= unc - 1u;
                    const double *s = sc.sph + (size_t)(base + j) * R_SPH_STRIDE;
                    const double tx_min = (s[R_C0X + sx] - o.x) * ia, tx_max = (s[R_C1X - sx] - o.x) * ia;
                    const double ty_min = (s[R_C0Y + sy] - o.y) * ib, ty_max = (s[R_C1Y - sy] - o.y) * ib;
                    const double tz_min = (s[R_C0Z + sz] - o.z) * ic, tz_max = (s[R_C1Z - sz] - o.z) * ic;
                    const double t0 = ref_max(tx_min, ref_max(ty_min, tz_min));
                    const double t1 = ref_min(tx_max, ref_min(ty_max, tz_max));
                    if (t0 < t1 && t1 > FLUX_T_MIN) m32 |= 1u << j;
                }
            }
            mask |= (unsigned long long)m32 << base;
        }
        if (COUNT) {
            cn[CN_BBOX_TESTS] += sc.ns;
            cn[CN_BBOX_PASS] += __popcll(mask);
        }
    }
#elif REGEN_TWO_PHASE
    // ray-invariant terms (shapes.rs:107-122), hoisted
    const double ia = 1.0 / d.x, ib = 1.0 / d.y, ic = 1.0 / d.z;
    const int sx = ia >= 0.0 ? 0 : 1, sy = ib >= 0.0 ? 0 : 1, sz = ic >= 0.0 ? 0 : 1;
    // ---- phase 1: slab tests ----
    unsigned long long mask = 0ull;
    {
        const double *s = sc.sph;
#pragma unroll 4
        for (uint32_t i = 0; i < sc.ns; i++, s += R_SPH_STRIDE) {
            const double tx_min = (s[R_C0X + sx] - o.x) * ia, tx_max = (s[R_C1X - sx] - o.x) * ia;
            const double ty_min = (s[R_C0Y + sy] - o.y) * ib, ty_max = (s[R_C1Y - sy] - o.y) * ib;
            const double tz_min = (s[R_C0Z + sz] - o.z) * ic, tz_max = (s[R_C1Z - sz] - o.z) * ic;
            const double t0 = ref_max(tx_min, ref_max(ty_min, tz_min));
            const double t1 = ref_min(tx_max, ref_min(ty_max, tz_max));
            const bool pass = t0 < t1 && t1 > FLUX_T_MIN;
            mask |= (unsigned long long)pass << i;
        }
        if (COUNT) {
            cn[CN_BBOX_TESTS] += sc.ns;
            cn[CN_BBOX_PASS] += __popcll(mask);
        }
    }
#endif
#if REGEN_TWO_PHASE
    // ---- phase 2: quadratics of the lanes' own candidates, in increasing shape order ----
    while (mask != 0ull) {  // SIMT: the warp iterates max-over-lanes(popcount(mask)) times
        const uint32_t i = (uint32_t)__ffsll((long long)mask) - 1u;
        mask &= mask - 1ull;
        const double *s = sc.sph + (size_t)i * R_SPH_STRIDE;
        const V3 temp = mk3(o.x - s[R_CX], o.y - s[R_CY], o.z - s[R_CZ]);
        const double b = 2.0 * dot3(temp, d);
        const double c = dot3(temp, temp) - s[R_RR];
        const double disc = b * b - A4 * c;
        if (disc < 0.0) continue;
        if (COUNT) cn[CN_DISC_NONNEG]++;
        const double e = sqrt(disc);
        double t = (-b - e) / A2;
        if (!(t > FLUX_T_MIN)) {
            if (COUNT) cn[CN_T2]++;
            t = (-b + e) / A2;
            if (!(t > FLUX_T_MIN)) continue;
        }
        if (COUNT) cn[CN_CANDIDATES]++;
        // candidates arrive in increasing shape id: a later sphere wins only if strictly closer
        if (best_id == 0xFFFFFFFFu || t < best_t) {
            best_t = t;
            best_id = sc.sph_id[i];
            best_ref = i;
        }
    }
#else
    const double ia = 1.0 / d.x, ib = 1.0 / d.y, ic = 1.0 / d.z;
    const int sx = ia >= 0.0 ? 0 : 1, sy = ib >= 0.0 ? 0 : 1, sz = ic >= 0.0 ? 0 : 1;
    const double *s = sc.sph;
#pragma unroll 2
    for (uint32_t i = 0; i < sc.ns; i++, s += R_SPH_STRIDE) {
        if (COUNT) cn[CN_BBOX_TESTS]++;
        // BoundingBox::hit, shapes.rs:98-133
        const double tx_min = (s[R_C0X + sx] - o.x) * ia, tx_max = (s[R_C1X - sx] - o.x) * ia;
        const double ty_min = (s[R_C0Y + sy] - o.y) * ib, ty_max = (s[R_C1Y - sy] - o.y) * ib;
        const double tz_min = (s[R_C0Z + sz] - o.z) * ic, tz_max = (s[R_C1Z - sz] - o.z) * ic;
        const double t0 = ref_max(tx_min, ref_max(ty_min, tz_min));
        const double t1 = ref_min(tx_max, ref_min(ty_max, tz_max));
        if (!(t0 < t1 && t1 > FLUX_T_MIN)) continue;
        if (COUNT) cn[CN_BBOX_PASS]++;
        // Sphere::hit, shapes.rs:176-212
        const V3 temp = mk3(o.x - s[R_CX], o.y - s[R_CY], o.z - s[R_CZ]);
        const double b = 2.0 * dot3(temp, d);
        const double c = dot3(temp, temp) - s[R_RR];
        const double disc = b * b - A4 * c;
        if (disc < 0.0) continue;
        if (COUNT) cn[CN_DISC_NONNEG]++;
        const double e = sqrt(disc);
        double t = (-b - e) / A2;
        if (!(t > FLUX_T_MIN)) {
            if (COUNT) cn[CN_T2]++;
            t = (-b + e) / A2;
            if (!(t > FLUX_T_MIN)) continue;
        }
        if (COUNT) cn[CN_CANDIDATES]++;
        // spheres arrive in increasing shape id: a later sphere wins only if strictly closer
        if (best_id == 0xFFFFFFFFu || t < best_t) {
            best_t = t;
            best_id = sc.sph_id[i];
            best_ref = i;
        }
    }
#endif
    const double *pl = sc.pln;
    for (uint32_t i = 0; i < sc.np; i++, pl += R_PLN_STRIDE) {
        if (COUNT) cn[CN_PLANE_TESTS]++;
        // Plane::hit, shapes.rs:137-139
        const V3 pn = mk3(pl[R_PNX], pl[R_PNY], pl[R_PNZ]);
        const double t = dot3(mk3(pl[R_PPX] - o.x, pl[R_PPY] - o.y, pl[R_PPZ] - o.z), pn) / dot3(d, pn);
        if (!(t > FLUX_T_MIN)) continue;
        if (COUNT) cn[CN_CANDIDATES]++;
        const uint32_t id = sc.pln_id[i];
        if (best_id == 0xFFFFFFFFu || t < best_t || (t == best_t && id < best_id)) {
            best_t = t;
            best_id = id;
            best_ref = 0x80000000u | i;
        }
    }
}

// BVH = true: closest hit through the 4-wide BVH over the scene in global memory (meshes, many spheres) instead of
// the shared-memory linear scan; shared memory then holds the per-thread traversal stacks.
// TRI: BVH scenes whose tree holds triangles only (DevScene::bvh_tree_spheres == 0; BvhTraversal<·, false>).
template <bool COUNT, bool BVH, bool TRI = false>
__global__ void __launch_bounds__(REGEN_THREADS, REGEN_MIN_BLOCKS) render_regen_kernel(const __grid_constant__ RenderParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint2 *bvh_stack = reinterpret_cast<uint2 *>(smem_raw) + threadIdx.x;   // [BVH_STACK][blockDim.x] when BVH
    // ---- stage the scene into shared memory (once per CTA) ----
    const uint32_t ns = BVH ? 0u : p.scene.n_spheres, np = BVH ? 0u : p.scene.n_planes, nm = BVH ? 0u : p.scene.n_materials;
    double *s_sph = reinterpret_cast<double *>(smem_raw);
    double *s_pln = s_sph + (size_t)ns * R_SPH_STRIDE;
    DevMaterial *s_mat = reinterpret_cast<DevMaterial *>(s_pln + (size_t)np * R_PLN_STRIDE);
    uint32_t *s_sph_id = reinterpret_cast<uint32_t *>(s_mat + nm);
    uint32_t *s_sph_mat = s_sph_id + ns;
    uint32_t *s_pln_id = s_sph_mat + ns;
    uint32_t *s_pln_mat = s_pln_id + np;
    for (uint32_t k = threadIdx.x; k < ns; k += blockDim.x) {
        const double *g = p.scene.sph;
        double *r = s_sph + (size_t)k * R_SPH_STRIDE;
        r[R_C0X] = g[SPH_C0X * ns + k]; r[R_C1X] = g[SPH_C1X * ns + k];
        r[R_C0Y] = g[SPH_C0Y * ns + k]; r[R_C1Y] = g[SPH_C1Y * ns + k];
        r[R_C0Z] = g[SPH_C0Z * ns + k]; r[R_C1Z] = g[SPH_C1Z * ns + k];
        r[R_CX] = g[SPH_CX * ns + k]; r[R_CY] = g[SPH_CY * ns + k]; r[R_CZ] = g[SPH_CZ * ns + k];
        r[R_RR] = g[SPH_RR * ns + k]; r[R_R] = g[SPH_R * ns + k]; r[R_INV] = g[SPH_INV * ns + k];
        s_sph_id[k] = p.scene.sph_meta[k];
        s_sph_mat[k] = p.scene.sph_meta[ns + k];
    }
    for (uint32_t k = threadIdx.x; k < np; k += blockDim.x) {
        double *r = s_pln + (size_t)k * R_PLN_STRIDE;
        for (int f = 0; f < PLN_FIELDS; f++) r[f] = p.scene.pln[(size_t)f * np + k];
        s_pln_id[k] = p.scene.pln_meta[k];
        s_pln_mat[k] = p.scene.pln_meta[np + k];
    }
    for (uint32_t k = threadIdx.x; k < nm; k += blockDim.x) s_mat[k] = p.scene.materials[k];
    __syncthreads();
    SmemScene sc;
    sc.sph = s_sph; sc.pln = s_pln; sc.sph_id = s_sph_id; sc.sph_mat = s_sph_mat;
    sc.pln_id = s_pln_id; sc.pln_mat = s_pln_mat; sc.mat = s_mat; sc.ns = ns; sc.np = np;

    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t W = p.cam.W;
    const uint32_t npix = p.n_rows * W;
    const uint32_t n = p.ss.n;
    const uint32_t max_depth = p.cam.max_depth;
    const DevCamera &cam = p.cam;
    unsigned long long cn[COUNT ? CN_COUNT : 1];
    if (COUNT)
        for (int k = 0; k < CN_COUNT; k++) cn[k] = 0;
    const double pixel_denom = 1.0 / (double)((unsigned long long)p.ss.root * p.ss.root);  // trace.rs:59

    double sf[FLUX_MAX_DEPTH_CAP][4];  // per-path (f.r, f.g, f.b, weight) stack, innermost last

    for (;;) {
        uint32_t pixel = 0;
        if (lane == 0) pixel = atomicAdd(p.work_counter, 1u);
        pixel = __shfl_sync(0xffffffffu, pixel, 0);
        if (pixel >= npix) break;
        const uint32_t rk = pixel / W;
        const uint32_t col = pixel - rk * W;
        const uint32_t row = p.rows[rk];
        const uint32_t set = p.set_index[(size_t)row * W + col];
        const double2 *ps = p.ss.pixel + (size_t)set * n;
        const double2 *ds = p.ss.disc + (size_t)set * n;
        const double *hs = p.ss.hemi + (size_t)set * p.ss.max_depth * n * 3;
        // trace.rs:72-73 pixel-invariant parts
        const double colf = (double)col - cam.half_w;
        const double rowf = (double)(cam.H - row) - cam.half_h;

        Rgb acc = Rgb{0.0, 0.0, 0.0};
        uint32_t next = 0;  // next unassigned sample index of this pixel (warp-uniform)
        bool alive = false;
        V3 o = mk3(0, 0, 0), d = mk3(0, 0, 0);
        double2 psamp = make_double2(0.0, 0.0);
        uint32_t i = 0, depth = 1, top = 0;

        for (;;) {
            // ---- regeneration: dead lanes take the next sample indices in lane order ----
            const uint32_t need = __ballot_sync(0xffffffffu, !alive);
            if (!alive) {
                const uint32_t mine = next + __popc(need & lt_mask);
                if (mine < n) {
                    i = mine;
                    psamp = ps[i];
                    const double2 l = ds[i];
                    // trace.rs:72-80 + Camera::ray_direction trace.rs:44-51
                    const double u = cam.aps * (colf + psamp.x);
                    const double v = cam.aps * (rowf + psamp.y);
                    const double lpx = l.x * cam.lens_radius;
                    const double lpy = l.y * cam.lens_radius;
                    const double px2 = u * cam.factor;
                    const double py2 = v * cam.factor;
                    d = normalize3(((px2 - lpx) * cam.u + (py2 - lpy) * cam.v) - cam.focal_w);
                    o = (cam.eye + lpx * cam.u) + lpy * cam.v;
                    depth = 1;
                    top = 0;
                    alive = true;
                    if (COUNT) cn[CN_SAMPLES]++;
                }
            }
            next += __popc(need);
            if (!__any_sync(0xffffffffu, alive)) break;
            if (!alive) continue;

            // ---- one segment: Scene::shade, scene.rs:162-172 ----
            bool term = false;
            Rgb L = Rgb{0.0, 0.0, 0.0};
            if (depth > max_depth) {  // scene.rs:164-165
                if (COUNT) cn[CN_DEPTH_CUT]++;
                term = true;
            } else {
                if (COUNT) cn[CN_SEGMENTS]++;
                double t = 0.0;
                uint32_t hid, href = 0;
                RayCtx rc;
                HitRef href_bvh;
                if (BVH) {
                    rc = make_ray(o, d);
                    href_bvh = closest_hit_bvh<COUNT, !TRI>(p.scene, rc, bvh_stack, blockDim.x, cn);
                    hid = href_bvh.shape_id;
                    t = href_bvh.t;
                } else {
                    closest_hit_smem<COUNT>(p, sc, o, d, t, hid, href, cn);
                }
                if (hid == 0xFFFFFFFFu) {  // scene.rs:168
                    if (COUNT) cn[CN_MISS]++;
                    L = Rgb{cam.bg[0], cam.bg[1], cam.bg[2]};
                    term = true;
                } else {
                    // hit record of the closest hit only (shapes.rs:140-147,191-198)
                    V3 normal;
                    uint32_t mi;
                    V3 point;
                    if (BVH) {
                        const HitRec hrec = build_hit(p.scene, rc, href_bvh);
                        normal = hrec.normal;
                        point = hrec.point;
                        mi = hrec.material;
                        if (COUNT) cn[href_bvh.kind == KIND_SPHERE ? CN_HIT_SPHERE : (href_bvh.kind == KIND_PLANE ? CN_HIT_PLANE : CN_HIT_TRI)]++;
                    } else if (href & 0x80000000u) {
                        point = o + t * d;
                        const uint32_t k = href & 0x7FFFFFFFu;
                        const double *pl = sc.pln + (size_t)k * R_PLN_STRIDE;
                        normal = mk3(pl[R_PNX], pl[R_PNY], pl[R_PNZ]);
                        mi = sc.pln_mat[k];
                        if (COUNT) cn[CN_HIT_PLANE]++;
                    } else {
                        point = o + t * d;
                        const double *s = sc.sph + (size_t)href * R_SPH_STRIDE;
                        const V3 temp = mk3(o.x - s[R_CX], o.y - s[R_CY], o.z - s[R_CZ]);
                        normal = ((temp + t * d) * s[R_INV]) / s[R_R];
                        mi = sc.sph_mat[href];
                        if (COUNT) cn[CN_HIT_SPHERE]++;
                    }
                    const DevMaterial &m = BVH ? p.scene.materials[mi] : sc.mat[mi];
                    const uint32_t kind = m.kind;
                    double c0 = m.c[0], c1 = m.c[1], c2 = m.c[2];
                    if (kind == FLUX_MAT_EMISSIVE) {  // materials.rs:42-49
                        if (COUNT) cn[CN_EMISSIVE]++;
                        if (dot3(normal * -1.0, d) > 0.0) L = Rgb{c0, c1, c2};
                        term = true;
                    } else {
                        V3 wi;
                        double weight;
                        if (kind == FLUX_MAT_MATTE) {  // materials.rs:19-33
                            if (COUNT) cn[CN_MATTE]++;
                            const double *hp = hs + ((size_t)(depth - 1) * n + i) * 3;
                            matte_sample(normal, mk3(hp[0], hp[1], hp[2]), wi, weight);
                        } else if (kind == FLUX_MAT_REFLECTIVE) {  // materials.rs:57-71, brdf.rs:39-45
                            if (COUNT) cn[CN_SPECULAR]++;
                            specular_sample(normal, d, wi, weight);
                        } else {  // materials.rs:57-71, brdf.rs:55-78
                            if (COUNT) cn[CN_GLOSSY]++;
                            double lobe;
                            bool flipped;
                            if (p.ss.ghemi) {  // lobe table: to_unit_hemi(pixel sample, exp) precomputed (same bits)
                                const double *gh = p.ss.ghemi + (((size_t)set * p.ss.gk + m.gidx) * n + i) * 3;
                                glossy_sample_hs(normal, d, mk3(gh[0], gh[1], gh[2]), m.exp, wi, weight, lobe, flipped);
                            } else {
                                glossy_sample(normal, d, psamp.x, psamp.y, m.exp, m.inv_e1, wi, weight, lobe, flipped);
                            }
                            if (COUNT && flipped) cn[CN_GLOSSY_FLIP]++;
                            c0 *= lobe;
                            c1 *= lobe;
                            c2 *= lobe;
                        }
                        sf[top][0] = c0;
                        sf[top][1] = c1;
                        sf[top][2] = c2;
                        sf[top][3] = weight;
                        top++;
                        o = point;
                        d = wi;
                        depth++;
                    }
                }
            }
            if (term) {
                while (top > 0) {  // (f (*) L) * w, innermost first: materials.rs:31-32,69-70
                    top--;
                    L.r = (sf[top][0] * L.r) * sf[top][3];
                    L.g = (sf[top][1] * L.g) * sf[top][3];
                    L.b = (sf[top][2] * L.b) * sf[top][3];
                }
                acc.r += L.r;  // trace.rs:82
                acc.g += L.g;
                acc.b += L.b;
                alive = false;
            }
        }
        // ---- fixed-shape reduction of the 32 lane sums, then trace.rs:85-86 + color.rs:35-44 ----
#pragma unroll
        for (uint32_t off = 16; off > 0; off >>= 1) {
            acc.r += __shfl_xor_sync(0xffffffffu, acc.r, off);
            acc.g += __shfl_xor_sync(0xffffffffu, acc.g, off);
            acc.b += __shfl_xor_sync(0xffffffffu, acc.b, off);
        }
        if (lane == 0) {
            double r = acc.r * pixel_denom, gg = acc.g * pixel_denom, b = acc.b * pixel_denom;
            const double mx1 = r > gg ? r : gg;
            const double mx2 = mx1 > b ? mx1 : b;
            if (mx2 > 1.0) {
                const double inv = 1.0 / mx2;
                r *= inv;
                gg *= inv;
                b *= inv;
            }
            double *out = p.out + (p.out_by_row ? (size_t)row * W + col : (size_t)pixel) * 3;
            out[0] = r;
            out[1] = gg;
            out[2] = b;
        }
    }
    if (COUNT) {
        for (int k = 0; k < CN_COUNT; k++)
            if (cn[k]) atomicAdd(p.counters + k, cn[k]);
    }
}

size_t regen_smem_bytes(const DevScene &sc) {
    if (sc.use_bvh) return (size_t)BVH_STACK * REGEN_THREADS * sizeof(uint2);   // per-thread traversal stacks
    size_t b = (size_t)sc.n_spheres * R_SPH_STRIDE * 8 + (size_t)sc.n_planes * R_PLN_STRIDE * 8 +
               (size_t)sc.n_materials * sizeof(DevMaterial) + (size_t)(2 * sc.n_spheres + 2 * sc.n_planes) * 4;
    return align_up(b, 16);
}

}  // namespace

// The regeneration kernel applies when a warp can own a pixel (spp >= 64) and either the scene goes through the
// BVH (meshes, more than 40 bounded shapes) or it has only spheres and planes and fits in shared memory next to
// two CTAs per SM.
bool regen_kernel_applicable(const RenderParams &p) {
    if (p.ss.n < 64) return false;
    if (p.scene.use_bvh) return true;
    return p.scene.n_tris == 0 && p.scene.n_spheres <= 64 && regen_smem_bytes(p.scene) <= 96 * 1024;
}

void launch_render_regen(const RenderParams &p, bool count, int sm_count, cudaStream_t stream) {
    const size_t smem = regen_smem_bytes(p.scene);
    const int threads = REGEN_THREADS;
    const uint64_t npix = (uint64_t)p.n_rows * p.cam.W;
    uint64_t want = (npix + (threads / 32) - 1) / (threads / 32);
    uint64_t cap = (uint64_t)sm_count * REGEN_MIN_BLOCKS;
    int blocks = (int)(want < cap ? (want ? want : 1) : cap);
    auto go = [&](auto kern) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kern<<<blocks, threads, smem, stream>>>(p);
    };
    const bool bvh = p.scene.use_bvh != 0;
    const bool tri = bvh && p.scene.bvh_tree_spheres == 0 && REGEN_TRI_ONLY;
    if (count) {
        if (tri) go(render_regen_kernel<true, true, true>);
        else if (bvh) go(render_regen_kernel<true, true>);
        else go(render_regen_kernel<true, false>);
    } else {
        if (tri) go(render_regen_kernel<false, true, true>);
        else if (bvh) go(render_regen_kernel<false, true>);
        else go(render_regen_kernel<false, false>);
    }
}
