// render_wave.cu — block-local wavefront render kernel for very high sample counts.
//
// Same arithmetic as render.cu / render_regen.cu (Camera::render trace.rs:53-97 → Scene::shade
// scene.rs:162-172 → Material::path_shade materials.rs:19-71); what changes is WHO evaluates WHAT, WHEN.
// A CTA owns one pixel and a pool of S = 256 path slots whose state lives in shared memory.  Each
// iteration advances every live path by one segment through stages separated by __syncthreads():
//
//   A regenerate   dead slots take the pixel's next sample indices (ballot + block prefix sum) and get
//                  their camera ray (trace.rs:72-80)
//   B box tests    slot owner: BoundingBox::hit (shapes.rs:98-133) against every sphere → bit mask;
//                  prefix sum over popcounts → compact list of (slot, sphere) candidate pairs
//   C quadratics   one thread per PAIR: Sphere::hit's quadratic (shapes.rs:176-212); the work is the sum
//                  of candidates, not (spheres any lane touched) × 32 lanes
//   D select       slot owner: closest candidate in shape order (scene.rs:156-160, common.rs:17-23), planes,
//                  hit record, terminal materials (Emissive, miss, depth cut: unwind + accumulate)
//   E shade        surviving hits are binned by material kind with one packed prefix sum; one thread per
//                  ITEM in kind order, so Matte / PerfectSpecular / GlossySpecular (pow-heavy) run in
//                  coherent warps (materials.rs:19-71, brdf.rs:20-78)
//
// ncu on the warp-per-pixel kernel showed why (profiles/r1b_render_regen_full.txt): the quadratic ran with
// 11 of 32 lanes, glossy shading with 8.5, pow with 8.5; 3400 instructions of straight-through code gave
// 1.6 stalled-for-instruction-fetch warps per issue.  Here every stage is a short dense loop.
//
// Determinism: every prefix sum is a function of the inputs only (no atomics), slot sums are combined by a
// fixed tree, so a pixel's value is independent of grid size, scheduling and sharding.
#include "flux_kernels.cuh"
#include "flux_shade.cuh"

#ifndef WAVE_MIN_BLOCKS
#define WAVE_MIN_BLOCKS 3
#endif
#define WAVE_S 256          // path slots per CTA == threads per CTA
#define WAVE_KCAP 1536      // candidate pairs processed per chunk
#define WAVE_MAX_DEPTH 8    // deeper jobs use the other kernels

namespace {

enum { W_C0X = 0, W_C1X, W_C0Y, W_C1Y, W_C0Z, W_C1Z, W_CX, W_CY, W_CZ, W_RR, W_R, W_INV, W_PAD, W_SPH_STRIDE };
enum { W_PPX = 0, W_PPY, W_PPZ, W_PNX, W_PNY, W_PNZ, W_PLN_STRIDE };
enum { ST_DEAD = 0, ST_ALIVE = 1 };
enum { K_NONE = 0, K_MATTE = 1, K_SPEC = 2, K_GLOSSY = 3 };

struct Rgb {
    double r, g, b;
};

// meta word of a slot: depth (bits 0-7) | top (8-15) | material of the pending hit (16-27) | state (28)
__device__ __forceinline__ uint32_t meta_pack(uint32_t depth, uint32_t top, uint32_t mat, uint32_t st) {
    return depth | (top << 8) | (mat << 16) | (st << 28);
}
__device__ __forceinline__ uint32_t meta_depth(uint32_t m) { return m & 0xFFu; }
__device__ __forceinline__ uint32_t meta_top(uint32_t m) { return (m >> 8) & 0xFFu; }
__device__ __forceinline__ uint32_t meta_mat(uint32_t m) { return (m >> 16) & 0xFFFu; }
__device__ __forceinline__ uint32_t meta_state(uint32_t m) { return (m >> 28) & 1u; }

struct WaveSmem {
    // scene
    double *sph, *pln;
    DevMaterial *mat;
    uint32_t *sph_id, *sph_mat, *pln_id, *pln_mat;
    // path slots
    double *ox, *oy, *oz, *dx, *dy, *dz, *a2, *a4, *nx, *ny, *nz;
    double *stk_w, *stk_lobe;   // [depth][slot]
    uint8_t *stk_mat;           // [depth][slot]
    uint32_t *si, *meta;
    // work lists
    uint32_t *pairs;            // (slot << 8) | sphere
    double *pair_t;             // hit distance or NaN
    uint32_t *shade;            // slot ids in kind order
    uint32_t *scratch;          // [4][8] warp totals (rotating)
};

__host__ __device__ inline size_t wave_smem_bytes(uint32_t ns, uint32_t np, uint32_t nm, uint32_t max_depth) {
    size_t b = 0;
    b += (size_t)ns * W_SPH_STRIDE * 8 + (size_t)np * W_PLN_STRIDE * 8;
    b += (size_t)nm * sizeof(DevMaterial);
    b += (size_t)(2 * ns + 2 * np) * 4;
    b = (b + 15) / 16 * 16;
    b += (size_t)11 * WAVE_S * 8;                      // ray, a2, a4, normal
    b += (size_t)2 * max_depth * WAVE_S * 8;           // stack weight, lobe
    b += (size_t)WAVE_KCAP * 8;                        // pair_t
    b += (size_t)2 * WAVE_S * 4;                       // si, meta
    b += (size_t)WAVE_KCAP * 4 + (size_t)WAVE_S * 4;   // pairs, shade
    b += 4 * 8 * 4;                                    // scratch
    b += (size_t)max_depth * WAVE_S;                   // stack material
    return (b + 15) / 16 * 16;
}

__device__ __forceinline__ WaveSmem carve(unsigned char *raw, uint32_t ns, uint32_t np, uint32_t nm, uint32_t max_depth) {
    WaveSmem w;
    double *d = reinterpret_cast<double *>(raw);
    w.sph = d; d += (size_t)ns * W_SPH_STRIDE;
    w.pln = d; d += (size_t)np * W_PLN_STRIDE;
    w.mat = reinterpret_cast<DevMaterial *>(d);
    uint32_t *u = reinterpret_cast<uint32_t *>(w.mat + nm);
    w.sph_id = u; u += ns;
    w.sph_mat = u; u += ns;
    w.pln_id = u; u += np;
    w.pln_mat = u; u += np;
    size_t off = (size_t)(reinterpret_cast<unsigned char *>(u) - raw);
    off = (off + 15) / 16 * 16;
    d = reinterpret_cast<double *>(raw + off);
    w.ox = d; d += WAVE_S; w.oy = d; d += WAVE_S; w.oz = d; d += WAVE_S;
    w.dx = d; d += WAVE_S; w.dy = d; d += WAVE_S; w.dz = d; d += WAVE_S;
    w.a2 = d; d += WAVE_S; w.a4 = d; d += WAVE_S;
    w.nx = d; d += WAVE_S; w.ny = d; d += WAVE_S; w.nz = d; d += WAVE_S;
    w.stk_w = d; d += (size_t)max_depth * WAVE_S;
    w.stk_lobe = d; d += (size_t)max_depth * WAVE_S;
    w.pair_t = d; d += WAVE_KCAP;
    u = reinterpret_cast<uint32_t *>(d);
    w.si = u; u += WAVE_S;
    w.meta = u; u += WAVE_S;
    w.pairs = u; u += WAVE_KCAP;
    w.shade = u; u += WAVE_S;
    w.scratch = u; u += 32;
    w.stk_mat = reinterpret_cast<uint8_t *>(u);
    return w;
}

// exclusive prefix sum of v over the CTA in thread order + total; one barrier; `scr` is an 8-word scratch row
// that must not be reused before the next barrier.
__device__ __forceinline__ uint32_t block_scan(uint32_t v, uint32_t *scr, uint32_t &total) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (uint32_t)o) incl += y;
    }
    if (lane == 31) scr[warp] = incl;
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (uint32_t w = 0; w < WAVE_S / 32; w++) {
        const uint32_t c = scr[w];
        if (w < warp) base += c;
        tot += c;
    }
    total = tot;
    return base + incl - v;
}

template <bool COUNT>
__global__ void __launch_bounds__(WAVE_S, WAVE_MIN_BLOCKS) render_wave_kernel(const __grid_constant__ RenderParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t ns = p.scene.n_spheres, np = p.scene.n_planes, nm = p.scene.n_materials;
    const uint32_t max_depth = p.cam.max_depth;
    const WaveSmem w = carve(smem_raw, ns, np, nm, max_depth);
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    // ---- stage the scene (once per CTA) ----
    for (uint32_t k = tid; k < ns; k += WAVE_S) {
        const double *g = p.scene.sph;
        double *r = w.sph + (size_t)k * W_SPH_STRIDE;
        r[W_C0X] = g[SPH_C0X * ns + k]; r[W_C1X] = g[SPH_C1X * ns + k];
        r[W_C0Y] = g[SPH_C0Y * ns + k]; r[W_C1Y] = g[SPH_C1Y * ns + k];
        r[W_C0Z] = g[SPH_C0Z * ns + k]; r[W_C1Z] = g[SPH_C1Z * ns + k];
        r[W_CX] = g[SPH_CX * ns + k]; r[W_CY] = g[SPH_CY * ns + k]; r[W_CZ] = g[SPH_CZ * ns + k];
        r[W_RR] = g[SPH_RR * ns + k]; r[W_R] = g[SPH_R * ns + k]; r[W_INV] = g[SPH_INV * ns + k];
        r[W_PAD] = 0.0;
        w.sph_id[k] = p.scene.sph_meta[k];
        w.sph_mat[k] = p.scene.sph_meta[ns + k];
    }
    for (uint32_t k = tid; k < np; k += WAVE_S) {
        double *r = w.pln + (size_t)k * W_PLN_STRIDE;
        for (int f = 0; f < PLN_FIELDS; f++) r[f] = p.scene.pln[(size_t)f * np + k];
        w.pln_id[k] = p.scene.pln_meta[k];
        w.pln_mat[k] = p.scene.pln_meta[np + k];
    }
    for (uint32_t k = tid; k < nm; k += WAVE_S) w.mat[k] = p.scene.materials[k];
    __syncthreads();

    const DevCamera &cam = p.cam;
    const uint32_t W = cam.W;
    const uint32_t npix = p.n_rows * W;
    const uint32_t n = p.ss.n;
    unsigned long long cn[COUNT ? CN_COUNT : 1];
    if (COUNT)
        for (int k = 0; k < CN_COUNT; k++) cn[k] = 0;
    const double pixel_denom = 1.0 / (double)((unsigned long long)p.ss.root * p.ss.root);  // trace.rs:59
    __shared__ uint32_t s_pixel;
    __shared__ double s_red[3][WAVE_S / 32];
    uint32_t rot = 0;  // rotating scratch row

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
        const double colf = (double)col - cam.half_w;          // trace.rs:72
        const double rowf = (double)(cam.H - row) - cam.half_h; // trace.rs:73

        Rgb acc = Rgb{0.0, 0.0, 0.0};   // radiance of the paths hosted by slot `tid`, in completion order
        uint32_t next = 0;              // next unassigned sample index (identical in all threads)
        w.meta[tid] = meta_pack(0, 0, 0, ST_DEAD);
        __syncthreads();

        for (;;) {
            // ================= A: regenerate =================
            uint32_t m = w.meta[tid];
            const bool dead = meta_state(m) == ST_DEAD;
            const uint32_t bal = __ballot_sync(0xffffffffu, dead);
            uint32_t *scr = w.scratch + 8 * (rot++ & 3u);
            if (lane == 0) scr[warp] = __popc(bal);
            __syncthreads();
            uint32_t base = 0, n_dead = 0;
#pragma unroll
            for (uint32_t q = 0; q < WAVE_S / 32; q++) {
                const uint32_t c = scr[q];
                if (q < warp) base += c;
                n_dead += c;
            }
            const uint32_t started = next >= n ? 0u : min(n_dead, n - next);
            if (WAVE_S - n_dead + started == 0) break;  // no live path and no sample left (uniform)
            V3 o, d;
            if (dead) {
                const uint32_t i = next + base + __popc(bal & ((1u << lane) - 1u));
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
                    d = normalize3(((px2 - lpx) * cam.u + (py2 - lpy) * cam.v) - cam.focal_w);
                    o = (cam.eye + lpx * cam.u) + lpy * cam.v;
                    w.ox[tid] = o.x; w.oy[tid] = o.y; w.oz[tid] = o.z;
                    w.dx[tid] = d.x; w.dy[tid] = d.y; w.dz[tid] = d.z;
                    w.si[tid] = i;
                    m = meta_pack(1, 0, 0, ST_ALIVE);
                    w.meta[tid] = m;
                    if (COUNT) cn[CN_SAMPLES]++;
                }
            } else {
                o = mk3(w.ox[tid], w.oy[tid], w.oz[tid]);
                d = mk3(w.dx[tid], w.dy[tid], w.dz[tid]);
            }
            next += n_dead;
            const bool alive = meta_state(m) == ST_ALIVE;
            const bool cut = alive && meta_depth(m) > max_depth;  // scene.rs:164-165
            const bool trace = alive && !cut;

            // ================= B: box tests (owner) =================
            unsigned long long mask = 0ull;
            if (trace) {
                if (COUNT) cn[CN_SEGMENTS]++;
                // ray-invariant terms (shapes.rs:107-122,177,180,187), hoisted
                const double ia = 1.0 / d.x, ib = 1.0 / d.y, ic = 1.0 / d.z;
                const int sx = ia >= 0.0 ? 0 : 1, sy = ib >= 0.0 ? 0 : 1, sz = ic >= 0.0 ? 0 : 1;
                const double A = dot3(d, d);
                w.a2[tid] = 2.0 * A;
                w.a4[tid] = 4.0 * A;
                const double *s = w.sph;
#pragma unroll 4
                for (uint32_t i = 0; i < ns; i++, s += W_SPH_STRIDE) {
                    const double tx_min = (s[W_C0X + sx] - o.x) * ia, tx_max = (s[W_C1X - sx] - o.x) * ia;
                    const double ty_min = (s[W_C0Y + sy] - o.y) * ib, ty_max = (s[W_C1Y - sy] - o.y) * ib;
                    const double tz_min = (s[W_C0Z + sz] - o.z) * ic, tz_max = (s[W_C1Z - sz] - o.z) * ic;
                    const double t0 = ref_max(tx_min, ref_max(ty_min, tz_min));
                    const double t1 = ref_min(tx_max, ref_min(ty_max, tz_max));
                    const bool pass = t0 < t1 && t1 > FLUX_T_MIN;
                    mask |= (unsigned long long)pass << i;
                }
                if (COUNT) {
                    cn[CN_BBOX_TESTS] += ns;
                    cn[CN_BBOX_PASS] += __popcll(mask);
                }
            }
            const uint32_t count = (uint32_t)__popcll(mask);
            uint32_t K;
            const uint32_t off = block_scan(count, w.scratch + 8 * (rot++ & 3u), K);

            // ================= C: quadratics, one thread per candidate pair =================
            double best_t = 0.0;
            uint32_t best_ref = 0xFFFFFFFFu;  // sphere index; plane = 0x80000000 | index; none = 0xFFFFFFFF
            for (uint32_t cbase = 0; cbase < K; cbase += WAVE_KCAP) {
                {   // owners publish their pairs that fall into this chunk
                    unsigned long long mm = mask;
                    uint32_t q = off;
                    while (mm) {
                        const uint32_t j = (uint32_t)__ffsll((long long)mm) - 1u;
                        mm &= mm - 1ull;
                        if (q >= cbase && q < cbase + WAVE_KCAP) w.pairs[q - cbase] = (tid << 8) | j;
                        q++;
                    }
                }
                __syncthreads();
                const uint32_t kn = min((uint32_t)WAVE_KCAP, K - cbase);
                for (uint32_t q = tid; q < kn; q += WAVE_S) {
                    const uint32_t pr = w.pairs[q];
                    const uint32_t sl = pr >> 8, j = pr & 0xFFu;
                    const double *s = w.sph + (size_t)j * W_SPH_STRIDE;
                    const V3 ro = mk3(w.ox[sl], w.oy[sl], w.oz[sl]);
                    const V3 rd = mk3(w.dx[sl], w.dy[sl], w.dz[sl]);
                    // Sphere::hit, shapes.rs:176-212
                    const V3 temp = mk3(ro.x - s[W_CX], ro.y - s[W_CY], ro.z - s[W_CZ]);
                    const double b = 2.0 * dot3(temp, rd);
                    const double c = dot3(temp, temp) - s[W_RR];
                    const double disc = b * b - w.a4[sl] * c;
                    double t = __longlong_as_double(0x7FF8000000000000ll);  // NaN = no hit
                    if (!(disc < 0.0)) {
                        if (COUNT) cn[CN_DISC_NONNEG]++;
                        const double e = sqrt(disc);
                        const double A2 = w.a2[sl];
                        double t1 = (-b - e) / A2;
                        if (!(t1 > FLUX_T_MIN)) {
                            if (COUNT) cn[CN_T2]++;
                            t1 = (-b + e) / A2;
                        }
                        if (t1 > FLUX_T_MIN) t = t1;
                    }
                    w.pair_t[q] = t;
                }
                __syncthreads();
                {   // owners fold their candidates of this chunk, in increasing shape order
                    unsigned long long mm = mask;
                    uint32_t q = off;
                    while (mm) {
                        const uint32_t j = (uint32_t)__ffsll((long long)mm) - 1u;
                        mm &= mm - 1ull;
                        if (q >= cbase && q < cbase + WAVE_KCAP) {
                            const double t = w.pair_t[q - cbase];
                            if (t == t) {  // not NaN
                                if (COUNT) cn[CN_CANDIDATES]++;
                                // a later sphere wins only if strictly closer (common.rs:17-23 + min_by)
                                if (best_ref == 0xFFFFFFFFu || t < best_t) {
                                    best_t = t;
                                    best_ref = j;
                                }
                            }
                        }
                        q++;
                    }
                }
                // the next chunk rewrites `pairs` (read only in the compute loop, finished at the barrier above)
                // and `pair_t` (rewritten only after the next barrier)
            }

            // ================= D: select, hit record, terminal materials (owner) =================
            uint32_t kind = K_NONE;
            if (alive) {
                bool term = false;
                Rgb L = Rgb{0.0, 0.0, 0.0};
                uint32_t top = meta_top(m);
                const uint32_t depth = meta_depth(m);
                if (cut) {
                    if (COUNT) cn[CN_DEPTH_CUT]++;
                    term = true;
                } else {
                    uint32_t best_id = best_ref == 0xFFFFFFFFu ? 0xFFFFFFFFu : w.sph_id[best_ref];
                    const double *pl = w.pln;
                    for (uint32_t i = 0; i < np; i++, pl += W_PLN_STRIDE) {
                        if (COUNT) cn[CN_PLANE_TESTS]++;
                        // Plane::hit, shapes.rs:137-139
                        const V3 pn = mk3(pl[W_PNX], pl[W_PNY], pl[W_PNZ]);
                        const double t = dot3(mk3(pl[W_PPX] - o.x, pl[W_PPY] - o.y, pl[W_PPZ] - o.z), pn) / dot3(d, pn);
                        if (!(t > FLUX_T_MIN)) continue;
                        if (COUNT) cn[CN_CANDIDATES]++;
                        const uint32_t id = w.pln_id[i];
                        if (best_id == 0xFFFFFFFFu || t < best_t || (t == best_t && id < best_id)) {
                            best_t = t;
                            best_id = id;
                            best_ref = 0x80000000u | i;
                        }
                    }
                    if (best_id == 0xFFFFFFFFu) {  // scene.rs:168
                        if (COUNT) cn[CN_MISS]++;
                        L = Rgb{cam.bg[0], cam.bg[1], cam.bg[2]};
                        term = true;
                    } else {
                        // hit record of the closest hit only (shapes.rs:140-147,191-198)
                        V3 normal;
                        uint32_t mi;
                        const V3 point = o + best_t * d;
                        if (best_ref & 0x80000000u) {
                            const uint32_t k = best_ref & 0x7FFFFFFFu;
                            const double *q = w.pln + (size_t)k * W_PLN_STRIDE;
                            normal = mk3(q[W_PNX], q[W_PNY], q[W_PNZ]);
                            mi = w.pln_mat[k];
                            if (COUNT) cn[CN_HIT_PLANE]++;
                        } else {
                            const double *s = w.sph + (size_t)best_ref * W_SPH_STRIDE;
                            const V3 temp = mk3(o.x - s[W_CX], o.y - s[W_CY], o.z - s[W_CZ]);
                            normal = ((temp + best_t * d) * s[W_INV]) / s[W_R];
                            mi = w.sph_mat[best_ref];
                            if (COUNT) cn[CN_HIT_SPHERE]++;
                        }
                        const uint32_t mk = w.mat[mi].kind;
                        if (mk == FLUX_MAT_EMISSIVE) {  // materials.rs:42-49
                            if (COUNT) cn[CN_EMISSIVE]++;
                            if (dot3(normal * -1.0, d) > 0.0) L = Rgb{w.mat[mi].c[0], w.mat[mi].c[1], w.mat[mi].c[2]};
                            term = true;
                        } else {
                            kind = mk == FLUX_MAT_MATTE ? K_MATTE : (mk == FLUX_MAT_REFLECTIVE ? K_SPEC : K_GLOSSY);
                            w.nx[tid] = normal.x; w.ny[tid] = normal.y; w.nz[tid] = normal.z;
                            w.ox[tid] = point.x; w.oy[tid] = point.y; w.oz[tid] = point.z;  // child ray origin
                            w.meta[tid] = meta_pack(depth, top, mi, ST_ALIVE);
                        }
                    }
                }
                if (term) {
                    while (top > 0) {  // (f (*) L) * w, innermost first: materials.rs:31-32,69-70
                        top--;
                        const DevMaterial &sm = w.mat[w.stk_mat[(size_t)top * WAVE_S + tid]];
                        const double lobe = w.stk_lobe[(size_t)top * WAVE_S + tid];
                        const double wt = w.stk_w[(size_t)top * WAVE_S + tid];
                        L.r = ((sm.c[0] * lobe) * L.r) * wt;
                        L.g = ((sm.c[1] * lobe) * L.g) * wt;
                        L.b = ((sm.c[2] * lobe) * L.b) * wt;
                    }
                    acc.r += L.r;  // trace.rs:82
                    acc.g += L.g;
                    acc.b += L.b;
                    w.meta[tid] = meta_pack(0, 0, 0, ST_DEAD);
                }
            }

            // ================= E: material-sorted shading, one thread per item =================
            // one packed prefix sum gives every item its position in (matte | specular | glossy) order
            const uint32_t packed = kind == K_MATTE ? 1u : (kind == K_SPEC ? (1u << 10) : (kind == K_GLOSSY ? (1u << 20) : 0u));
            uint32_t tot;
            const uint32_t ex = block_scan(packed, w.scratch + 8 * (rot++ & 3u), tot);
            const uint32_t n_matte = tot & 0x3FFu, n_spec = (tot >> 10) & 0x3FFu, n_gloss = (tot >> 20) & 0x3FFu;
            if (kind == K_MATTE) w.shade[ex & 0x3FFu] = tid;
            else if (kind == K_SPEC) w.shade[n_matte + ((ex >> 10) & 0x3FFu)] = tid;
            else if (kind == K_GLOSSY) w.shade[n_matte + n_spec + ((ex >> 20) & 0x3FFu)] = tid;
            __syncthreads();
            if (tid < n_matte + n_spec + n_gloss) {
                const uint32_t sl = w.shade[tid];
                const uint32_t sm = w.meta[sl];
                const uint32_t depth = meta_depth(sm), top = meta_top(sm), mi = meta_mat(sm);
                const V3 normal = mk3(w.nx[sl], w.ny[sl], w.nz[sl]);
                const V3 dir = mk3(w.dx[sl], w.dy[sl], w.dz[sl]);
                const uint32_t i = w.si[sl];
                V3 wi;
                double weight, lobe = 1.0;
                if (tid < n_matte) {  // materials.rs:19-33
                    if (COUNT) cn[CN_MATTE]++;
                    const double *hp = hs + ((size_t)(depth - 1) * n + i) * 3;
                    matte_sample(normal, mk3(hp[0], hp[1], hp[2]), wi, weight);
                } else if (tid < n_matte + n_spec) {  // materials.rs:57-71, brdf.rs:39-45
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
                w.stk_w[(size_t)top * WAVE_S + sl] = weight;
                w.stk_lobe[(size_t)top * WAVE_S + sl] = lobe;
                w.stk_mat[(size_t)top * WAVE_S + sl] = (uint8_t)mi;
                w.dx[sl] = wi.x; w.dy[sl] = wi.y; w.dz[sl] = wi.z;
                w.meta[sl] = meta_pack(depth + 1, top + 1, 0, ST_ALIVE);
            }
            __syncthreads();
        }

        // ---- fixed-shape reduction of the 256 slot sums, then trace.rs:85-86 + color.rs:35-44 ----
#pragma unroll
        for (uint32_t o2 = 16; o2 > 0; o2 >>= 1) {
            acc.r += __shfl_xor_sync(0xffffffffu, acc.r, o2);
            acc.g += __shfl_xor_sync(0xffffffffu, acc.g, o2);
            acc.b += __shfl_xor_sync(0xffffffffu, acc.b, o2);
        }
        if (lane == 0) {
            s_red[0][warp] = acc.r;
            s_red[1][warp] = acc.g;
            s_red[2][warp] = acc.b;
        }
        __syncthreads();
        if (tid == 0) {
            double r = 0.0, g = 0.0, b = 0.0;
            // fixed pairwise tree over the 8 warp sums
            double tr[WAVE_S / 32], tg[WAVE_S / 32], tb[WAVE_S / 32];
            for (int k = 0; k < WAVE_S / 32; k++) { tr[k] = s_red[0][k]; tg[k] = s_red[1][k]; tb[k] = s_red[2][k]; }
            for (int st = 1; st < WAVE_S / 32; st <<= 1)
                for (int k = 0; k + st < WAVE_S / 32; k += 2 * st) { tr[k] += tr[k + st]; tg[k] += tg[k + st]; tb[k] += tb[k + st]; }
            r = tr[0] * pixel_denom; g = tg[0] * pixel_denom; b = tb[0] * pixel_denom;
            const double mx1 = r > g ? r : g;
            const double mx2 = mx1 > b ? mx1 : b;
            if (mx2 > 1.0) {
                const double inv = 1.0 / mx2;
                r *= inv;
                g *= inv;
                b *= inv;
            }
            double *out = p.out + (size_t)pixel * 3;
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
}

}  // namespace

// The wavefront kernel applies when a CTA can own a pixel (spp >= 4096 keeps the per-pixel drain tail
// under ~2 %), the scene has only spheres and planes, at most 64 spheres / 4096 materials, and depth <= 8.
bool wave_kernel_applicable(const RenderParams &p) {
    return p.ss.n >= 4096 && p.scene.n_tris == 0 && !p.scene.use_bvh && p.scene.n_spheres <= 64 &&
           p.scene.n_materials <= 255 && p.cam.max_depth >= 1 && p.cam.max_depth <= WAVE_MAX_DEPTH &&
           wave_smem_bytes(p.scene.n_spheres, p.scene.n_planes, p.scene.n_materials, p.cam.max_depth) <= 100 * 1024;
}

void launch_render_wave(const RenderParams &p, bool count, int sm_count, cudaStream_t stream) {
    const size_t smem = wave_smem_bytes(p.scene.n_spheres, p.scene.n_planes, p.scene.n_materials, p.cam.max_depth);
    const uint64_t npix = (uint64_t)p.n_rows * p.cam.W;
    const uint64_t cap = (uint64_t)sm_count * WAVE_MIN_BLOCKS;
    const int blocks = (int)(npix < cap ? (npix ? npix : 1) : cap);
    if (count) {
        cudaFuncSetAttribute(render_wave_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        render_wave_kernel<true><<<blocks, WAVE_S, smem, stream>>>(p);
    } else {
        cudaFuncSetAttribute(render_wave_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        render_wave_kernel<false><<<blocks, WAVE_S, smem, stream>>>(p);
    }
}
