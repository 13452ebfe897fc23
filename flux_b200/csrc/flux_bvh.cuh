// flux_bvh.cuh — EXTENSION: 4-wide bounding-volume hierarchy over sphere and triangle
// boxes, and its closest-hit traversal.  The reference has no acceleration structure: Scene::hit is a
// linear scan (fluxcore/src/scene.rs:156-160).  The contract of this file is therefore "return exactly
// what the linear scan returns" — same shape id (ties to the lower id, common.rs:17-23) and the same
// bits of t — only faster for scenes with thousands to millions of shapes.
//
// Why the result is identical:
//   * primitives reached in a leaf are tested by the very same device functions as the linear scan
//     (sphere_t = BoundingBox::hit + quadratic, tri_t) and folded with the order-independent consider();
//   * a node may be skipped only if no primitive below it can yield a candidate.  Node boxes are unions
//     of the primitive boxes (spheres: centre -+ r exactly as Sphere::new, shapes.rs:154-169; triangles:
//     vertex min/max) grown by BVH_PAD_REL * scene extent.  The slab expression (c - o) * (1/d) is the
//     one BoundingBox::hit uses and is monotone in c under IEEE rounding, so a box that contains a
//     primitive box passes whenever the primitive's own test passes; NaN slabs (0 * inf) never cull
//     (fmax/fmin ignore NaN, the comparisons below are false for NaN);
//   * distance pruning compares the node's entry distance with t_best plus a margin of 1e-9 relative
//     (seven orders of magnitude above the rounding difference between a quadratic root and a slab
//     entry), so a root that rounds to just in front of its own box is never lost.
// Planes are unbounded and a few very large spheres (environment spheres) would bloat the top of the
// tree: both are kept in a linear list that is tested before the traversal.
//
// Node boxes are stored and tested in FP32 — conservatively: the builder rounds the (padded) f64 child boxes outward
// to f32, the traversal forms each slab value as fma(c, 1/d, -(o/d)) in f32 with the operand -(o/d) lowered (entry)
// or raised (exit) by a per-axis bound E_k >= the f32 evaluation error (same derivation as the box classification
// of render_wave2.cu), so the f32 entry distance never exceeds and the f32 exit distance never falls below the
// value the reference's f64 expression would give for the same box.  A node is skipped only on those one-sided
// bounds; zero / denormal direction components use a clamped reciprocal of magnitude 1e30 with the sign of d_k
// (a box containing o_k on that axis then imposes no constraint, one not containing it is missed — what 1/0 = inf
// does in the reference), non-finite rays cull nothing.  Leaves run the exact f64 primitive tests.
//
// Layout: BvhNode4 is 128 B (one line): child boxes SoA (float4 per axis and bound), 4 child references; a child
// reference is an inner node index, a leaf (offset, count) into bvh_prims, or EMPTY.  Leaf primitives are fetched
// from 112 B / 96 B array-of-structure records (one or two lines per primitive) instead of the 12 strided planes
// of the linear-scan SoA.
#pragma once
#include <math_constants.h>

#include "flux_intersect.cuh"

#define BVH_EMPTY 0xFFFFFFFFu
#define BVH_LEAF 0x80000000u
#define BVH_DIRECT 0x40000000u   // leaf of ONE primitive, named in the reference itself: kind << 28 | index (28 bits) — no
                                 // trip through bvh_prims, one dependent load less per leaf visit
#define BVH_STACK 32          // entries per thread (shared memory); the builder keeps 3*depth+1 below it
#define BVH_PAD_REL 1e-7
#define BVH_PRUNE_REL 1e-9
// Inner nodes a lane may walk per descend() run before the warp moves on to its leaf tests (flux_bvh.cuh descend()).
// Unbounded is the classic while-while order; a small bound keeps the lanes that need 0-2 nodes to their next leaf
// from idling behind the few that have just started a ray at the root (~6 nodes).  Measured r2 (one B200):
// config 5, 20 M rays: unbounded 1841, 1 / 2 / 3 / 4 / 5 nodes 1690 / 2030 / 2175 / 2314 / 2245 Mrays/s;
// config 3 (regeneration kernel): unbounded 633, 1 / 2 / 3 / 4 / 5 nodes 639 / 675 / 670 / 659 / 653 Msamples/s.
#ifndef BVH_DESCEND_MAX
#define BVH_DESCEND_MAX 1u          // closest_hit_bvh: the render kernels (a warp's lanes start their segments together); 2 until the
                                    // triangle-only instantiation: with it 1 / 2 / 3 = 908 / 893 / 865 Msamples/s on config 3 (r2W)
#endif
#ifndef TRACE_DESCEND_MAX
#define TRACE_DESCEND_MAX 8u        // trace_rays_bvh_kernel: lanes refill one by one (4 while leaves held two primitives; with
                                    // leaves of one, r2F: 4 / 6 / 8 = 2496 / 2562 / 2600 Mrays/s)
#endif
#ifdef BVH_KEY_NAN_FIRST
#define BVH_KEY(tn) ((tn) == (tn) ? (tn) : -CUDART_INF_F)   // r1: an unknown entry bound sorts nearest
#else
#define BVH_KEY(tn) (tn)
#endif
#ifndef BVH_FIRST_LEAF
#define BVH_FIRST_LEAF 1      // leaf size tried first by the builder (doubled while the tree is too deep for the stack);
                              // measured r1: 2 vs 4 = +8 % on config 5, +1 % on config 3; r2, with bounded descend runs
                              // (r2E, one box): 1 / 2 / 4 = 2497 / 2192 / 1970 Mrays/s on config 5, 745 / 695 / 683
                              // Msamples/s on config 3 — a leaf visit is an f64 primitive test, a node visit four f32 ones
#endif

struct __align__(128) BvhNode4 {
    float lo[3][4];      // [axis][child], rounded down
    float hi[3][4];      // rounded up
    uint32_t child[4];   // EMPTY | inner node index | LEAF | offset << 3 | (count - 1)
    uint32_t pad[4];
};
static_assert(sizeof(BvhNode4) == 128, "BvhNode4 layout");

// leaf records
struct __align__(16) SphRec {   // 112 B
    double c0x, c1x, c0y, c1y, c0z, c1z;   // bounding box, shapes.rs:156-161
    double cx, cy, cz, rr;
    double r, inv;
    uint32_t shape_id, material, index, pad;
};
struct __align__(16) TriRec {   // 80 B... padded to 96
    double v0x, v0y, v0z, e1x, e1y, e1z, e2x, e2y, e2z;
    uint32_t shape_id, material;
    uint32_t index, pad[3];
};

#ifdef __CUDACC__
__device__ __forceinline__ double2 ldg2(const double *p) { return __ldg(reinterpret_cast<const double2 *>(p)); }

// Node fetches of the ray-batch kernel are split between the two data paths of the SM's L1: bit k of the mask sends
// fetch k (near x, far x, near y, far y, near z, far z, children; 16 bytes each) through the texture unit, the rest
// through the LSU.  ncu on config 5 (r2H): the kernel sat at 91 % of l1tex__data_pipe_lsu_wavefronts with issue slots
// 37 % busy — an incoherent 16-byte load costs the LSU data pipe one wavefront per LANE, seven of them per lane and node —
// and one more resident CTA per SM changed nothing.  Measured on one B200, 20 M rays: all seven through the LSU 2698
// Mrays/s, all through the texture unit 2890, split 3 + 4 (mask 7) 3419, 4 + 3 (mask 15) 3442, near x / far x / near y +
// children (71) 3452.  (Three 256-bit LDG.E.256 loads at fixed offsets with near / far picked in registers: 2914 — the
// requests fall by 36 % but the data-pipe wavefronts only by 9 %.)  The render kernels keep every fetch on the LSU
// (mask 0, the r1 code): config 3 with any mask 862 - 890 against 934 Msamples/s — their L1 hit rate is 51 %, an L1 hit
// is quick through the LSU and as slow as a miss through the texture pipe, and they are issue-bound as much as L1-bound
// (the children fetch alone through the texture unit: 871; the plane arrays picked by a 32-bit index instead of by
// pointer, all on the LSU: 869 — the compiler keeps six offsets live per ray and the kernel spills; the triangle records
// of the leaves through a texture of their own, all four 16-byte pieces or two of them: 879 / 882 / 870).  Pairs of predicated
// loads at immediate offsets (`@p ld [node+lo]; @!p ld [node+hi]`, no address arithmetic at all): config 5 3269 against
// 3462, config 3 630 against 934 — a load whose predicate is off is not free.
#ifndef TRACE_TEX_MASK
#define TRACE_TEX_MASK 71
#endif
#ifndef BVH_TEX_MASK
#define BVH_TEX_MASK 0
#endif
__device__ __forceinline__ float4 u4f(uint4 v) {
    return make_float4(__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z), __uint_as_float(v.w));
}

// Sphere::hit distance on a leaf record (same expressions as sphere_t in flux_intersect.cuh).
template <bool COUNT>
__device__ __forceinline__ bool sphere_t_rec(const RayCtx &r, const SphRec *__restrict__ s, double &t_out,
                                             unsigned long long *cn) {
    if (COUNT) cn[CN_BBOX_TESTS]++;
    const double2 bx = ldg2(&s->c0x), by = ldg2(&s->c0y), bz = ldg2(&s->c0z);
    if (!bbox_hit(r, bx.x, by.x, bz.x, bx.y, by.y, bz.y)) return false;
    if (COUNT) cn[CN_BBOX_PASS]++;
    const double2 c01 = ldg2(&s->cx), c23 = ldg2(&s->cz);
    V3 temp = r.o - mk3(c01.x, c01.y, c23.x);
    double b = 2.0 * dot3(temp, r.d);
    double cc = dot3(temp, temp) - c23.y;
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

__device__ __forceinline__ bool tri_t_rec(const RayCtx &r, const TriRec *__restrict__ q, double &t_out) {
    const double2 a = ldg2(&q->v0x), b = ldg2(&q->v0z), c = ldg2(&q->e1y), d = ldg2(&q->e2x);
    const double e2z = __ldg(&q->e2z);
    const V3 v0 = mk3(a.x, a.y, b.x), e1 = mk3(b.y, c.x, c.y), e2 = mk3(d.x, d.y, e2z);
    V3 p = cross3(r.d, e2);
    double det = dot3(e1, p);
    if (det == 0.0) return false;
    V3 s = r.o - v0;
    const double sp = dot3(s, p);
    // (r2, measured and dropped: rejecting from the signs and magnitudes of sp and det before the IEEE division — exact,
    // the division is ~25 instructions and most leaf triangles fail the u test — made config 3 slower, 681 against 695
    // Msamples/s: two more branches in a loop that already runs at 9 of 32 lanes cost more than the divisions saved)
    double inv = 1.0 / det;
    double u = sp * inv;
    if (!(u >= 0.0 && u <= 1.0)) return false;
    V3 qv = cross3(s, e1);
    double v = dot3(r.d, qv) * inv;
    if (!(v >= 0.0 && u + v <= 1.0)) return false;
    double t = dot3(e2, qv) * inv;
    if (!(t > FLUX_T_MIN)) return false;
    t_out = t;
    return true;
}

// Scene::hit through the BVH as an explicit traversal state, so that a kernel can interleave traversals with
// fetching new rays (trace_rays_kernel) as well as run one to completion (closest_hit_bvh).  `stack` points at this
// thread's column of a [BVH_STACK][stride] uint2 array (entry e at stack[e * stride]); shared or local memory.
//
// "while-while" order: descend() keeps a lane on inner nodes until it holds a leaf (or is done) before leaf() runs the
// f64 primitive tests, so the two very different code paths each execute with most of the warp (ncu on an if/else
// loop: 6.9 of 32 lanes per instruction on incoherent rays, profiles/r1s_bvh_kernels_full.txt).
// SPH = false: the tree holds no sphere (every sphere of the scene, if any, is on the linear list handled in begin()),
// so after begin() the ray's reciprocals and d.d terms are dead and the leaf code is the triangle test alone — fewer
// live registers in kernels that spill (the render kernels on mesh scenes).
template <bool COUNT, bool SPH = true>
struct BvhTraversal {
    RayCtx r;
    HitRef best;
    double t_prune, tscale;
    float t_prune32;
    float ia32[3], nlo[3], nhi[3];
    bool pos[3];
    uint32_t sp, cur;   // cur: inner node index | leaf reference | BVH_EMPTY (done; carries the leaf bit)

    __device__ __forceinline__ bool done() const { return cur == BVH_EMPTY; }

    __device__ __forceinline__ void candidate(double t, uint32_t id, uint32_t kind, uint32_t index, unsigned long long *cn) {
        if (COUNT) cn[CN_CANDIDATES]++;
        consider(best, t, id, kind, index);
        t_prune = best.t + BVH_PRUNE_REL * (fabs(best.t) + tscale);
        t_prune32 = __double2float_ru(t_prune);   // >= t_prune: what the f32 node test prunes against
    }

    // pop the nearest stacked subtree that can still hold a closer hit
    __device__ __forceinline__ void pop(const uint2 *stack, uint32_t stride) {
        cur = BVH_EMPTY;
        while (sp) {
            sp--;
            const uint2 e = stack[(size_t)sp * stride];
#ifdef BVH_POP_F64
            if ((double)__uint_as_float(e.y) > t_prune) continue;   // became prunable since it was pushed
#else
            // became prunable since it was pushed?  Against the f32 bound (>= t_prune: never prunes what f64 would keep):
            // no conversion to double on the XU pipe in a loop that runs at 3 - 5 lanes (9 % of config 3's issue slots)
            if (__uint_as_float(e.y) > t_prune32) continue;
#endif
            cur = e.x;
            break;
        }
    }

    // unbounded / oversized shapes first (they tighten t_prune early), then the f32 ray constants
    __device__ __forceinline__ void begin(const DevScene &sc, const RayCtx &ray, unsigned long long *cn) {
        r = ray;
        best.t = 0.0;
        best.shape_id = 0xFFFFFFFFu;
        best.kind = 0;
        best.index = 0;
        const double inf = __longlong_as_double(0x7FF0000000000000ll);
        // margin scale: one unit of coordinate error moves t by at most 1/|d| <= min_k |1/d_k|
        tscale = sc.bvh_extent * fmin(fabs(r.ia), fmin(fabs(r.ib), fabs(r.ic)));
        t_prune = inf;
        t_prune32 = CUDART_INF_F;   // >= t_prune, kept in step by candidate()
        for (uint32_t i = 0; i < sc.n_planes; i++) {
            double t;
            if (COUNT) cn[CN_PLANE_TESTS]++;
            if (plane_t(r, sc.pln, sc.n_planes, i, t)) candidate(t, __ldg(sc.pln_meta + i), KIND_PLANE, i, cn);
        }
        for (uint32_t k = 0; k < sc.bvh_n_linear; k++) {
            const uint32_t i = __ldg(sc.bvh_linear + k);
            double t;
            if (sphere_t<COUNT>(r, sc.sph, sc.n_spheres, i, t, cn)) candidate(t, __ldg(sc.sph_meta + i), KIND_SPHERE, i, cn);
        }
        sp = 0;
        cur = sc.bvh_n_nodes ? 0u : BVH_EMPTY;   // root
        // f32 ray constants: near = fma(c_near, ia, nlo) <= exact entry, far = fma(c_far, ia, nhi) >= exact exit
        const double dd[3] = {r.d.x, r.d.y, r.d.z}, oo[3] = {r.o.x, r.o.y, r.o.z};
        const float ext = __double2float_ru(sc.bvh_extent);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float df = (float)dd[k], of = (float)oo[k];
            float a;
            asm("rcp.approx.f32 %0, %1;" : "=f"(a) : "f"(df));   // max relative error 2^-23; 1/+-0 = +-inf
            if (!(fabsf(a) < 1e30f)) a = copysignf(1e30f, df);
            const float noa = -(of * a);
            // |f32 slab - exact slab| <= (2^-22 + 2^-23)(|c| + |o|)|ia| + 2^-24 |slab| <= E
            const float E = fabsf(a) * (ext + fabsf(of)) * (1.01f * 9.5367431640625e-07f);
            const bool sane = (df == df) && (fabsf(noa) < 1e37f) && (E < 1e37f);
            pos[k] = !signbit(a);                 // = (1/d_k >= 0) of BoundingBox::hit, shapes.rs:108,115,122
            ia32[k] = sane ? a : 0.0f;
            nlo[k] = sane ? noa - E : -CUDART_INF_F;   // an axis that cannot be bounded constrains nothing
            nhi[k] = sane ? noa + E : CUDART_INF_F;
        }
    }

    // inner nodes until this lane holds a leaf or has nothing left — or, with max_steps, for at most that many nodes:
    // the lane may then still hold an inner node (no leaf bit) and goes on in the caller's next round
    // TEXM: which of the seven 16-byte fetches of a node go through the texture unit (TRACE_TEX_MASK above); needs
    // sc.bvh_tex.  0 = all through the LSU.
    template <int TEXM = BVH_TEX_MASK>
    __device__ __forceinline__ void descend(const DevScene &sc, uint2 *stack, uint32_t stride, unsigned long long *cn,
                                            uint32_t max_steps = 0xFFFFFFFFu) {
        const BvhNode4 *__restrict__ nodes = reinterpret_cast<const BvhNode4 *>(sc.bvh_nodes);
        for (uint32_t step = 0; !(cur & BVH_LEAF) && step < max_steps; step++) {
            if (COUNT) cn[CN_NODES]++;
            const BvhNode4 *nd = nodes + cur;
            float4 ax, bx, ay, by, az, bz;
            uint4 ch;
            if (TEXM != 0) {
                const uint32_t tb = cur * 8u;   // eight uint4 texels per node; the plane arrays picked by index
                const cudaTextureObject_t tx = (cudaTextureObject_t)sc.bvh_tex;
                const uint4 *nq = reinterpret_cast<const uint4 *>(nd);
#define BVH_F4(slot, i) (((TEXM >> (slot)) & 1) ? tex1Dfetch<uint4>(tx, (int)(tb + (i))) : __ldg(nq + (i)))
                ax = u4f(BVH_F4(0, pos[0] ? 0u : 3u)); bx = u4f(BVH_F4(1, pos[0] ? 3u : 0u));
                ay = u4f(BVH_F4(2, pos[1] ? 1u : 4u)); by = u4f(BVH_F4(3, pos[1] ? 4u : 1u));
                az = u4f(BVH_F4(4, pos[2] ? 2u : 5u)); bz = u4f(BVH_F4(5, pos[2] ? 5u : 2u));
                ch = BVH_F4(6, 6u);
#undef BVH_F4
            } else {
                const float4 *lo4 = reinterpret_cast<const float4 *>(nd->lo), *hi4 = reinterpret_cast<const float4 *>(nd->hi);
                ax = __ldg(pos[0] ? lo4 + 0 : hi4 + 0); bx = __ldg(pos[0] ? hi4 + 0 : lo4 + 0);
                ay = __ldg(pos[1] ? lo4 + 1 : hi4 + 1); by = __ldg(pos[1] ? hi4 + 1 : lo4 + 1);
                az = __ldg(pos[2] ? lo4 + 2 : hi4 + 2); bz = __ldg(pos[2] ? hi4 + 2 : lo4 + 2);
                ch = __ldg(reinterpret_cast<const uint4 *>(nd->child));
            }
            float key[4];
            uint32_t ref[4];
            ref[0] = ch.x; ref[1] = ch.y; ref[2] = ch.z; ref[3] = ch.w;
#define BVH_CHILD(k, C)                                                                                             \
    {                                                                                                               \
        const float tn = fmaxf(fmaxf(fmaf(ax.C, ia32[0], nlo[0]), fmaf(ay.C, ia32[1], nlo[1])), fmaf(az.C, ia32[2], nlo[2])); \
        const float tf = fminf(fminf(fmaf(bx.C, ia32[0], nhi[0]), fmaf(by.C, ia32[1], nhi[1])), fmaf(bz.C, ia32[2], nhi[2])); \
        /* fmaxf / fminf drop NaN slabs (inf * 0): they constrain nothing; every comparison is false for NaN */    \
        const bool skip = (tn > tf) || (tf < 0.000499f) || (tn > t_prune32);                                       \
        /* unused slots hold the empty box (+inf, -inf): tn > tf skips them like any missed child.  A NaN bound (all  */ \
        /* three slabs NaN) sorts anywhere: a real child is still entered or pushed and never pruned, an unused slot  */ \
        /* keeps its BVH_EMPTY reference and is neither                                                               */ \
        if (skip) ref[k] = BVH_EMPTY;                                                                               \
        key[k] = skip ? CUDART_INF_F : BVH_KEY(tn);                                                                 \
    }
            BVH_CHILD(0, x) BVH_CHILD(1, y) BVH_CHILD(2, z) BVH_CHILD(3, w)
#undef BVH_CHILD
            // order: EMPTY last, then by entry bound (unknown = nearest); 5-comparator network
#define BVH_CSWAP(a, b)                                                         \
    if (key[b] < key[a]) {                                                      \
        const float tk = key[a]; key[a] = key[b]; key[b] = tk;                  \
        const uint32_t tr = ref[a]; ref[a] = ref[b]; ref[b] = tr;               \
    }
            BVH_CSWAP(0, 1) BVH_CSWAP(2, 3) BVH_CSWAP(0, 2) BVH_CSWAP(1, 3) BVH_CSWAP(1, 2)
#undef BVH_CSWAP
            // push far children first so that the nearest is popped first; continue with the nearest
#pragma unroll
            for (int k = 3; k >= 1; k--)
                if (ref[k] != BVH_EMPTY) {
                    stack[(size_t)sp * stride] = make_uint2(ref[k], __float_as_uint(key[k]));   // entry bound: re-tested when popped
                    sp++;
                }
            if (ref[0] != BVH_EMPTY) cur = ref[0];
            else pop(stack, stride);
        }
    }

    // the f64 primitive tests of the leaf this lane holds (the very functions of the linear scan), then the next subtree
    __device__ __forceinline__ void leaf(const DevScene &sc, const uint2 *stack, uint32_t stride, unsigned long long *cn) {
        const SphRec *__restrict__ srec = reinterpret_cast<const SphRec *>(sc.bvh_sph);
        const TriRec *__restrict__ trec = reinterpret_cast<const TriRec *>(sc.bvh_tri);
        const bool direct = (cur & BVH_DIRECT) != 0u;
        const uint32_t off = (cur & 0x3FFFFFFFu) >> 3, cnt = direct ? 1u : (cur & 7u) + 1u;
        for (uint32_t k = 0; k < cnt; k++) {
            const uint32_t pr = direct ? (((cur >> 28) & 3u) << 30) | (cur & 0x0FFFFFFFu) : __ldg(sc.bvh_prims + off + k);
            const uint32_t idx = pr & 0x3FFFFFFFu;
            double t;
            if (SPH && (pr >> 30) == KIND_SPHERE) {
                const SphRec *s = srec + idx;
                if (sphere_t_rec<COUNT>(r, s, t, cn)) candidate(t, __ldg(&s->shape_id), KIND_SPHERE, idx, cn);
            } else {
                const TriRec *q = trec + idx;
                if (COUNT) cn[CN_TRI_TESTS]++;
                if (tri_t_rec(r, q, t)) candidate(t, __ldg(&q->shape_id), KIND_TRI, idx, cn);
            }
        }
        pop(stack, stride);
    }
};

template <bool COUNT, bool SPH = true>
__device__ __forceinline__ HitRef closest_hit_bvh(const DevScene &sc, const RayCtx &r, uint2 *stack, uint32_t stride,
                                                  unsigned long long *cn) {
    BvhTraversal<COUNT, SPH> T;
    T.begin(sc, r, cn);
    while (!T.done()) {
        T.template descend<BVH_TEX_MASK>(sc, stack, stride, cn, BVH_DESCEND_MAX);
        if (T.done()) break;
        if (T.cur & BVH_LEAF) T.leaf(sc, stack, stride, cn);
    }
    return T.best;
}
#endif  // __CUDACC__

// ---- host builder (bvh_build.cu) ----
#include <string>
#include <vector>
#include "host_slices.h"
struct BvhBuild {
    flux_raw_vector<BvhNode4> nodes;   // (resize() does not zero these three: host_slices.h)
    flux_raw_vector<uint32_t> prims;   // (kind << 30) | index, in leaf order
    std::vector<uint32_t> linear;    // sphere indices kept out of the tree
    std::vector<SphRec> sph;         // [n_spheres], indexed like the SoA arrays
    flux_raw_vector<TriRec> tri;     // [n_triangles]
    double extent = 0.0;             // max |coordinate| over the boxes in the tree
    uint32_t depth = 0;              // levels of 4-wide nodes
    uint32_t leaf_size = 0;
};
// sph: SoA [SPH_FIELDS][ns] as uploaded; tri: SoA [TRI_FIELDS][nt]; v1/v2: original vertices (for the boxes)
bool build_bvh4(const double *sph, const uint32_t *sph_meta, uint32_t ns, const double *tri, const uint32_t *tri_meta,
                const double *tri_v1, const double *tri_v2, uint32_t nt, BvhBuild &out, std::string &err);
