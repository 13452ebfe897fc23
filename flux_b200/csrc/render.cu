// render.cu — the per-pixel render loop on sm_100a.
//
// Camera::render (fluxcore/src/trace.rs:53-97) → Scene::shade (scene.rs:162-172)
// → Material::path_shade (materials.rs:19-71), iterated instead of recursed:
// each material spawns exactly one child ray, so a path is a chain; the
// (f, weight) pair of every bounce is kept on a per-path stack and radiance is
// combined innermost-first exactly as the recursion returns it
// (SURVEY.md H4):  L_k = (f_k (*) L_{k+1}) * w_k.
//
// Work decomposition: a group of G = min(32, pow2ceil(spp)) lanes owns one
// pixel; lane g takes samples g, g+G, ...; the per-pixel sum is a fixed-shape
// xor-shuffle tree, so the result does not depend on grid size, scheduling or
// the number of GPUs.
#include "flux_bvh.cuh"
#include "flux_cull.cuh"
#include "flux_intersect.cuh"
#include "flux_kernels.cuh"
#include "flux_shade.cuh"

struct Rgb {
    double r, g, b;
};

// Body of the per-sample loop, trace.rs:71-80.
__device__ __forceinline__ void primary_ray(const DevCamera &cam, uint32_t row, uint32_t col, double2 ps, double2 ds,
                                            V3 &o, V3 &d) {
    double u = cam.aps * (((double)col - cam.half_w) + ps.x);
    double v = cam.aps * (((double)(cam.H - row) - cam.half_h) + ps.y);
    double lpx = ds.x * cam.lens_radius;
    double lpy = ds.y * cam.lens_radius;
    double px2 = u * cam.factor;
    double py2 = v * cam.factor;
    d = normalize3_dev(((px2 - lpx) * cam.u + (py2 - lpy) * cam.v) - cam.focal_w);   // = normalize3, one shared reciprocal (flux_math.cuh)
    o = (cam.eye + lpx * cam.u) + lpy * cam.v;
}

// scene.shade(&r, 1, ..) for one camera sample.
template <bool COUNT, bool BVH>
__device__ __forceinline__ Rgb trace_path(const RenderParams &p, V3 o, V3 d, uint32_t set, uint32_t i, double2 ps,
                                          uint2 *stack, unsigned long long *cn) {
    double sf[FLUX_MAX_DEPTH_CAP][4];  // (f.r, f.g, f.b, weight) per bounce
    uint32_t top = 0;
    Rgb L;
    uint32_t depth = 1;
    const DevScene &sc = p.scene;
    const bool culled = sc.n_tris == 0 && sc.n_spheres <= FLUX_CULL_MAX;   // uniform
    for (;;) {
        if (depth > p.cam.max_depth) {  // scene.rs:164-165
            if (COUNT) cn[CN_DEPTH_CUT]++;
            L = Rgb{0.0, 0.0, 0.0};
            break;
        }
        if (COUNT) cn[CN_SEGMENTS]++;
        RayCtx r;
        HitRef h;
        if (!BVH && culled) {   // spheres and planes only: FP32-classified boxes (flux_cull.cuh); the hit record needs o and d only
            r.o = o;
            r.d = d;
            h = closest_hit_linear_culled<COUNT>(p, o, d, cn);
        } else {
            r = make_ray(o, d);
            h = BVH ? closest_hit_bvh<COUNT>(sc, r, stack, blockDim.x, cn) : closest_hit_linear<COUNT>(sc, r, cn);
        }
        if (h.shape_id == 0xFFFFFFFFu) {  // scene.rs:168
            if (COUNT) cn[CN_MISS]++;
            L = Rgb{p.cam.bg[0], p.cam.bg[1], p.cam.bg[2]};
            break;
        }
        if (COUNT) cn[h.kind == KIND_SPHERE ? CN_HIT_SPHERE : (h.kind == KIND_PLANE ? CN_HIT_PLANE : CN_HIT_TRI)]++;
        HitRec hr = build_hit(sc, r, h);
        const DevMaterial *m = sc.materials + hr.material;
        uint32_t kind = m->kind;
        double c0 = m->c[0], c1 = m->c[1], c2 = m->c[2];
        if (kind == FLUX_MAT_EMISSIVE) {  // materials.rs:42-49
            if (COUNT) cn[CN_EMISSIVE]++;
            if (dot3(hr.normal * -1.0, d) > 0.0)
                L = Rgb{c0, c1, c2};
            else
                L = Rgb{0.0, 0.0, 0.0};
            break;
        }
        V3 wi;
        double weight;
        if (kind == FLUX_MAT_MATTE) {  // materials.rs:19-33
            if (COUNT) cn[CN_MATTE]++;
            const double *hp = p.ss.hemi + (((size_t)set * p.ss.max_depth + (depth - 1)) * p.ss.n + i) * 3;
            matte_sample(hr.normal, mk3(hp[0], hp[1], hp[2]), wi, weight);
        } else if (kind == FLUX_MAT_REFLECTIVE) {  // materials.rs:57-71, brdf.rs:39-45
            if (COUNT) cn[CN_SPECULAR]++;
            specular_sample(hr.normal, d, wi, weight);
        } else {  // glossy: materials.rs:57-71, brdf.rs:55-78
            if (COUNT) cn[CN_GLOSSY]++;
            double lobe;
            bool flipped;
            glossy_sample(hr.normal, d, ps.x, ps.y, m->exp, m->inv_e1, wi, weight, lobe, flipped);
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
        o = hr.point;
        d = wi;
        depth++;
    }
    while (top > 0) {  // unwind: (f (*) L) * w, materials.rs:31-32,69-70
        top--;
        L.r = (sf[top][0] * L.r) * sf[top][3];
        L.g = (sf[top][1] * L.g) * sf[top][3];
        L.b = (sf[top][2] * L.b) * sf[top][3];
    }
    return L;
}

#ifndef RENDER_MIN_BLOCKS
#define RENDER_MIN_BLOCKS 4   // linear-scan instantiations only (the BVH ones keep the compiler's own 113 registers).  Config 1,
                              // r2S/r2T: 1 / 2 / 3 / 4 blocks = 4425 / 4526 / 5060-5093 / 5423 Msamples/s (round 1 kernel: 4136)
#endif
template <bool COUNT, bool BVH>
__global__ void __launch_bounds__(256, BVH ? 1 : RENDER_MIN_BLOCKS) render_kernel(const __grid_constant__ RenderParams p, uint32_t G) {
    extern __shared__ uint2 bvh_stack[];  // [BVH_STACK][blockDim.x] when BVH
    uint2 *stack = BVH ? bvh_stack + threadIdx.x : nullptr;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t ppw = 32u / G;  // pixels per warp
    const uint32_t g = lane % G;
    const uint32_t W = p.cam.W;
    const uint32_t npix = p.n_rows * W;
    const uint32_t n_items = (npix + ppw - 1) / ppw;
    const uint32_t n = p.ss.n;
    unsigned long long cn[COUNT ? CN_COUNT : 1];
    if (COUNT)
        for (int k = 0; k < CN_COUNT; k++) cn[k] = 0;
    const double pixel_denom = 1.0 / (double)((unsigned long long)p.ss.root * p.ss.root);  // trace.rs:59

    for (;;) {
        uint32_t item = 0;
        if (lane == 0) item = atomicAdd(p.work_counter, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= n_items) break;
        const uint32_t pixel = item * ppw + lane / G;
        const bool valid = pixel < npix;
        Rgb acc = Rgb{0.0, 0.0, 0.0};
        uint32_t row = 0, col = 0;
        if (valid) {
            const uint32_t rk = pixel / W;
            col = pixel - rk * W;
            row = p.rows[rk];
            const uint32_t set = p.set_index[(size_t)row * W + col];
            const double2 *ps = p.ss.pixel + (size_t)set * n;
            const double2 *ds = p.ss.disc + (size_t)set * n;
            for (uint32_t i = p.i_begin + g; i < p.i_end; i += G) {
                double2 s = ps[i];
                double2 l = ds[i];
                V3 o, d;
                primary_ray(p.cam, row, col, s, l, o, d);
                if (COUNT) cn[CN_SAMPLES]++;
                Rgb c = trace_path<COUNT, BVH>(p, o, d, set, i, s, stack, cn);
                acc.r += c.r;
                acc.g += c.g;
                acc.b += c.b;
            }
        }
        for (uint32_t off = G >> 1; off > 0; off >>= 1) {
            acc.r += __shfl_xor_sync(0xffffffffu, acc.r, off);
            acc.g += __shfl_xor_sync(0xffffffffu, acc.g, off);
            acc.b += __shfl_xor_sync(0xffffffffu, acc.b, off);
        }
        if (valid && g == 0 && p.accum) {  // progressive pass: the sum of this pass joins the sums of the earlier ones
            double *a = p.accum + (size_t)pixel * 3;
            a[0] += acc.r;
            a[1] += acc.g;
            a[2] += acc.b;
        } else if (valid && g == 0) {  // trace.rs:85-86, color.rs:35-44
            double r = acc.r * pixel_denom, gg = acc.g * pixel_denom, b = acc.b * pixel_denom;
            double mx1 = r > gg ? r : gg;
            double mx2 = mx1 > b ? mx1 : b;
            if (mx2 > 1.0) {
                double inv = 1.0 / mx2;
                r *= inv;
                gg *= inv;
                b *= inv;
            }
            double *o = p.out + (p.out_by_row ? (size_t)row * W + col : (size_t)pixel) * 3;
            o[0] = r;
            o[1] = gg;
            o[2] = b;
        }
    }
    if (COUNT) {
        for (int k = 0; k < CN_COUNT; k++)
            if (cn[k]) atomicAdd(p.counters + k, cn[k]);
    }
}

// the image of the samples accumulated so far: sum * (1 / count), then max_to_one (trace.rs:85-86, color.rs:35-44)
__global__ void resolve_accum_kernel(const double *__restrict__ accum, double *__restrict__ out, uint32_t npix, double inv_count) {
    const uint32_t pixel = blockIdx.x * blockDim.x + threadIdx.x;
    if (pixel >= npix) return;
    double r = accum[(size_t)pixel * 3] * inv_count, gg = accum[(size_t)pixel * 3 + 1] * inv_count, b = accum[(size_t)pixel * 3 + 2] * inv_count;
    const double mx1 = r > gg ? r : gg;
    const double mx2 = mx1 > b ? mx1 : b;
    if (mx2 > 1.0) {
        const double inv = 1.0 / mx2;
        r *= inv;
        gg *= inv;
        b *= inv;
    }
    out[(size_t)pixel * 3] = r;
    out[(size_t)pixel * 3 + 1] = gg;
    out[(size_t)pixel * 3 + 2] = b;
}

void launch_resolve_accum(const double *accum, double *out, uint32_t npix, double inv_count, cudaStream_t stream) {
    if (npix) resolve_accum_kernel<<<(npix + 255) / 256, 256, 0, stream>>>(accum, out, npix, inv_count);
}

static uint32_t group_width(uint32_t n) {
    uint32_t G = 1;
    while (G < n && G < 32) G <<= 1;
    return G;
}

void launch_render(const RenderParams &p, bool count, int sm_count, cudaStream_t stream) {
    const uint32_t G = group_width(p.i_end - p.i_begin);
    const int threads = 256;
    const uint32_t ppw = 32 / G;
    const uint64_t npix = (uint64_t)p.n_rows * p.cam.W;
    const uint64_t n_items = (npix + ppw - 1) / ppw;
    uint64_t want = (n_items + (threads / 32) - 1) / (threads / 32);
    uint64_t cap = (uint64_t)sm_count * 4;
    int blocks = (int)(want < cap ? (want ? want : 1) : cap);
    const bool bvh = p.scene.use_bvh != 0;
    const size_t smem = bvh ? (size_t)BVH_STACK * threads * sizeof(uint2) : 0;
    auto go = [&](auto kern) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kern<<<blocks, threads, smem, stream>>>(p, G);
    };
    if (count) {
        if (bvh) go(render_kernel<true, true>);
        else go(render_kernel<true, false>);
    } else {
        if (bvh) go(render_kernel<false, true>);
        else go(render_kernel<false, false>);
    }
}

// ---- Scene::hit on an explicit ray batch (K6) -------------------------------
template <bool BVH>
__global__ void __launch_bounds__(128) trace_rays_kernel(const __grid_constant__ DevScene sc, uint64_t n,
                                                         const double *__restrict__ o, const double *__restrict__ d,
                                                         int32_t *__restrict__ hit, double *__restrict__ t) {
    extern __shared__ uint2 bvh_stack[];  // [BVH_STACK][blockDim.x] when BVH
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        RayCtx r = make_ray(mk3(o[3 * i], o[3 * i + 1], o[3 * i + 2]), mk3(d[3 * i], d[3 * i + 1], d[3 * i + 2]));
        HitRef h = BVH ? closest_hit_bvh<false>(sc, r, bvh_stack + threadIdx.x, blockDim.x, nullptr)
                       : closest_hit_linear<false>(sc, r, nullptr);
        if (h.shape_id == 0xFFFFFFFFu) {
            hit[i] = -1;
            t[i] = __longlong_as_double(0x7FF0000000000000ll);
        } else {
            hit[i] = (int32_t)h.shape_id;
            t[i] = h.t;
        }
    }
}

// BVH variant with per-lane ray refill (persistent "while-while" traversal in bounded runs): a lane whose ray is
// finished takes the next ray of its warp's current chunk at once (ballot rank, no atomics), so the node loop keeps
// running with most lanes instead of waiting for the warp's longest traversal — ncu on the ray-per-thread form showed
// 6.9 of 32 lanes per instruction (profiles/r1s_bvh_kernels_full.txt), 9.2 with the refill, 16.1 with bounded runs
// (profiles/r2h_bvh_kernels_full.txt), 18.8 with leaves of one primitive (profiles/r2H_bvh_kernels_full.txt).  Chunks of TRACE_CHUNK rays come from one global counter (one atomic per chunk),
// so a launch of any size ends with every warp busy until the rays run out: with r1's static split of the batch into
// equal shares per warp a 4 - 6 M-ray launch spent its last wave half empty.
#ifndef TRACE_CHUNK
#define TRACE_CHUNK 128u
#endif
#ifndef TRACE_MIN_BLOCKS
#define TRACE_MIN_BLOCKS 1
#endif
template <bool COUNT, int TEXM>
__global__ void __launch_bounds__(128, TRACE_MIN_BLOCKS) trace_rays_bvh_kernel(const __grid_constant__ DevScene sc, uint64_t n,
                                                             unsigned long long *__restrict__ work,
                                                             const double *__restrict__ o, const double *__restrict__ d,
                                                             int32_t *__restrict__ hit, double *__restrict__ t,
                                                             unsigned long long *__restrict__ counters) {
    extern __shared__ uint2 bvh_stack[];  // [BVH_STACK][blockDim.x]
    uint2 *stack = bvh_stack + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31u, lt_mask = (1u << lane) - 1u;
    uint64_t next = 0, end = 0;   // the warp's current chunk (uniform)
    bool exhausted = false;
    BvhTraversal<COUNT> T;
    unsigned long long cn[COUNT ? CN_COUNT : 1];
    if (COUNT)
        for (int k = 0; k < CN_COUNT; k++) cn[k] = 0;
    bool has = false;
    uint64_t mine = 0;
    for (;;) {
        const uint32_t need = __ballot_sync(0xffffffffu, !has);
        if (need != 0u && !exhausted) {
            if (next >= end) {   // uniform: fetch the next chunk of the batch
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(work, (unsigned long long)TRACE_CHUNK);
                base = __shfl_sync(0xffffffffu, base, 0);
                if (base >= n) {
                    exhausted = true;
                    next = end = n;
                } else {
                    next = base;
                    end = base + TRACE_CHUNK < n ? base + TRACE_CHUNK : n;
                }
            }
            if (!has) {
                const uint64_t i = next + __popc(need & lt_mask);
                if (i < end) {
                    if (COUNT) cn[CN_SEGMENTS]++;
                    T.begin(sc, make_ray(mk3(o[3 * i], o[3 * i + 1], o[3 * i + 2]), mk3(d[3 * i], d[3 * i + 1], d[3 * i + 2])), cn);
                    has = true;
                    mine = i;
                }
            }
            const uint64_t adv = next + __popc(need);
            next = adv < end ? adv : end;   // lanes left without a ray ask again next round (from the next chunk)
        }
        if (!__any_sync(0xffffffffu, has)) {
            if (exhausted) break;
            continue;
        }
        if (has) {
            T.template descend<TEXM>(sc, stack, blockDim.x, cn, TRACE_DESCEND_MAX);
            if (!T.done() && (T.cur & BVH_LEAF)) T.leaf(sc, stack, blockDim.x, cn);
            if (T.done()) {
                if (T.best.shape_id == 0xFFFFFFFFu) {
                    if (COUNT) cn[CN_MISS]++;
                    hit[mine] = -1;
                    t[mine] = __longlong_as_double(0x7FF0000000000000ll);
                } else {
                    hit[mine] = (int32_t)T.best.shape_id;
                    t[mine] = T.best.t;
                }
                has = false;
            }
        }
    }
    if (COUNT) {
        for (int k = 0; k < CN_COUNT; k++)
            if (cn[k]) atomicAdd(counters + k, cn[k]);
    }
}

#ifndef TRACE_CTAS_PER_SM
#define TRACE_CTAS_PER_SM 5
#endif
void launch_trace_rays(const DevScene &sc, uint64_t n, const double *o, const double *d, int32_t *hit, double *t,
                       int sm_count, cudaStream_t stream, unsigned long long *counters, unsigned long long *work) {
    if (n == 0) return;
    const int threads = 128;
    uint64_t want = (n + threads - 1) / threads;
    if (sc.use_bvh) {
        // persistent grid: as many CTAs as stay resident (5 per SM at 96 registers and 32 KB of traversal stacks), fed
        // from the chunk counter `work` (zeroed here, on the launching stream)
        const size_t smem = (size_t)BVH_STACK * threads * sizeof(uint2);
        const uint64_t cap = (uint64_t)sm_count * TRACE_CTAS_PER_SM;
        const int blocks = (int)(want < cap ? want : cap);
        cudaMemsetAsync(work, 0, sizeof(unsigned long long), stream);
        // with counters (flux_enable_counters): segments = rays, nodes_visited, bbox / triangle tests, candidates, misses
        // node fetches split between the texture unit and the LSU (flux_bvh.cuh TRACE_TEX_MASK) when the nodes have a
        // texture object; a tree too large for a linear texture goes through the LSU alone
        if (counters) trace_rays_bvh_kernel<true, 0><<<blocks, threads, smem, stream>>>(sc, n, work, o, d, hit, t, counters);
        else if (sc.bvh_tex != 0ull && TRACE_TEX_MASK != 0)
            trace_rays_bvh_kernel<false, TRACE_TEX_MASK><<<blocks, threads, smem, stream>>>(sc, n, work, o, d, hit, t, nullptr);
        else trace_rays_bvh_kernel<false, 0><<<blocks, threads, smem, stream>>>(sc, n, work, o, d, hit, t, nullptr);
    } else {
        uint64_t cap = (uint64_t)sm_count * 16;
        int blocks = (int)(want < cap ? want : cap);
        trace_rays_kernel<false><<<blocks, threads, 0, stream>>>(sc, n, o, d, hit, t);
    }
}
