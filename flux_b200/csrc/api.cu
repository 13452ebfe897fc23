// api.cu — the C-ABI of include/fluxb200.h: context, scene flattening to SoA
// device buffers, sample-set upload / generation, render and trace entry points.
//
// Host-side counterpart of LocalWorker's loop body (fluxcore/src/workers.rs:46-64):
// flux_set_scene = Scene::from_data + CameraBasis::new + Camera::new (minus the
// samples), flux_set_samples / flux_generate_samples = MasterSampleSets::new,
// flux_render_rows = Camera::render.
#include "../../include/fluxb200.h"
#include "flux_bvh.cuh"
#include "flux_kernels.cuh"
#include "host_slices.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <new>
#include <string>
#include <vector>

namespace {

std::string g_create_error;

template <class T> struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;  // elements
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc((void **)&p, std::max<size_t>(n, 1) * sizeof(T));
        if (e == cudaSuccess) cap = std::max<size_t>(n, 1);
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

}  // namespace

struct flux_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t s_in = nullptr, s_out = nullptr;   // copy streams of the pipelined host-buffer ray batches (flux_trace_rays)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string err;

    bool have_scene = false, have_samples = false, have_index = false;
    bool samples_in_range = true;   // pixel samples in [0,1]^2, lens samples in the unit disc (RenderParams::primary_mask_ok)
    // what the set-index map on the device was validated (flux_set_set_index) or generated for: it is usable only
    // while the image size and the number of sample sets are still these (index_valid)
    uint32_t idx_W = 0, idx_H = 0, idx_sets = 0;
    flux_job_config cfg{};
    DevCamera cam{};
    DevScene scene{};
    DevSamples ss{};
    int accel_mode = 0;
    int kernel_mode = 0;  // 0 auto, 1 direct (render.cu), 2 regeneration (render_regen.cu), 4 wavefront (render_wave2.cu)
    bool count = false;
    float last_ms = 0.f;
    bool ms_pending = false;   // last_ms is still to be read from ev0 / ev1 (device-pointer calls do not synchronise)
    uint64_t launches = 0;

    DevBuf<double> sph, pln, tri, hemi, out, ray_o, ray_d, ray_t, sink, ghemi, ginv, accum;
    // progressive passes: the rows being refined and how many samples per pixel the accumulator holds
    std::vector<uint32_t> prog_rows;
    uint32_t prog_done = 0;
    bool prog_active = false;
    std::vector<double> g_inv_e1;  // 1/(exp+1) of the distinct glossy exponents of the scene
    bool use_gtable = true;
    DevBuf<uint32_t> sph_meta, pln_meta, tri_meta, set_index, rows;
    DevBuf<int32_t> ray_hit;
    DevBuf<DevMaterial> materials;
    DevBuf<double2> pixel, disc;
    DevBuf<unsigned long long> counters;
    DevBuf<unsigned int> work_counter;
    DevBuf<unsigned long long> trace_work;   // chunk counter of the ray-batch kernel
    // BVH extension (flux_bvh.cuh)
    DevBuf<BvhNode4> bvh_nodes;
    cudaTextureObject_t bvh_tex = 0;   // bvh_nodes as a linear uint4 texture (DevScene::bvh_tex)
    DevBuf<SphRec> bvh_sph;
    DevBuf<TriRec> bvh_tri;
    DevBuf<uint32_t> bvh_prims, bvh_linear;
    uint32_t bvh_depth = 0, bvh_leaf = 0;
    float cull[FLUX_CULL_MAX][4];   // f32 spheres for render_wave2.cu (RenderParams::cull)
    float cull_cmax = 0.f;

    bool index_valid() const {
        return have_index && have_scene && have_samples && idx_W == cam.W && idx_H == cam.H && idx_sets == ss.num_sets;
    }
    // Called with the context's device current (flux_ctx_destroy, and the failure paths of flux_ctx_create).
    ~flux_ctx() {
        if (stream) cudaStreamSynchronize(stream);
        sph.release(); pln.release(); tri.release(); hemi.release(); out.release(); ray_o.release(); ray_d.release();
        ray_t.release(); sink.release(); ghemi.release(); ginv.release(); accum.release();
        sph_meta.release(); pln_meta.release(); tri_meta.release(); set_index.release(); rows.release();
        ray_hit.release(); materials.release(); pixel.release(); disc.release(); counters.release();
        work_counter.release(); trace_work.release(); bvh_nodes.release(); bvh_sph.release(); bvh_tri.release(); bvh_prims.release();
        bvh_linear.release();
        if (bvh_tex) cudaDestroyTextureObject(bvh_tex);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (s_in) cudaStreamDestroy(s_in);
        if (s_out) cudaStreamDestroy(s_out);
        if (stream) cudaStreamDestroy(stream);
        cudaGetLastError();
    }
};

namespace {

int fail(flux_ctx *c, int code, const std::string &msg) {
    if (c) c->err = msg;
    return code;
}

// No exception crosses the C ABI (std::bad_alloc from a host vector or a message string is the one that can arise): the
// entry points that allocate are function-try-blocks ending here.
int caught(flux_ctx *c, const char *where) noexcept {
    try {
        const std::string msg = std::string(where) + ": out of host memory (or another host-side exception)";
        if (c) c->err = msg;
        else g_create_error = msg;
    } catch (...) {
    }
    return FLUX_ERR_INVALID;
}

#define CK(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e__ = (call);                                                                      \
        if (e__ != cudaSuccess)                                                                        \
            return fail(ctx, FLUX_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));      \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// CameraBasis::new, fluxcore/src/scene.rs:29-34 (host, f64, same op order)
void camera_basis(V3 eye, V3 look_at, V3 up, V3 &u, V3 &v, V3 &w) {
    w = normalize3(eye - look_at);
    u = normalize3(cross3(up, w));
    v = cross3(w, u);
}

// spheres as SoA + bounding boxes as Sphere::new (shapes.rs:154-169)
void flatten_spheres(const flux_scene_flat *s, std::vector<double> &sph, std::vector<uint32_t> &sph_meta) {
    const uint32_t ns = s->n_spheres;
    sph.assign((size_t)SPH_FIELDS * ns, 0.0);
    sph_meta.assign((size_t)2 * ns, 0u);
    for (uint32_t i = 0; i < ns; i++) {
        const double *c = s->sphere_center + 3 * (size_t)i;
        const double r = s->sphere_radius[i];
        sph[(size_t)SPH_CX * ns + i] = c[0];
        sph[(size_t)SPH_CY * ns + i] = c[1];
        sph[(size_t)SPH_CZ * ns + i] = c[2];
        sph[(size_t)SPH_R * ns + i] = r;
        sph[(size_t)SPH_RR * ns + i] = r * r;                       // shapes.rs:179
        sph[(size_t)SPH_INV * ns + i] = s->sphere_invert[i] ? -1.0 : 1.0;  // shapes.rs:181
        sph[(size_t)SPH_C0X * ns + i] = c[0] - r;
        sph[(size_t)SPH_C0Y * ns + i] = c[1] - r;
        sph[(size_t)SPH_C0Z * ns + i] = c[2] - r;
        sph[(size_t)SPH_C1X * ns + i] = c[0] + r;
        sph[(size_t)SPH_C1Y * ns + i] = c[1] + r;
        sph[(size_t)SPH_C1Z * ns + i] = c[2] + r;
        sph_meta[i] = s->sphere_shape_id[i];
        sph_meta[ns + i] = s->sphere_material[i];
    }
}

// triangles (EXTENSION) as SoA: v0, e1 = v1 - v0, e2 = v2 - v0
void flatten_triangles(const flux_scene_flat *s, flux_raw_vector<double> &tri, flux_raw_vector<uint32_t> &tri_meta) {
    const uint32_t nt = s->n_triangles;
    tri.resize((size_t)TRI_FIELDS * nt);   // not zeroed (host_slices.h): every element is written below
    tri_meta.resize((size_t)2 * nt);
    flux_in_slices(nt, [&](uint32_t lo, uint32_t hi) {   // a million triangles: on several host threads
        for (uint32_t i = lo; i < hi; i++) {
            for (int k = 0; k < 3; k++) {
                const double v0 = s->tri_v0[3 * (size_t)i + k];
                tri[(size_t)(TRI_V0X + k) * nt + i] = v0;
                tri[(size_t)(TRI_E1X + k) * nt + i] = s->tri_v1[3 * (size_t)i + k] - v0;
                tri[(size_t)(TRI_E2X + k) * nt + i] = s->tri_v2[3 * (size_t)i + k] - v0;
            }
            tri_meta[i] = s->tri_shape_id[i];
            tri_meta[nt + i] = s->tri_material[i];
        }
    });
}

V3 ld3(const double *p) { return mk3(p[0], p[1], p[2]); }

}  // namespace

extern "C" {

const char *flux_version(void) { return "fluxb200 0.2.0 (sm_100a)"; }

const char *flux_last_error(const flux_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int flux_ctx_create(int device, flux_ctx **out) {
    if (!out) {
        g_create_error = "flux_ctx_create: out is null";
        return FLUX_ERR_INVALID;
    }
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        g_create_error = std::string("flux_ctx_create: no CUDA device (") +
                         (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0") +
                         "); libfluxb200 has no CPU fallback";
        return FLUX_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= n) {
        g_create_error = "flux_ctx_create: device " + std::to_string(device) + " out of range (" + std::to_string(n) + " devices)";
        return FLUX_ERR_NO_DEVICE;
    }
    flux_ctx *ctx = new (std::nothrow) flux_ctx();
    if (!ctx) {
        g_create_error = "flux_ctx_create: out of memory";
        return FLUX_ERR_INVALID;
    }
    ctx->device = device;
    DeviceGuard g(device);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaEventCreate(&ctx->ev0)) != cudaSuccess || (e = cudaEventCreate(&ctx->ev1)) != cudaSuccess) {
        g_create_error = std::string("flux_ctx_create: ") + cudaGetErrorString(e);
        delete ctx;
        return FLUX_ERR_CUDA;
    }
    if (prop.major < 10) {
        g_create_error = "flux_ctx_create: device is sm_" + std::to_string(prop.major * 10 + prop.minor) +
                         "; libfluxb200 is built for sm_100a only";
        delete ctx;
        return FLUX_ERR_NO_DEVICE;
    }
    ctx->sm_count = prop.multiProcessorCount;
    if (ctx->counters.reserve(CN_COUNT) != cudaSuccess || ctx->work_counter.reserve(1) != cudaSuccess || ctx->trace_work.reserve(1) != cudaSuccess ||
        ctx->sink.reserve(1) != cudaSuccess) {
        g_create_error = "flux_ctx_create: cudaMalloc failed";
        delete ctx;
        return FLUX_ERR_CUDA;
    }
    cudaMemsetAsync(ctx->counters.p, 0, CN_COUNT * sizeof(unsigned long long), ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    *out = ctx;
    return FLUX_OK;
}

int flux_ctx_destroy(flux_ctx *ctx) {
    if (!ctx) return FLUX_ERR_INVALID;
    DeviceGuard g(ctx->device);
    delete ctx;   // ~flux_ctx releases every device buffer (the progressive accumulator included), events and the stream
    return FLUX_OK;
}

static int build_glossy_table(flux_ctx *ctx);

int flux_set_scene(flux_ctx *ctx, const flux_scene_flat *s, const flux_job_config *cfg) try {
    if (!ctx) return FLUX_ERR_INVALID;
    if (!s || !cfg) return fail(ctx, FLUX_ERR_INVALID, "flux_set_scene: null scene or config");
    if (s->image_width == 0 || s->image_height == 0) return fail(ctx, FLUX_ERR_INVALID, "flux_set_scene: empty image");
    if (cfg->sample_root == 0) return fail(ctx, FLUX_ERR_INVALID, "flux_set_scene: sample_root must be >= 1");
    if (cfg->max_trace_depth > FLUX_MAX_DEPTH_CAP)
        return fail(ctx, FLUX_ERR_INVALID, "flux_set_scene: max_trace_depth exceeds " + std::to_string(FLUX_MAX_DEPTH_CAP));
    if ((s->n_materials && !s->materials) || (s->n_spheres && (!s->sphere_center || !s->sphere_radius || !s->sphere_invert || !s->sphere_shape_id || !s->sphere_material)) ||
        (s->n_planes && (!s->plane_point || !s->plane_normal || !s->plane_shape_id || !s->plane_material)) ||
        (s->n_triangles && (!s->tri_v0 || !s->tri_v1 || !s->tri_v2 || !s->tri_shape_id || !s->tri_material)))
        return fail(ctx, FLUX_ERR_INVALID, "flux_set_scene: null array with non-zero count");
    auto check_ids = [&](const uint32_t *ids, const uint32_t *mats, uint32_t n, const char *what) -> bool {
        for (uint32_t i = 0; i < n; i++) {
            if (mats[i] >= s->n_materials) {
                ctx->err = std::string("flux_set_scene: ") + what + " material index out of range";
                return false;
            }
            if (i && ids[i] <= ids[i - 1]) {
                ctx->err = std::string("flux_set_scene: ") + what + " shape ids must be strictly increasing";
                return false;
            }
        }
        return true;
    };
    if (!check_ids(s->sphere_shape_id, s->sphere_material, s->n_spheres, "sphere") ||
        !check_ids(s->plane_shape_id, s->plane_material, s->n_planes, "plane") ||
        !check_ids(s->tri_shape_id, s->tri_material, s->n_triangles, "triangle"))
        return FLUX_ERR_INVALID;
    for (uint32_t i = 0; i < s->n_materials; i++)
        if (s->materials[i].kind > FLUX_MAT_GLOSSY) return fail(ctx, FLUX_ERR_INVALID, "flux_set_scene: unknown material kind");
    // Triangles are this library's extension and their vertices must be finite — whichever way the closest hit is then
    // found, so that adding an unrelated shape (and with it the BVH, whose builder has no box for such a triangle)
    // never turns a scene that rendered into an error.  Spheres and planes keep the reference's semantics: anything goes.
    {
        std::atomic<uint32_t> first_bad{0xFFFFFFFFu};   // the lowest index, whichever host thread meets it
        flux_in_slices(s->n_triangles, [&](uint32_t lo, uint32_t hi) {
            for (uint32_t i = lo; i < hi; i++) {
                bool finite = true;
                for (int k = 0; k < 3; k++)
                    finite = finite && std::isfinite(s->tri_v0[3 * (size_t)i + k]) && std::isfinite(s->tri_v1[3 * (size_t)i + k]) &&
                             std::isfinite(s->tri_v2[3 * (size_t)i + k]);
                if (!finite) {
                    uint32_t seen = first_bad.load();
                    while (i < seen && !first_bad.compare_exchange_weak(seen, i)) {}
                    return;   // later triangles of this slice have higher indices
                }
            }
        });
        if (first_bad.load() != 0xFFFFFFFFu)
            return fail(ctx, FLUX_ERR_INVALID, "flux_set_scene: triangle " + std::to_string(first_bad.load()) + " has a non-finite vertex");
    }

    DeviceGuard g(ctx->device);
    ctx->have_scene = false;
    ctx->prog_active = false;
    // ---- materials: per-material constants (same IEEE products as the reference) ----
    std::vector<DevMaterial> mats(s->n_materials);
    for (uint32_t i = 0; i < s->n_materials; i++) {
        const flux_material &m = s->materials[i];
        DevMaterial d{};
        d.kind = m.kind;
        for (int k = 0; k < 3; k++) {
            double ck = m.color[k] * m.k;                      // cd*kd | color*power | cr*kr | cs*ks
            d.c[k] = (m.kind == FLUX_MAT_MATTE) ? ck * FLUX_INV_PI : ck;  // brdf.rs:29
        }
        d.exp = m.exp;
        d.inv_e1 = 1.0 / (m.exp + 1.0);  // samplers/src/lib.rs:135
        d.gidx = 0;
        mats[i] = d;
    }
    // distinct glossy exponents -> columns of the lobe table (DevSamples::ghemi)
    ctx->g_inv_e1.clear();
    {
        std::vector<double> exps;
        for (uint32_t i = 0; i < s->n_materials; i++) {
            if (mats[i].kind != FLUX_MAT_GLOSSY) continue;
            size_t k = 0;
            while (k < exps.size() && std::memcmp(&exps[k], &mats[i].exp, sizeof(double)) != 0) k++;
            if (k == exps.size()) {
                exps.push_back(mats[i].exp);
                ctx->g_inv_e1.push_back(mats[i].inv_e1);
            }
            mats[i].gidx = (uint32_t)k;
        }
    }
    // ---- spheres: SoA + bounding boxes as Sphere::new (shapes.rs:154-169) ----
    const uint32_t ns = s->n_spheres, np = s->n_planes, nt = s->n_triangles;
    std::vector<double> sph;
    std::vector<uint32_t> sph_meta;
    flatten_spheres(s, sph, sph_meta);
    std::vector<double> pln((size_t)PLN_FIELDS * np);
    std::vector<uint32_t> pln_meta((size_t)2 * np);
    for (uint32_t i = 0; i < np; i++) {
        for (int k = 0; k < 3; k++) {
            pln[(size_t)(PLN_PX + k) * np + i] = s->plane_point[3 * i + k];
            pln[(size_t)(PLN_NX + k) * np + i] = s->plane_normal[3 * i + k];
        }
        pln_meta[i] = s->plane_shape_id[i];
        pln_meta[np + i] = s->plane_material[i];
    }
    flux_raw_vector<double> tri;
    flux_raw_vector<uint32_t> tri_meta;
    flatten_triangles(s, tri, tri_meta);
    CK(ctx->materials.reserve(mats.size()));
    CK(ctx->sph.reserve(sph.size()));
    CK(ctx->sph_meta.reserve(sph_meta.size()));
    CK(ctx->pln.reserve(pln.size()));
    CK(ctx->pln_meta.reserve(pln_meta.size()));
    CK(ctx->tri.reserve(tri.size()));
    CK(ctx->tri_meta.reserve(tri_meta.size()));
    auto up = [&](void *dst, const void *src, size_t bytes) {
        return bytes ? cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream) : cudaSuccess;
    };
    CK(up(ctx->materials.p, mats.data(), mats.size() * sizeof(DevMaterial)));
    CK(up(ctx->sph.p, sph.data(), sph.size() * sizeof(double)));
    CK(up(ctx->sph_meta.p, sph_meta.data(), sph_meta.size() * sizeof(uint32_t)));
    CK(up(ctx->pln.p, pln.data(), pln.size() * sizeof(double)));
    CK(up(ctx->pln_meta.p, pln_meta.data(), pln_meta.size() * sizeof(uint32_t)));
    CK(up(ctx->tri.p, tri.data(), tri.size() * sizeof(double)));
    CK(up(ctx->tri_meta.p, tri_meta.data(), tri_meta.size() * sizeof(uint32_t)));
    CK(cudaStreamSynchronize(ctx->stream));  // host vectors go out of scope

    DevScene &sc = ctx->scene;
    sc = DevScene{};
    sc.n_spheres = ns;
    sc.n_planes = np;
    sc.n_tris = nt;
    sc.n_materials = s->n_materials;
    sc.sph = ctx->sph.p;
    sc.sph_meta = ctx->sph_meta.p;
    sc.pln = ctx->pln.p;
    sc.pln_meta = ctx->pln_meta.p;
    sc.tri = ctx->tri.p;
    sc.tri_meta = ctx->tri_meta.p;
    sc.materials = ctx->materials.p;

    // ---- f32 spheres for the conservative slab pre-test (render_wave2.cu) ----
    {
        float cmax = 0.f;
        const float up = std::numeric_limits<float>::infinity();
        for (uint32_t i = 0; i < FLUX_CULL_MAX; i++) {
            for (int k = 0; k < 4; k++) ctx->cull[i][k] = std::numeric_limits<float>::quiet_NaN();
            if (i >= ns) continue;
            const double c[3] = {sph[(size_t)SPH_CX * ns + i], sph[(size_t)SPH_CY * ns + i], sph[(size_t)SPH_CZ * ns + i]};
            const double r = sph[(size_t)SPH_R * ns + i];
            const bool ok = std::isfinite(c[0]) && std::isfinite(c[1]) && std::isfinite(c[2]) && std::isfinite(r) && r >= 0.0 &&
                            std::fabs(c[0]) < 1e30 && std::fabs(c[1]) < 1e30 && std::fabs(c[2]) < 1e30 && r < 1e30;
            if (!ok) continue;   // stays NaN: always the exact test
            for (int k = 0; k < 3; k++) {
                ctx->cull[i][k] = (float)c[k];
                cmax = std::max(cmax, std::nextafter((float)(std::fabs(c[k]) + r), up));
            }
            ctx->cull[i][3] = (float)r;
        }
        ctx->cull_cmax = std::nextafter(cmax, up);
    }

    // ---- acceleration structure (EXTENSION): the reference scans linearly (scene.rs:156-160); scenes beyond
    // FLUX_LINEAR_LIMIT bounded shapes get a BVH that returns the same answer (flux_bvh.cuh) ----
    ctx->bvh_depth = ctx->bvh_leaf = 0;
    const bool want_bvh = (ns + nt > 0) && (ctx->accel_mode == 2 || (ctx->accel_mode == 0 && (uint64_t)ns + nt > FLUX_LINEAR_LIMIT));
    if (want_bvh) {
        BvhBuild bb;
        std::string berr;
        if (!build_bvh4(sph.data(), sph_meta.data(), ns, tri.data(), tri_meta.data(), s->tri_v1, s->tri_v2, nt, bb, berr))
            return fail(ctx, FLUX_ERR_INVALID, "flux_set_scene: " + berr);
        CK(ctx->bvh_nodes.reserve(bb.nodes.size()));
        CK(ctx->bvh_sph.reserve(bb.sph.size()));
        CK(ctx->bvh_tri.reserve(bb.tri.size()));
        CK(ctx->bvh_prims.reserve(bb.prims.size()));
        CK(ctx->bvh_linear.reserve(bb.linear.size()));
        CK(up(ctx->bvh_nodes.p, bb.nodes.data(), bb.nodes.size() * sizeof(BvhNode4)));
        CK(up(ctx->bvh_sph.p, bb.sph.data(), bb.sph.size() * sizeof(SphRec)));
        CK(up(ctx->bvh_tri.p, bb.tri.data(), bb.tri.size() * sizeof(TriRec)));
        CK(up(ctx->bvh_prims.p, bb.prims.data(), bb.prims.size() * sizeof(uint32_t)));
        CK(up(ctx->bvh_linear.p, bb.linear.data(), bb.linear.size() * sizeof(uint32_t)));
        CK(cudaStreamSynchronize(ctx->stream));
        if (ctx->bvh_tex) {
            cudaDestroyTextureObject(ctx->bvh_tex);
            ctx->bvh_tex = 0;
        }
        // the nodes once more as a linear texture, eight uint4 texels per node (flux_bvh.cuh TRACE_TEX_MASK).  A tree too
        // large for a linear texture — or FLUXB200_NO_NODE_TEXTURE in the environment, a test hook — leaves bvh_tex 0 and
        // the ray-batch kernel on its LSU-only instantiation
        if (!std::getenv("FLUXB200_NO_NODE_TEXTURE")) {
            cudaResourceDesc rd{};
            rd.resType = cudaResourceTypeLinear;
            rd.res.linear.devPtr = ctx->bvh_nodes.p;
            rd.res.linear.desc = cudaCreateChannelDesc(32, 32, 32, 32, cudaChannelFormatKindUnsigned);
            rd.res.linear.sizeInBytes = bb.nodes.size() * sizeof(BvhNode4);
            cudaTextureDesc td{};
            td.readMode = cudaReadModeElementType;
            if (cudaCreateTextureObject(&ctx->bvh_tex, &rd, &td, nullptr) != cudaSuccess) {
                ctx->bvh_tex = 0;
                cudaGetLastError();
            }
        }
        sc.use_bvh = 1;
        sc.bvh_nodes = ctx->bvh_nodes.p;
        sc.bvh_tex = (unsigned long long)ctx->bvh_tex;
        sc.bvh_n_nodes = (uint32_t)bb.nodes.size();
        sc.bvh_prims = ctx->bvh_prims.p;
        sc.bvh_sph = ctx->bvh_sph.p;
        sc.bvh_tri = ctx->bvh_tri.p;
        sc.bvh_linear = ctx->bvh_linear.p;
        sc.bvh_n_linear = (uint32_t)bb.linear.size();
        sc.bvh_tree_spheres = ns - (uint32_t)bb.linear.size();
        sc.bvh_extent = bb.extent;
        ctx->bvh_depth = bb.depth;
        ctx->bvh_leaf = bb.leaf_size;
    }

    // ---- camera: CameraBasis::new (scene.rs:29-34) + render() prologue (trace.rs:54-60) ----
    DevCamera &cam = ctx->cam;
    cam = DevCamera{};
    cam.eye = ld3(s->eye);
    camera_basis(cam.eye, ld3(s->look_at), ld3(s->up), cam.u, cam.v, cam.w);
    cam.aps = s->pixel_size / s->zoom_factor;
    cam.half_w = (double)s->image_width * 0.5;
    cam.half_h = (double)s->image_height * 0.5;
    cam.factor = s->focal_distance / s->view_plane_distance;
    cam.focal = s->focal_distance;
    cam.lens_radius = s->lens_radius;
    cam.focal_w = s->focal_distance * cam.w;
    for (int k = 0; k < 3; k++) cam.bg[k] = s->background[k];
    cam.W = s->image_width;
    cam.H = s->image_height;
    cam.max_depth = cfg->max_trace_depth;
    ctx->cfg = *cfg;
    ctx->have_scene = true;
    // samples / set index belong to a job: a new scene invalidates them if shapes changed
    if (ctx->have_samples && (ctx->ss.root != cfg->sample_root || ctx->ss.max_depth != cfg->max_trace_depth))
        ctx->have_samples = false;
    // the set-index map is [H][W] of set numbers: another image size (even one that fits the old allocation) or
    // fewer sample sets make it stale
    if (ctx->have_index && (ctx->idx_W != cam.W || ctx->idx_H != cam.H || !ctx->have_samples)) ctx->have_index = false;
    if (ctx->have_samples) return build_glossy_table(ctx);  // exponents may have changed
    ctx->ss.ghemi = nullptr;
    ctx->ss.gk = 0;
    return FLUX_OK;
} catch (...) {
    return caught(ctx, "flux_set_scene");
}

// (Re)build the glossy lobe table for the current scene and sample sets.  Skipped (kernels fall back to
// inline evaluation, same values) when the scene has no glossy material or the table would exceed 16 GiB.
static int build_glossy_table(flux_ctx *ctx) {
    ctx->ss.ghemi = nullptr;
    ctx->ss.gk = 0;
    const size_t gk = ctx->g_inv_e1.size();
    if (!ctx->use_gtable || gk == 0 || !ctx->have_scene) return FLUX_OK;
    const size_t elems = (size_t)ctx->ss.num_sets * gk * ctx->ss.n * 3;
    if (elems * sizeof(double) > (16ull << 30)) return FLUX_OK;
    CK(ctx->ghemi.reserve(elems));
    CK(ctx->ginv.reserve(gk));
    CK(cudaMemcpyAsync(ctx->ginv.p, ctx->g_inv_e1.data(), gk * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    launch_build_glossy_table(ctx->pixel.p, ctx->ss.n, ctx->ss.num_sets, (uint32_t)gk, ctx->ginv.p, ctx->ghemi.p,
                              ctx->sm_count, ctx->stream);
    ctx->launches += 1;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->ss.ghemi = ctx->ghemi.p;
    ctx->ss.gk = (uint32_t)gk;
    return FLUX_OK;
}

static int alloc_samples(flux_ctx *ctx, uint32_t root, uint32_t max_depth, uint32_t num_sets) {
    const size_t n = (size_t)root * root;
    CK(ctx->pixel.reserve(n * num_sets));
    CK(ctx->disc.reserve(n * num_sets));
    CK(ctx->hemi.reserve(n * num_sets * max_depth * 3));
    ctx->ss.root = root;
    ctx->ss.n = (uint32_t)n;
    ctx->ss.max_depth = max_depth;
    ctx->ss.num_sets = num_sets;
    ctx->ss.pixel = ctx->pixel.p;
    ctx->ss.disc = ctx->disc.p;
    ctx->ss.hemi = ctx->hemi.p;
    return FLUX_OK;
}

int flux_set_samples(flux_ctx *ctx, uint32_t root, uint32_t max_depth, uint32_t num_sets, const double *pixel_xy,
                     const double *disc_xy, const double *hemi_xyz) try {
    if (!ctx) return FLUX_ERR_INVALID;
    if (!ctx->have_scene) return fail(ctx, FLUX_ERR_STATE, "flux_set_samples: call flux_set_scene first");
    if (root == 0 || num_sets == 0 || !pixel_xy || !disc_xy || (max_depth && !hemi_xyz))
        return fail(ctx, FLUX_ERR_INVALID, "flux_set_samples: bad arguments");
    if (root != ctx->cfg.sample_root || max_depth != ctx->cfg.max_trace_depth)
        return fail(ctx, FLUX_ERR_INVALID, "flux_set_samples: root/max_depth differ from the job configuration");
    if ((uint64_t)root * root > 0xFFFFFFFFull) return fail(ctx, FLUX_ERR_INVALID, "flux_set_samples: root too large");
    DeviceGuard g(ctx->device);
    ctx->have_samples = false;
    ctx->prog_active = false;
    // a map validated against another number of sets (or another image) may name sets that no longer exist
    if (ctx->have_index && (ctx->idx_sets != num_sets || ctx->idx_W != ctx->cam.W || ctx->idx_H != ctx->cam.H)) ctx->have_index = false;
    int rc = alloc_samples(ctx, root, max_depth, num_sets);
    if (rc) return rc;
    const size_t n = (size_t)root * root;
    {   // the reference's sets are unit-square / unit-disc samples (samplers/src/lib.rs:46-131); a caller may hand over
        // anything, and the wavefront kernel's per-pixel primary mask must then stand aside
        bool ok = true;
        const size_t m = n * num_sets;
        for (size_t k = 0; k < m && ok; k++) {
            const double px = pixel_xy[2 * k], py = pixel_xy[2 * k + 1], dx = disc_xy[2 * k], dy = disc_xy[2 * k + 1];
            ok = px >= 0.0 && px <= 1.0 && py >= 0.0 && py <= 1.0 && dx * dx + dy * dy <= 1.0 + 1e-12;
        }
        ctx->samples_in_range = ok;
    }
    CK(cudaMemcpyAsync(ctx->pixel.p, pixel_xy, n * num_sets * 16, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->disc.p, disc_xy, n * num_sets * 16, cudaMemcpyHostToDevice, ctx->stream));
    if (max_depth) CK(cudaMemcpyAsync(ctx->hemi.p, hemi_xyz, n * num_sets * max_depth * 24, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->have_samples = true;
    return build_glossy_table(ctx);
} catch (...) {
    return caught(ctx, "flux_set_samples");
}

int flux_generate_samples(flux_ctx *ctx, uint64_t seed, uint32_t num_sets) try {
    if (!ctx) return FLUX_ERR_INVALID;
    if (!ctx->have_scene) return fail(ctx, FLUX_ERR_STATE, "flux_generate_samples: call flux_set_scene first");
    if (num_sets == 0) return fail(ctx, FLUX_ERR_INVALID, "flux_generate_samples: num_sets must be >= 1");
    const uint32_t root = ctx->cfg.sample_root, depth = ctx->cfg.max_trace_depth;
    if ((size_t)4 * root * root > 200 * 1024)
        return fail(ctx, FLUX_ERR_INVALID, "flux_generate_samples: sample_root too large for on-device permutations (use flux_set_samples)");
    if ((size_t)num_sets * 4 > 200 * 1024)
        return fail(ctx, FLUX_ERR_INVALID, "flux_generate_samples: num_sets too large (use flux_set_set_index)");
    DeviceGuard g(ctx->device);
    ctx->have_samples = false;
    ctx->prog_active = false;
    ctx->have_index = false;
    int rc = alloc_samples(ctx, root, depth, num_sets);
    if (rc) return rc;
    CK(ctx->set_index.reserve((size_t)ctx->cam.W * ctx->cam.H));
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    ctx->samples_in_range = true;   // by construction (samplegen.cu)
    launch_generate_samples(seed, root, depth, num_sets, ctx->pixel.p, ctx->disc.p, ctx->hemi.p, ctx->stream);
    launch_generate_set_index(seed, ctx->cam.H, ctx->cam.W, num_sets, ctx->set_index.p, ctx->stream);
    ctx->launches += 2;
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));
    ctx->have_samples = true;
    ctx->have_index = true;
    ctx->idx_W = ctx->cam.W; ctx->idx_H = ctx->cam.H; ctx->idx_sets = num_sets;
    return build_glossy_table(ctx);
} catch (...) {
    return caught(ctx, "flux_generate_samples");
}

int flux_get_samples(flux_ctx *ctx, double *pixel_xy, double *disc_xy, double *hemi_xyz) {
    if (!ctx) return FLUX_ERR_INVALID;
    if (!ctx->have_samples) return fail(ctx, FLUX_ERR_STATE, "flux_get_samples: no samples on the device");
    DeviceGuard g(ctx->device);
    const size_t n = (size_t)ctx->ss.n * ctx->ss.num_sets;
    if (pixel_xy) CK(cudaMemcpyAsync(pixel_xy, ctx->pixel.p, n * 16, cudaMemcpyDeviceToHost, ctx->stream));
    if (disc_xy) CK(cudaMemcpyAsync(disc_xy, ctx->disc.p, n * 16, cudaMemcpyDeviceToHost, ctx->stream));
    if (hemi_xyz && ctx->ss.max_depth)
        CK(cudaMemcpyAsync(hemi_xyz, ctx->hemi.p, n * ctx->ss.max_depth * 24, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return FLUX_OK;
}

int flux_get_set_index(flux_ctx *ctx, uint32_t *idx) {
    if (!ctx) return FLUX_ERR_INVALID;
    if (!ctx->index_valid()) return fail(ctx, FLUX_ERR_STATE, "flux_get_set_index: no set-index map on the device");
    if (!idx) return fail(ctx, FLUX_ERR_INVALID, "flux_get_set_index: null output");
    DeviceGuard g(ctx->device);
    CK(cudaMemcpyAsync(idx, ctx->set_index.p, (size_t)ctx->cam.W * ctx->cam.H * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return FLUX_OK;
}

int flux_set_set_index(flux_ctx *ctx, const uint32_t *idx) try {
    if (!ctx) return FLUX_ERR_INVALID;
    if (!ctx->have_scene) return fail(ctx, FLUX_ERR_STATE, "flux_set_set_index: call flux_set_scene first");
    if (!ctx->have_samples) return fail(ctx, FLUX_ERR_STATE, "flux_set_set_index: set or generate samples first");
    if (!idx) return fail(ctx, FLUX_ERR_INVALID, "flux_set_set_index: null map");
    const size_t n = (size_t)ctx->cam.W * ctx->cam.H;
    for (size_t i = 0; i < n; i++)
        if (idx[i] >= ctx->ss.num_sets) return fail(ctx, FLUX_ERR_INVALID, "flux_set_set_index: set index out of range");
    DeviceGuard g(ctx->device);
    ctx->have_index = false;
    CK(ctx->set_index.reserve(n));
    CK(cudaMemcpyAsync(ctx->set_index.p, idx, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->have_index = true;
    ctx->idx_W = ctx->cam.W; ctx->idx_H = ctx->cam.H; ctx->idx_sets = ctx->ss.num_sets;
    return FLUX_OK;
} catch (...) {
    return caught(ctx, "flux_set_set_index");
}

int flux_shard_rows(uint32_t image_height, uint32_t tile_rows, uint32_t rank, uint32_t world, uint32_t *rows,
                    uint32_t *n_rows) {
    if (!n_rows || tile_rows == 0 || world == 0 || rank >= world) return FLUX_ERR_INVALID;
    uint32_t k = 0;
    for (uint32_t r = 0; r < image_height; r++)
        if ((r / tile_rows) % world == rank) {
            if (rows) rows[k] = r;
            k++;
        }
    *n_rows = k;
    return FLUX_OK;
}

static int render_common(flux_ctx *ctx, const uint32_t *rows, uint32_t n_rows, double *d_out, cudaStream_t user_stream,
                         bool sync_to_user, bool out_by_row = false) {
    if (!ctx->have_scene) return fail(ctx, FLUX_ERR_STATE, "render: scene not set");
    if (!ctx->have_samples) return fail(ctx, FLUX_ERR_STATE, "render: sample sets not set");
    if (!ctx->index_valid()) return fail(ctx, FLUX_ERR_STATE, "render: set-index map not set (or set for another image size / number of sample sets)");
    if (n_rows == 0) return FLUX_OK;
    if (!rows || !d_out) return fail(ctx, FLUX_ERR_INVALID, "render: null rows or output");
    for (uint32_t k = 0; k < n_rows; k++) {
        if (rows[k] >= ctx->cam.H) return fail(ctx, FLUX_ERR_INVALID, "render: row out of range");
        if (k && rows[k] <= rows[k - 1]) return fail(ctx, FLUX_ERR_INVALID, "render: rows must be strictly ascending");
    }
    if ((uint64_t)n_rows * ctx->cam.W > 0xFFFFFFFFull) return fail(ctx, FLUX_ERR_INVALID, "render: too many pixels in one call");
    cudaStream_t st = ctx->stream;
    CK(ctx->rows.reserve(n_rows));
    if (sync_to_user) {
        // order after work already queued on the caller's stream
        CK(cudaEventRecord(ctx->ev1, user_stream));
        CK(cudaStreamWaitEvent(st, ctx->ev1, 0));
    }
    CK(cudaMemcpyAsync(ctx->rows.p, rows, (size_t)n_rows * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(ctx->work_counter.p, 0, sizeof(unsigned int), st));
    RenderParams p{};
    p.scene = ctx->scene;
    p.cam = ctx->cam;
    p.ss = ctx->ss;
    p.set_index = ctx->set_index.p;
    p.rows = ctx->rows.p;
    p.n_rows = n_rows;
    p.out = d_out;
    p.out_by_row = out_by_row ? 1u : 0u;
    p.counters = ctx->counters.p;
    p.work_counter = ctx->work_counter.p;
    std::memcpy(p.cull, ctx->cull, sizeof(p.cull));
    p.cull_cmax = ctx->cull_cmax;
    p.primary_mask_ok = ctx->samples_in_range ? 1u : 0u;
    p.i_begin = 0;
    p.i_end = ctx->ss.n;
    p.accum = nullptr;
    CK(cudaEventRecord(ctx->ev0, st));
    const bool regen_ok = regen_kernel_applicable(p);
    if (ctx->kernel_mode == 2 && !regen_ok)
        return fail(ctx, FLUX_ERR_INVALID, "render: regeneration kernel needs spp >= 64 and a BVH scene or a sphere/plane scene that fits shared memory");
    const bool wave2_ok = wave2_kernel_applicable(p);
    if (ctx->kernel_mode == 4 && !wave2_ok)
        return fail(ctx, FLUX_ERR_INVALID, "render: wavefront-2 kernel needs spp >= 256, depth <= 8 and a sphere/plane scene of at most 128 spheres");
    // automatic choice on BVH scenes (measured r1): the wavefront kernel wins while the tree is small (config 4, 67
    // spheres: 2.72 vs 1.90 Gsamples/s), the regeneration kernel on deep trees whose traversal lengths vary a lot
    // (config 3, 1 M triangles: 567 vs 426 Msamples/s)
    const bool wave2_auto = wave2_ok && (!p.scene.use_bvh || p.scene.bvh_n_nodes <= 2048);   // (1024 while leaves held two primitives: the same scenes)
    // ... and on sphere/plane scenes the wavefront kernel's FP32-culled linear scan beats its own BVH owner stage up
    // to about 120 spheres (measured r1, config-4 scene with 40 / 85 / 125 spheres: 5.21 vs 2.69, 2.98 vs 2.34, 2.29 vs
    // 2.32 Gsamples/s), so a BVH built for the other kernels is left aside here unless it was asked for
    RenderParams q = p;
    q.scene.use_bvh = 0;
    const bool wave2_linear = ctx->accel_mode == 0 && p.scene.use_bvh && p.scene.n_tris == 0 &&
                              p.scene.n_spheres <= FLUX_WAVE2_LINEAR_MAX && wave2_kernel_applicable(q);
    if (wave2_linear && (ctx->kernel_mode == 0 || ctx->kernel_mode == 4))
        launch_render_wave2(q, ctx->count, ctx->sm_count, st);
    else if ((wave2_auto && ctx->kernel_mode == 0) || (wave2_ok && ctx->kernel_mode == 4))
        launch_render_wave2(p, ctx->count, ctx->sm_count, st);
    else if (regen_ok && ctx->kernel_mode != 1)
        launch_render_regen(p, ctx->count, ctx->sm_count, st);
    else
        launch_render(p, ctx->count, ctx->sm_count, st);
    ctx->launches += 1;
    CK(cudaEventRecord(ctx->ev1, st));
    ctx->ms_pending = true;
    CK(cudaGetLastError());
    if (sync_to_user) CK(cudaStreamWaitEvent(user_stream, ctx->ev1, 0));
    return FLUX_OK;
}

int flux_render_row_list_device(flux_ctx *ctx, const uint32_t *rows, uint32_t n_rows, double *d_out_rgb, void *cuda_stream) try {
    if (!ctx) return FLUX_ERR_INVALID;
    DeviceGuard g(ctx->device);
    // rows are copied from pageable host memory: the copy is staged before return
    return render_common(ctx, rows, n_rows, d_out_rgb, (cudaStream_t)cuda_stream, true);
} catch (...) {
    return caught(ctx, "flux_render_row_list_device");
}

int flux_render_row_list(flux_ctx *ctx, const uint32_t *rows, uint32_t n_rows, double *out_rgb) try {
    if (!ctx) return FLUX_ERR_INVALID;
    if (n_rows && !out_rgb) return fail(ctx, FLUX_ERR_INVALID, "render: null output");
    DeviceGuard g(ctx->device);
    const size_t elems = (size_t)n_rows * ctx->cam.W * 3;
    CK(ctx->out.reserve(elems));
    int rc = render_common(ctx, rows, n_rows, ctx->out.p, nullptr, false);
    if (rc) return rc;
    if (elems) CK(cudaMemcpyAsync(out_rgb, ctx->out.p, elems * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (n_rows) CK(cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));
    return FLUX_OK;
} catch (...) {
    return caught(ctx, "flux_render_row_list");
}

// ---- multi-GPU frame assembly over peer memory (include/fluxb200.h) ----------------------------------------------
}  // extern "C"

struct flux_frame {
    int device = 0;          // device the mapping belongs to (the caller's GPU, not necessarily the owner's)
    double *p = nullptr;     // [H][W][3] on the owner's GPU
    uint32_t W = 0, H = 0;
    int kind = 0;            // 0 owner (cudaMalloc), 1 CUDA IPC mapping, 2 peer alias inside one process
};

extern "C" {

int flux_frame_create(flux_ctx *ctx, uint32_t W, uint32_t H, flux_frame **out) try {
    if (!ctx) return FLUX_ERR_INVALID;
    if (!out || W == 0 || H == 0) return fail(ctx, FLUX_ERR_INVALID, "flux_frame_create: bad arguments");
    *out = nullptr;
    DeviceGuard g(ctx->device);
    flux_frame *f = new (std::nothrow) flux_frame();
    if (!f) return fail(ctx, FLUX_ERR_INVALID, "flux_frame_create: out of memory");
    f->device = ctx->device; f->W = W; f->H = H; f->kind = 0;
    const size_t bytes = (size_t)W * H * 3 * sizeof(double);
    // plain cudaMalloc (not a pool, not VMM): the allocation must be exportable as a CUDA IPC handle
    cudaError_t e = cudaMalloc((void **)&f->p, bytes);
    if (e == cudaSuccess) e = cudaMemsetAsync(f->p, 0, bytes, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        if (f->p) cudaFree(f->p);
        delete f;
        return fail(ctx, FLUX_ERR_CUDA, std::string("flux_frame_create: ") + cudaGetErrorString(e));
    }
    *out = f;
    return FLUX_OK;
} catch (...) {
    return caught(ctx, "flux_frame_create");
}

int flux_frame_export(flux_frame *f, unsigned char *handle) {
    if (!f || !handle || f->kind != 0) return FLUX_ERR_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == FLUX_FRAME_HANDLE_BYTES, "CUDA IPC handle size");
    DeviceGuard g(f->device);
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, f->p) != cudaSuccess) {
        g_create_error = std::string("flux_frame_export: ") + cudaGetErrorString(cudaGetLastError());
        return FLUX_ERR_CUDA;
    }
    std::memcpy(handle, &h, sizeof h);
    return FLUX_OK;
}

int flux_frame_open_ipc(flux_ctx *ctx, const unsigned char *handle, uint32_t W, uint32_t H, flux_frame **out) try {
    if (!ctx) return FLUX_ERR_INVALID;
    if (!out || !handle || W == 0 || H == 0) return fail(ctx, FLUX_ERR_INVALID, "flux_frame_open_ipc: bad arguments");
    *out = nullptr;
    DeviceGuard g(ctx->device);
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof h);
    void *p = nullptr;
    CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    flux_frame *f = new (std::nothrow) flux_frame();
    if (!f) {
        cudaIpcCloseMemHandle(p);
        return fail(ctx, FLUX_ERR_INVALID, "flux_frame_open_ipc: out of memory");
    }
    f->device = ctx->device; f->p = (double *)p; f->W = W; f->H = H; f->kind = 1;
    *out = f;
    return FLUX_OK;
} catch (...) {
    return caught(ctx, "flux_frame_open_ipc");
}

int flux_frame_open_peer(flux_ctx *ctx, flux_frame *owner, flux_frame **out) try {
    if (!ctx) return FLUX_ERR_INVALID;
    if (!out || !owner || owner->kind != 0) return fail(ctx, FLUX_ERR_INVALID, "flux_frame_open_peer: bad arguments");
    *out = nullptr;
    DeviceGuard g(ctx->device);
    if (owner->device != ctx->device) {
        int can = 0;
        CK(cudaDeviceCanAccessPeer(&can, ctx->device, owner->device));
        if (!can)
            return fail(ctx, FLUX_ERR_CUDA, "flux_frame_open_peer: device " + std::to_string(ctx->device) +
                                                " has no peer access to device " + std::to_string(owner->device));
        const cudaError_t e = cudaDeviceEnablePeerAccess(owner->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
            return fail(ctx, FLUX_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
        cudaGetLastError();
    }
    flux_frame *f = new (std::nothrow) flux_frame();
    if (!f) return fail(ctx, FLUX_ERR_INVALID, "flux_frame_open_peer: out of memory");
    f->device = ctx->device; f->p = owner->p; f->W = owner->W; f->H = owner->H; f->kind = 2;
    *out = f;
    return FLUX_OK;
} catch (...) {
    return caught(ctx, "flux_frame_open_peer");
}

int flux_render_row_list_into_frame(flux_ctx *ctx, const uint32_t *rows, uint32_t n_rows, flux_frame *frame, void *cuda_stream) try {
    if (!ctx) return FLUX_ERR_INVALID;
    if (!frame) return fail(ctx, FLUX_ERR_INVALID, "render: null frame");
    if (ctx->have_scene && (frame->W != ctx->cam.W || frame->H != ctx->cam.H))
        return fail(ctx, FLUX_ERR_INVALID, "render: the frame has another size than the scene's image");
    if (frame->device != ctx->device) return fail(ctx, FLUX_ERR_INVALID, "render: the frame was opened for another device");
    DeviceGuard g(ctx->device);
    return render_common(ctx, rows, n_rows, frame->p, (cudaStream_t)cuda_stream, true, true);
} catch (...) {
    return caught(ctx, "flux_render_row_list_into_frame");
}

int flux_ctx_sync(flux_ctx *ctx) {
    if (!ctx) return FLUX_ERR_INVALID;
    DeviceGuard g(ctx->device);
    CK(cudaStreamSynchronize(ctx->stream));
    return FLUX_OK;
}

int flux_frame_read(flux_frame *f, double *host_rgb) {
    if (!f || !host_rgb) return FLUX_ERR_INVALID;
    DeviceGuard g(f->device);
    if (cudaMemcpy(host_rgb, f->p, (size_t)f->W * f->H * 3 * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) {
        g_create_error = std::string("flux_frame_read: ") + cudaGetErrorString(cudaGetLastError());
        return FLUX_ERR_CUDA;
    }
    return FLUX_OK;
}

int flux_frame_device_ptr(flux_frame *f, void **p) {
    if (!f || !p) return FLUX_ERR_INVALID;
    *p = f->p;
    return FLUX_OK;
}

int flux_frame_close(flux_frame *f) {
    if (!f) return FLUX_ERR_INVALID;
    DeviceGuard g(f->device);
    if (f->kind == 0) cudaFree(f->p);
    else if (f->kind == 1) cudaIpcCloseMemHandle(f->p);
    cudaGetLastError();
    delete f;
    return FLUX_OK;
}

int flux_progressive_begin(flux_ctx *ctx, const uint32_t *rows, uint32_t n_rows) try {
    if (!ctx) return FLUX_ERR_INVALID;
    ctx->prog_active = false;
    if (!ctx->have_scene) return fail(ctx, FLUX_ERR_STATE, "progressive: scene not set");
    if (n_rows && !rows) return fail(ctx, FLUX_ERR_INVALID, "progressive: null rows");
    for (uint32_t k = 0; k < n_rows; k++) {
        if (rows[k] >= ctx->cam.H) return fail(ctx, FLUX_ERR_INVALID, "progressive: row out of range");
        if (k && rows[k] <= rows[k - 1]) return fail(ctx, FLUX_ERR_INVALID, "progressive: rows must be strictly ascending");
    }
    if ((uint64_t)n_rows * ctx->cam.W > 0xFFFFFFFFull) return fail(ctx, FLUX_ERR_INVALID, "progressive: too many pixels in one call");
    DeviceGuard g(ctx->device);
    const size_t elems = (size_t)n_rows * ctx->cam.W * 3;
    CK(ctx->accum.reserve(elems));
    CK(ctx->rows.reserve(n_rows));
    if (elems) CK(cudaMemsetAsync(ctx->accum.p, 0, elems * sizeof(double), ctx->stream));
    ctx->prog_rows.assign(rows, rows + n_rows);
    ctx->prog_done = 0;
    ctx->prog_active = true;
    return FLUX_OK;
} catch (...) {
    return caught(ctx, "flux_progressive_begin");
}

int flux_progressive_pass(flux_ctx *ctx, uint32_t sample_begin, uint32_t sample_end, double *out_rgb) try {
    if (!ctx) return FLUX_ERR_INVALID;
    if (!ctx->prog_active) return fail(ctx, FLUX_ERR_STATE, "progressive: flux_progressive_begin not called (or the scene changed since)");
    if (!ctx->have_samples) return fail(ctx, FLUX_ERR_STATE, "render: sample sets not set");
    if (!ctx->index_valid()) return fail(ctx, FLUX_ERR_STATE, "render: set-index map not set (or set for another image size / number of sample sets)");
    if (sample_begin != ctx->prog_done) return fail(ctx, FLUX_ERR_INVALID, "progressive: passes must continue where the last one ended");
    if (sample_end <= sample_begin || sample_end > ctx->ss.n) return fail(ctx, FLUX_ERR_INVALID, "progressive: bad sample range");
    const uint32_t n_rows = (uint32_t)ctx->prog_rows.size();
    if (n_rows == 0) return FLUX_OK;
    DeviceGuard g(ctx->device);
    cudaStream_t st = ctx->stream;
    const uint32_t npix = n_rows * ctx->cam.W;
    CK(cudaMemcpyAsync(ctx->rows.p, ctx->prog_rows.data(), (size_t)n_rows * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(ctx->work_counter.p, 0, sizeof(unsigned int), st));
    RenderParams p{};
    p.scene = ctx->scene;
    p.cam = ctx->cam;
    p.ss = ctx->ss;
    p.set_index = ctx->set_index.p;
    p.rows = ctx->rows.p;
    p.n_rows = n_rows;
    p.out = nullptr;
    p.counters = ctx->counters.p;
    p.work_counter = ctx->work_counter.p;
    std::memcpy(p.cull, ctx->cull, sizeof(p.cull));   // the direct kernel classifies sphere boxes in FP32 too (flux_cull.cuh)
    p.cull_cmax = ctx->cull_cmax;
    p.primary_mask_ok = ctx->samples_in_range ? 1u : 0u;
    p.i_begin = sample_begin;
    p.i_end = sample_end;
    p.accum = ctx->accum.p;
    CK(cudaEventRecord(ctx->ev0, st));
    launch_render(p, ctx->count, ctx->sm_count, st);
    ctx->launches += 1;
    ctx->prog_done = sample_end;
    if (out_rgb) {
        CK(ctx->out.reserve((size_t)npix * 3));
        // 1 / count; with every sample in, exactly the reference's pixel_denom (trace.rs:59)
        launch_resolve_accum(ctx->accum.p, ctx->out.p, npix, 1.0 / (double)sample_end, st);
        ctx->launches += 1;
    }
    CK(cudaEventRecord(ctx->ev1, st));
    CK(cudaGetLastError());
    if (out_rgb) CK(cudaMemcpyAsync(out_rgb, ctx->out.p, (size_t)npix * 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));
    return FLUX_OK;
} catch (...) {
    return caught(ctx, "flux_progressive_pass");
}

static int row_range(flux_ctx *ctx, uint32_t a, uint32_t b, std::vector<uint32_t> &rows) {
    if (!ctx->have_scene) return fail(ctx, FLUX_ERR_STATE, "render: scene not set");   // (before a row list of any length is made)
    if (b < a) return fail(ctx, FLUX_ERR_INVALID, "render: row_end < row_start");
    if (b >= ctx->cam.H) return fail(ctx, FLUX_ERR_INVALID, "render: row out of range");
    rows.resize((size_t)b - a + 1);
    for (uint32_t r = a; r <= b; r++) rows[r - a] = r;
    return FLUX_OK;
}

int flux_render_rows(flux_ctx *ctx, uint32_t row_start, uint32_t row_end_inclusive, double *out_rgb) try {
    if (!ctx) return FLUX_ERR_INVALID;
    std::vector<uint32_t> rows;
    int rc = row_range(ctx, row_start, row_end_inclusive, rows);
    if (rc) return rc;
    return flux_render_row_list(ctx, rows.data(), (uint32_t)rows.size(), out_rgb);
} catch (...) {
    return caught(ctx, "flux_render_rows");
}

int flux_render_rows_device(flux_ctx *ctx, uint32_t row_start, uint32_t row_end_inclusive, double *d_out_rgb, void *cuda_stream) try {
    if (!ctx) return FLUX_ERR_INVALID;
    std::vector<uint32_t> rows;
    int rc = row_range(ctx, row_start, row_end_inclusive, rows);
    if (rc) return rc;
    return flux_render_row_list_device(ctx, rows.data(), (uint32_t)rows.size(), d_out_rgb, cuda_stream);
} catch (...) {
    return caught(ctx, "flux_render_rows_device");
}

int flux_trace_rays_device(flux_ctx *ctx, uint64_t n, const double *d_o, const double *d_d, int32_t *d_hit, double *d_t,
                           void *cuda_stream) try {
    if (!ctx) return FLUX_ERR_INVALID;
    if (!ctx->have_scene) return fail(ctx, FLUX_ERR_STATE, "flux_trace_rays: scene not set");
    if (n == 0) return FLUX_OK;
    if (!d_o || !d_d || !d_hit || !d_t) return fail(ctx, FLUX_ERR_INVALID, "flux_trace_rays: null pointer");
    DeviceGuard g(ctx->device);
    cudaStream_t us = (cudaStream_t)cuda_stream;
    CK(cudaEventRecord(ctx->ev1, us));
    CK(cudaStreamWaitEvent(ctx->stream, ctx->ev1, 0));
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    launch_trace_rays(ctx->scene, n, d_o, d_d, d_hit, d_t, ctx->sm_count, ctx->stream, ctx->count ? ctx->counters.p : nullptr, ctx->trace_work.p);
    ctx->launches += 1;
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    ctx->ms_pending = true;
    CK(cudaGetLastError());
    CK(cudaStreamWaitEvent(us, ctx->ev1, 0));
    return FLUX_OK;
} catch (...) {
    return caught(ctx, "flux_trace_rays_device");
}

// Host-buffer form: the batch goes through the device in pieces of TRACE_PIECE rays on three streams — upload of
// piece k+1, traversal of piece k and download of piece k-1 overlap (the PCIe link, 48 bytes in and 12 out per ray, is
// the bound of this entry point: the kernel alone runs at several times its rate).  Pinned caller memory makes the
// copies truly asynchronous; pageable memory works, staged by the driver.
int flux_trace_rays(flux_ctx *ctx, uint64_t n, const double *o, const double *d, int32_t *hit, double *t) try {
    if (!ctx) return FLUX_ERR_INVALID;
    if (!ctx->have_scene) return fail(ctx, FLUX_ERR_STATE, "flux_trace_rays: scene not set");
    if (n == 0) return FLUX_OK;
    if (!o || !d || !hit || !t) return fail(ctx, FLUX_ERR_INVALID, "flux_trace_rays: null pointer");
    DeviceGuard g(ctx->device);
    constexpr uint64_t TRACE_PIECE = 1ull << 22;   // 4 Mi rays: 200 MB in, 50 MB out per piece
    constexpr int NB = 3;
    const uint64_t piece = std::min<uint64_t>(n, TRACE_PIECE);
    const int nb = n > piece ? NB : 1;
    CK(ctx->ray_o.reserve(3 * piece * nb));
    CK(ctx->ray_d.reserve(3 * piece * nb));
    CK(ctx->ray_t.reserve(piece * nb));
    CK(ctx->ray_hit.reserve(piece * nb));
    if (!ctx->s_in) CK(cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
    if (!ctx->s_out) CK(cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
    const uint64_t n_pieces = (n + piece - 1) / piece;
    struct Ev {
        std::vector<cudaEvent_t> v;
        ~Ev() { for (cudaEvent_t e : v) if (e) cudaEventDestroy(e); }
    } ev;   // per piece: uploaded, kernel start, kernel end; per buffer: downloaded
    ev.v.assign(3 * n_pieces + NB, nullptr);
    for (size_t k = 0; k < ev.v.size(); k++) {
        const bool timing = k < 3 * n_pieces && (k % 3) != 0;
        CK(cudaEventCreateWithFlags(&ev.v[k], timing ? cudaEventDefault : cudaEventDisableTiming));
    }
    cudaEvent_t *freed = ev.v.data() + 3 * n_pieces;
    struct Drain {   // an error return must not leave copies to or from the caller's buffers in flight
        flux_ctx *c;
        ~Drain() { cudaStreamSynchronize(c->s_in); cudaStreamSynchronize(c->stream); cudaStreamSynchronize(c->s_out); cudaGetLastError(); }
    } drain{ctx};
    for (uint64_t k = 0; k < n_pieces; k++) {
        const uint64_t off = k * piece, c = std::min(piece, n - off);
        const int b = (int)(k % nb);
        double *bo = ctx->ray_o.p + 3 * piece * b, *bd = ctx->ray_d.p + 3 * piece * b, *bt = ctx->ray_t.p + piece * b;
        int32_t *bh = ctx->ray_hit.p + piece * b;
        if (k >= (uint64_t)nb) CK(cudaStreamWaitEvent(ctx->s_in, freed[b], 0));   // the buffer's last results have left
        CK(cudaMemcpyAsync(bo, o + 3 * off, 24 * c, cudaMemcpyHostToDevice, ctx->s_in));
        CK(cudaMemcpyAsync(bd, d + 3 * off, 24 * c, cudaMemcpyHostToDevice, ctx->s_in));
        CK(cudaEventRecord(ev.v[3 * k], ctx->s_in));
        CK(cudaStreamWaitEvent(ctx->stream, ev.v[3 * k], 0));
        CK(cudaEventRecord(ev.v[3 * k + 1], ctx->stream));
        launch_trace_rays(ctx->scene, c, bo, bd, bh, bt, ctx->sm_count, ctx->stream, ctx->count ? ctx->counters.p : nullptr, ctx->trace_work.p);
        ctx->launches += 1;
        CK(cudaEventRecord(ev.v[3 * k + 2], ctx->stream));
        CK(cudaGetLastError());
        CK(cudaStreamWaitEvent(ctx->s_out, ev.v[3 * k + 2], 0));
        CK(cudaMemcpyAsync(hit + off, bh, 4 * c, cudaMemcpyDeviceToHost, ctx->s_out));
        CK(cudaMemcpyAsync(t + off, bt, 8 * c, cudaMemcpyDeviceToHost, ctx->s_out));
        CK(cudaEventRecord(freed[b], ctx->s_out));
    }
    CK(cudaStreamSynchronize(ctx->s_out));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaStreamSynchronize(ctx->s_in));
    float total_ms = 0.f;
    for (uint64_t k = 0; k < n_pieces; k++) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, ev.v[3 * k + 1], ev.v[3 * k + 2]));
        total_ms += ms;
    }
    ctx->last_ms = total_ms;   // the kernels alone
    ctx->ms_pending = false;
    return FLUX_OK;
} catch (...) {
    return caught(ctx, "flux_trace_rays");
}

int flux_enable_counters(flux_ctx *ctx, int enable) {
    if (!ctx) return FLUX_ERR_INVALID;
    ctx->count = enable != 0;
    return FLUX_OK;
}

int flux_get_counters(flux_ctx *ctx, flux_counters *out) {
    if (!ctx) return FLUX_ERR_INVALID;
    if (!out) return fail(ctx, FLUX_ERR_INVALID, "flux_get_counters: null output");
    static_assert(sizeof(flux_counters) == CN_COUNT * sizeof(uint64_t), "flux_counters layout");
    DeviceGuard g(ctx->device);
    CK(cudaMemcpyAsync(out, ctx->counters.p, sizeof(flux_counters), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return FLUX_OK;
}

int flux_reset_counters(flux_ctx *ctx) {
    if (!ctx) return FLUX_ERR_INVALID;
    DeviceGuard g(ctx->device);
    CK(cudaMemsetAsync(ctx->counters.p, 0, sizeof(flux_counters), ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return FLUX_OK;
}

int flux_last_kernel_ms(flux_ctx *ctx, float *ms) {
    if (!ctx || !ms) return FLUX_ERR_INVALID;
    DeviceGuard g(ctx->device);
    // device-pointer calls do not synchronise: resolve the events lazily here
    if (ctx->ms_pending && cudaEventQuery(ctx->ev1) == cudaSuccess) {
        float v = 0.f;
        if (cudaEventElapsedTime(&v, ctx->ev0, ctx->ev1) == cudaSuccess) ctx->last_ms = v;
        ctx->ms_pending = false;
    }
    cudaGetLastError();
    *ms = ctx->last_ms;
    return FLUX_OK;
}

int flux_launch_count(flux_ctx *ctx, uint64_t *n) {
    if (!ctx || !n) return FLUX_ERR_INVALID;
    *n = ctx->launches;
    return FLUX_OK;
}

int flux_set_accel_mode(flux_ctx *ctx, int mode) {
    if (!ctx) return FLUX_ERR_INVALID;
    if (mode < 0 || mode > 2) return fail(ctx, FLUX_ERR_INVALID, "flux_set_accel_mode: mode must be 0, 1 or 2");
    ctx->accel_mode = mode;
    return FLUX_OK;
}

int flux_bvh_hash(const flux_scene_flat *s, uint64_t *hash) try {
    if (!s || !hash) return FLUX_ERR_INVALID;
    const uint32_t ns = s->n_spheres, nt = s->n_triangles;
    if ((ns && (!s->sphere_center || !s->sphere_radius || !s->sphere_invert || !s->sphere_shape_id || !s->sphere_material)) ||
        (nt && (!s->tri_v0 || !s->tri_v1 || !s->tri_v2 || !s->tri_shape_id || !s->tri_material)))
        return FLUX_ERR_INVALID;
    std::vector<double> sph;
    std::vector<uint32_t> sph_meta;
    flux_raw_vector<double> tri;
    flux_raw_vector<uint32_t> tri_meta;
    flatten_spheres(s, sph, sph_meta);
    flatten_triangles(s, tri, tri_meta);
    BvhBuild bb;
    std::string err;
    if (!build_bvh4(sph.data(), sph_meta.data(), ns, tri.data(), tri_meta.data(), s->tri_v1, s->tri_v2, nt, bb, err)) {
        g_create_error = err;
        return FLUX_ERR_INVALID;
    }
    uint64_t h = 1469598103934665603ull;   // FNV-1a over everything the traversal reads
    auto mix = [&h](const void *p, size_t n) {
        const unsigned char *b = static_cast<const unsigned char *>(p);
        for (size_t i = 0; i < n; i++) h = (h ^ b[i]) * 1099511628211ull;
    };
    mix(bb.nodes.data(), bb.nodes.size() * sizeof(BvhNode4));
    mix(bb.prims.data(), bb.prims.size() * sizeof(uint32_t));
    mix(bb.linear.data(), bb.linear.size() * sizeof(uint32_t));
    mix(bb.sph.data(), bb.sph.size() * sizeof(SphRec));
    mix(bb.tri.data(), bb.tri.size() * sizeof(TriRec));
    const uint64_t tail[3] = {bb.depth, bb.leaf_size, 0};
    mix(tail, sizeof tail);
    mix(&bb.extent, sizeof bb.extent);
    *hash = h;
    return FLUX_OK;
} catch (...) {
    return caught(nullptr, "flux_bvh_hash");
}

int flux_bvh_describe(const flux_scene_flat *s, uint64_t out[8]) try {
    if (!s || !out) return FLUX_ERR_INVALID;
    const uint32_t ns = s->n_spheres, nt = s->n_triangles;
    if ((ns && (!s->sphere_center || !s->sphere_radius || !s->sphere_invert || !s->sphere_shape_id || !s->sphere_material)) ||
        (nt && (!s->tri_v0 || !s->tri_v1 || !s->tri_v2 || !s->tri_shape_id || !s->tri_material)))
        return FLUX_ERR_INVALID;
    std::vector<double> sph;
    std::vector<uint32_t> sph_meta;
    flux_raw_vector<double> tri;
    flux_raw_vector<uint32_t> tri_meta;
    flatten_spheres(s, sph, sph_meta);
    flatten_triangles(s, tri, tri_meta);
    BvhBuild bb;
    std::string err;
    if (!build_bvh4(sph.data(), sph_meta.data(), ns, tri.data(), tri_meta.data(), s->tri_v1, s->tri_v2, nt, bb, err)) {
        g_create_error = err;
        return FLUX_ERR_INVALID;
    }
    uint64_t violations = 0;
    std::vector<uint32_t> seen_s(ns, 0), seen_t(nt, 0);
    for (uint32_t i : bb.linear) seen_s[i]++;
    // walk the tree: (node, slot box of the parent) pairs
    struct Job { uint32_t ref; double lo[3], hi[3]; };
    std::vector<Job> stack;
    const double inf = std::numeric_limits<double>::infinity();
    if (!bb.nodes.empty()) {
        Job j; j.ref = 0;
        for (int k = 0; k < 3; k++) j.lo[k] = -inf, j.hi[k] = inf;
        stack.push_back(j);
    }
    while (!stack.empty()) {
        const Job j = stack.back();
        stack.pop_back();
        if (j.ref & BVH_LEAF) {
            const bool direct = (j.ref & BVH_DIRECT) != 0u;
            const uint32_t off = (j.ref & 0x3FFFFFFFu) >> 3, cnt = direct ? 1u : (j.ref & 7u) + 1u;
            for (uint32_t k = 0; k < cnt; k++) {
                const uint32_t pr = direct ? (((j.ref >> 28) & 3u) << 30) | (j.ref & 0x0FFFFFFFu) : bb.prims[off + k], idx = pr & 0x3FFFFFFFu;
                double lo[3], hi[3];
                if ((pr >> 30) == KIND_SPHERE) {
                    seen_s[idx]++;
                    const SphRec &q = bb.sph[idx];
                    lo[0] = std::min(q.c0x, q.c1x); hi[0] = std::max(q.c0x, q.c1x);
                    lo[1] = std::min(q.c0y, q.c1y); hi[1] = std::max(q.c0y, q.c1y);
                    lo[2] = std::min(q.c0z, q.c1z); hi[2] = std::max(q.c0z, q.c1z);
                } else {
                    seen_t[idx]++;
                    for (int a = 0; a < 3; a++) {
                        const double v0 = s->tri_v0[3 * (size_t)idx + a], v1 = s->tri_v1[3 * (size_t)idx + a], v2 = s->tri_v2[3 * (size_t)idx + a];
                        lo[a] = std::min({v0, v1, v2});
                        hi[a] = std::max({v0, v1, v2});
                    }
                }
                for (int a = 0; a < 3; a++)
                    if (!(j.lo[a] < lo[a] && hi[a] < j.hi[a])) violations++;   // strictly inside: the padding
            }
            continue;
        }
        const BvhNode4 &n = bb.nodes[j.ref];
        for (int c = 0; c < 4; c++) {
            if (n.child[c] == BVH_EMPTY) continue;
            Job q; q.ref = n.child[c];
            for (int a = 0; a < 3; a++) {
                q.lo[a] = n.lo[a][c]; q.hi[a] = n.hi[a][c];
                if (!(j.lo[a] <= q.lo[a] && q.hi[a] <= j.hi[a])) violations++;
            }
            stack.push_back(q);
        }
    }
    uint64_t miscount = 0;
    for (uint32_t v : seen_s) miscount += v != 1;
    for (uint32_t v : seen_t) miscount += v != 1;
    out[0] = bb.nodes.size(); out[1] = bb.depth; out[2] = bb.leaf_size; out[3] = bb.linear.size();
    out[4] = bb.prims.size(); out[5] = violations; out[6] = miscount;
    out[7] = ((uint64_t)ns + nt > FLUX_LINEAR_LIMIT) ? 1 : 0;
    return FLUX_OK;
} catch (...) {
    return caught(nullptr, "flux_bvh_describe");
}

int flux_set_kernel_mode(flux_ctx *ctx, int mode) {
    if (!ctx) return FLUX_ERR_INVALID;
    if (mode < 0 || mode > 4 || mode == 3) return fail(ctx, FLUX_ERR_INVALID, "flux_set_kernel_mode: mode must be 0, 1, 2 or 4 (3, the first-generation wavefront kernel, was removed)");
    ctx->kernel_mode = mode;
    return FLUX_OK;
}

int flux_set_glossy_table(flux_ctx *ctx, int enable) {
    if (!ctx) return FLUX_ERR_INVALID;
    ctx->use_gtable = enable != 0;
    return FLUX_OK;
}

int flux_measure_fp64_peak(flux_ctx *ctx, double *ginstr_per_s) {
    if (!ctx || !ginstr_per_s) return FLUX_ERR_INVALID;
    DeviceGuard g(ctx->device);
    double best = 0.0;
    launch_fp64_peak(ctx->sm_count, 2000, ctx->sink.p, ctx->stream);  // warm-up
    for (int rep = 0; rep < 3; rep++) {
        CK(cudaEventRecord(ctx->ev0, ctx->stream));
        double instr = launch_fp64_peak(ctx->sm_count, 20000, ctx->sink.p, ctx->stream);
        CK(cudaEventRecord(ctx->ev1, ctx->stream));
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(ctx->stream));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        ctx->launches += 1;
        if (ms > 0.f) best = std::max(best, instr / (ms * 1e-3) / 1e9);
    }
    ctx->launches += 1;
    *ginstr_per_s = best;
    return FLUX_OK;
}

// Image::write, fluxcore/src/image.rs:42-60: "P3\n{W} {H}\n65535\n" then one
// "r g b" line per pixel, (c * 65535.99) as u16 (saturating, NaN -> 0).
int flux_write_ppm(const char *path, uint32_t width, uint32_t height, const double *rgb) {
    if (!path || !rgb) return FLUX_ERR_INVALID;
    FILE *f = fopen(path, "w");
    if (!f) return FLUX_ERR_INVALID;
    auto q = [](double c) -> unsigned {
        double v = c * 65535.99;
        if (!(v == v) || v <= 0.0) return 0u;
        if (v >= 65535.0) return 65535u;
        return (unsigned)v;
    };
    fprintf(f, "P3\n%u %u\n65535\n", width, height);
    const size_t n = (size_t)width * height;
    for (size_t i = 0; i < n; i++) fprintf(f, "%u %u %u\n", q(rgb[3 * i]), q(rgb[3 * i + 1]), q(rgb[3 * i + 2]));
    return fclose(f) == 0 ? FLUX_OK : FLUX_ERR_INVALID;
}

}  // extern "C"
