/*
 * fluxb200.h — C-ABI of libfluxb200.so, the B200-native replacement for the
 * per-pixel render loop of jtdaugherty/flux.
 *
 * Drop-in boundary.  In the reference the hot path is entered
 * at exactly one call site, fluxcore/src/workers.rs:46-64:
 *
 *     let scene  = Scene::from_data(job.scene_data, job.config);   // scene.rs:128
 *     let camera = Camera::new(.., job.config, image_width, ..);    // trace.rs:26
 *     loop { let r = camera.render(&scene, unit); .. RowsReady(r) } // trace.rs:53
 *
 * A Rust `GpuWorker: Worker` (manager.rs:232-236) binds the functions below
 * through a `-sys` crate (INTEGRATION.md shows the stub).  Everything that
 * crosses is plain data: pointers, sizes, f64/u32/u8 scalars.  No C++ or torch
 * types, no exceptions, no aborts.  Every function returns 0 on success or a
 * FLUX_ERR_* code; flux_last_error() returns the message for the last failure
 * on that context.
 *
 * Ownership: the caller owns every host buffer.  flux_set_* copy their inputs
 * to the device before returning and never retain caller pointers.  Output
 * buffers are fully written before a (host-pointer) call returns.
 *
 * Threading: a flux_ctx is single-caller (like the one LocalWorker thread,
 * workers.rs:42-75); several contexts may coexist.
 *
 * There is no CPU fallback: without a usable CUDA device flux_ctx_create
 * fails with FLUX_ERR_NO_DEVICE.
 */
#ifndef FLUXB200_H
#define FLUXB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes ------------------------------------------------------- */
enum {
    FLUX_OK = 0,
    FLUX_ERR_INVALID = 1,   /* bad argument (null pointer, index out of range, ...) */
    FLUX_ERR_CUDA = 2,      /* a CUDA runtime call failed; see flux_last_error   */
    FLUX_ERR_STATE = 3,     /* call order: scene / samples / set index not set   */
    FLUX_ERR_NO_DEVICE = 4  /* no CUDA device, or device id out of range         */
};

/* ---- plain-data scene ---------------------------------------------------- */

/* MaterialData variants, fluxcore/src/shapes.rs:42-83. */
enum {
    FLUX_MAT_MATTE = 0,      /* MatteData: color = diffuse_color, k = diffuse_coefficient
                                (ambient_color is carried by the reference but never read
                                on the path: materials.rs:19-33 uses diffuse_brdf only)  */
    FLUX_MAT_EMISSIVE = 1,   /* EmissiveData: color, k = power                            */
    FLUX_MAT_REFLECTIVE = 2, /* ReflectiveData: color = reflect_color, k = reflect_amount */
    FLUX_MAT_GLOSSY = 3      /* GlossyReflectiveData: + exp = reflect_exponent            */
};

typedef struct flux_material {
    uint32_t kind;     /* FLUX_MAT_*        */
    uint32_t _pad;
    double color[3];
    double k;
    double exp;
} flux_material;

/*
 * Flattened SceneData (fluxcore/src/scene.rs:42-66) + the shapes of
 * shapes.rs:18-37 as per-kind arrays.  `*_shape_id` is the position of the
 * shape in the YAML `shapes:` list; spheres, planes and triangles share that
 * one index space because closest-hit ties go to the earlier shape
 * (scene.rs:156-160 + common.rs:17-23).  Within each per-kind array shape ids
 * must be strictly increasing.
 *
 * Triangles are an EXTENSION (the reference has only Sphere and Plane,
 * scene.rs:71-74); see DESIGN.md "extensions" for their semantics.
 */
typedef struct flux_scene_flat {
    /* OutputSettings, scene.rs:58-63 */
    uint32_t image_width;
    uint32_t image_height;
    double pixel_size;
    /* background, scene.rs:45 */
    double background[3];
    /* CameraSettings, scene.rs:11-16 */
    double eye[3];
    double look_at[3];
    double up[3];
    /* CameraData, scene.rs:50-56 */
    double zoom_factor;
    double view_plane_distance;
    double focal_distance;
    double lens_radius;

    uint32_t n_materials;
    const flux_material *materials;

    uint32_t n_spheres;
    const double *sphere_center;       /* [n_spheres][3]                      */
    const double *sphere_radius;       /* [n_spheres]                         */
    const uint8_t *sphere_invert;      /* [n_spheres] 0/1, shapes.rs:22       */
    const uint32_t *sphere_shape_id;   /* [n_spheres]                         */
    const uint32_t *sphere_material;   /* [n_spheres] index into materials    */

    uint32_t n_planes;
    const double *plane_point;         /* [n_planes][3]                       */
    const double *plane_normal;        /* [n_planes][3], used as given        */
    const uint32_t *plane_shape_id;
    const uint32_t *plane_material;

    uint32_t n_triangles;              /* EXTENSION                           */
    const double *tri_v0;              /* [n_triangles][3]                    */
    const double *tri_v1;
    const double *tri_v2;
    const uint32_t *tri_shape_id;
    const uint32_t *tri_material;
} flux_scene_flat;

/* JobConfiguration, fluxcore/src/job.rs:49-53. */
typedef struct flux_job_config {
    uint32_t sample_root;        /* samples per pixel = sample_root^2 (trace.rs:59) */
    uint32_t max_trace_depth;    /* scene.rs:164                                    */
    uint32_t rows_per_work_unit; /* carried for the caller; not used by the library  */
} flux_job_config;

/*
 * Event counters of one or more render calls: the inputs of the
 * algorithmic-op model.  Only filled when counting is enabled
 * (flux_enable_counters), which selects an instrumented kernel.
 */
typedef struct flux_counters {
    uint64_t samples;       /* camera samples (paths)                    */
    uint64_t segments;      /* ray segments traced (Scene::hit calls)    */
    uint64_t bbox_tests;    /* BoundingBox::hit evaluations              */
    uint64_t bbox_pass;     /* ... that returned true                    */
    uint64_t disc_nonneg;   /* sphere quadratics with disc >= 0          */
    uint64_t t2_evals;      /* second root evaluated                     */
    uint64_t plane_tests;
    uint64_t tri_tests;
    uint64_t candidates;    /* Some(Hit) returned by a shape             */
    uint64_t hit_sphere;    /* closest hit was a sphere                  */
    uint64_t hit_plane;
    uint64_t hit_tri;
    uint64_t emissive;      /* path_shade events per material kind       */
    uint64_t matte;
    uint64_t specular;
    uint64_t glossy;
    uint64_t glossy_flip;   /* brdf.rs:67-71 branch taken                */
    uint64_t depth_cut;     /* scene.rs:164 returned black               */
    uint64_t miss;          /* scene.rs:168 returned background          */
    uint64_t nodes_visited; /* BVH nodes visited (extension path)        */
} flux_counters;

typedef struct flux_ctx flux_ctx;

/* ---- lifecycle ----------------------------------------------------------- */

/* Create a context bound to CUDA device `device`.  Replaces LocalWorker::new
 * (workers.rs:26).  One context drives one GPU; a multi-GPU job uses one
 * context per GPU, each rendering its shard (flux_shard_rows). */
int flux_ctx_create(int device, flux_ctx **out);
int flux_ctx_destroy(flux_ctx *ctx);
const char *flux_last_error(const flux_ctx *ctx); /* ctx may be NULL: last create error */
const char *flux_version(void);

/* ---- job setup ----------------------------------------------------------- */

/* Scene::from_data (scene.rs:128-154) + CameraBasis::new (scene.rs:29-34) +
 * the non-sample part of Camera::new (trace.rs:26-42).  Builds the sphere
 * bounding boxes exactly as Sphere::new (shapes.rs:154-169) and, when the
 * scene has more shapes than the linear-scan limit, the BVH extension. */
int flux_set_scene(flux_ctx *ctx, const flux_scene_flat *scene, const flux_job_config *cfg);

/* MasterSampleSets (sampling.rs:5-33) supplied explicitly, reference layout:
 *   pixel_xy [num_sets][root^2][2]             UnitSquareSample {x,y}
 *   disc_xy  [num_sets][root^2][2]             UnitDiscSample {x,y}
 *   hemi_xyz [num_sets][max_depth][root^2][3]  Vector3 */
int flux_set_samples(flux_ctx *ctx, uint32_t sample_root, uint32_t max_depth, uint32_t num_sets,
                     const double *pixel_xy, const double *disc_xy, const double *hemi_xyz);

/* MasterSampleSets::new on the device (N1): correlated multi-jittered pixel and
 * disc sets, multi-jittered e=0 hemisphere sets (sampling.rs:16-29 →
 * samplers/src/lib.rs:46-182) from a counter-based PRNG keyed by `seed`;
 * also fills the set-index map (one permutation per row, sampling.rs:35-40).
 * Requires flux_set_scene first (for root, depth, image size). */
int flux_generate_samples(flux_ctx *ctx, uint64_t seed, uint32_t num_sets);

/* Copy the device-resident sample sets / set-index map back (for parity tests
 * against the oracle).  Any pointer may be NULL to skip it. */
int flux_get_samples(flux_ctx *ctx, double *pixel_xy, double *disc_xy, double *hemi_xyz);
int flux_get_set_index(flux_ctx *ctx, uint32_t *idx /* [image_height][image_width] */);

/* Per-row permutation of sample-set indices (trace.rs:64,68-69) as an explicit
 * map: idx[row*image_width + col] in [0, num_sets). */
int flux_set_set_index(flux_ctx *ctx, const uint32_t *idx /* [image_height][image_width] */);

/* ---- the hot path -------------------------------------------------------- */

/* Camera::render(&scene, WorkUnit{row_start,row_end}) (trace.rs:53-97).
 * out_rgb: [(row_end-row_start+1)][image_width][3] f64, linear, averaged and
 * max_to_one-clamped exactly like WorkUnitResult.rows (manager.rs:25-28). */
int flux_render_rows(flux_ctx *ctx, uint32_t row_start, uint32_t row_end_inclusive,
                     double *out_rgb /* host */);

/* Progressive refinement (the reference's preview re-renders the whole job at another sample
 * root, flux/src/main.rs:296-315, and cancels between work units, manager.rs:66-69).  flux_progressive_begin
 * names the rows (strictly ascending) and clears their accumulator; each flux_progressive_pass traces samples
 * [sample_begin, sample_end) of every pixel of those rows — the same sample sets, set-index map and per-sample
 * arithmetic as Camera::render (trace.rs:53-97) — and adds their radiance to the accumulator.  Passes must follow
 * one another (sample_begin = the previous sample_end, 0 first; sample_end <= sample_root^2).  out_rgb (host,
 * [n_rows][image_width][3], may be NULL) receives the image of the samples so far: sum * (1 / sample_end), then
 * max_to_one.  After the last pass that is Camera::render's result up to the order of the per-pixel sum; one pass
 * over all samples is bit-identical to flux_render_rows with kernel mode 1.  The caller cancels by not issuing the
 * next pass.  flux_set_scene and the sample-set calls end a progression. */
int flux_progressive_begin(flux_ctx *ctx, const uint32_t *rows, uint32_t n_rows);
int flux_progressive_pass(flux_ctx *ctx, uint32_t sample_begin, uint32_t sample_end, double *out_rgb /* host or NULL */);

/* Same, writing to device memory on `cuda_stream` (a cudaStream_t, may be NULL)
 * without synchronising: for callers that keep the framebuffer on the GPU
 * (multi-GPU gather over NVLink). */
int flux_render_rows_device(flux_ctx *ctx, uint32_t row_start, uint32_t row_end_inclusive,
                            double *d_out_rgb, void *cuda_stream);

/* Multi-GPU sharding (replaces the bounded(1) work queue, manager.rs:100):
 * shard `rank` of `world` owns rows r with (r / tile_rows) % world == rank.
 * flux_shard_rows lists them in ascending order (rows may be NULL to count). */
int flux_shard_rows(uint32_t image_height, uint32_t tile_rows, uint32_t rank, uint32_t world,
                    uint32_t *rows, uint32_t *n_rows);
/* Render an arbitrary ascending row list; output packed in list order. */
int flux_render_row_list(flux_ctx *ctx, const uint32_t *rows, uint32_t n_rows, double *out_rgb);
int flux_render_row_list_device(flux_ctx *ctx, const uint32_t *rows, uint32_t n_rows,
                                double *d_out_rgb, void *cuda_stream);

/* ---- multi-GPU frame assembly over NVLink peer memory ------------------------
 * Replaces the reference's RowsReady stream from every worker into the manager's one ImageBuilder
 * (workers.rs:46-64 -> manager.rs:100,156-162; network form workers.rs:119-243): ONE framebuffer
 * [image_height][image_width][3] f64 lives on the owner's GPU and every GPU of the box renders its rows straight
 * into it — the render kernel's final per-pixel store goes through an NVLink / NVSwitch peer mapping — so there is
 * no gather collective, no packed per-GPU slice and no un-interleave copy.  The only cross-GPU synchronisation a
 * caller needs is "every GPU's render has finished" (flux_ctx_sync + a barrier of its own, or stream order).
 *
 *   owner process / thread:  flux_frame_create -> (flux_frame_export) -> ... -> flux_frame_read -> flux_frame_close
 *   other GPU, same process: flux_frame_open_peer(ctx, owner_frame)        (cudaDeviceEnablePeerAccess)
 *   other process:           flux_frame_open_ipc(ctx, handle, W, H)        (CUDA IPC memory handle, 64 bytes)
 *   every GPU:               flux_render_row_list_into_frame(ctx, rows, n_rows, frame, stream)
 */
typedef struct flux_frame flux_frame;
#define FLUX_FRAME_HANDLE_BYTES 64
int flux_frame_create(flux_ctx *ctx, uint32_t image_width, uint32_t image_height, flux_frame **out);
/* handle: FLUX_FRAME_HANDLE_BYTES opaque bytes another process on the same box passes to flux_frame_open_ipc. */
int flux_frame_export(flux_frame *frame, unsigned char *handle);
int flux_frame_open_ipc(flux_ctx *ctx, const unsigned char *handle, uint32_t image_width, uint32_t image_height,
                        flux_frame **out);
int flux_frame_open_peer(flux_ctx *ctx, flux_frame *owner_frame, flux_frame **out);
/* Camera::render for an ascending row list with every row written at its place in `frame` (row r of the image at
 * frame[r]).  Asynchronous like flux_render_row_list_device: ordered after work already queued on `cuda_stream`
 * (may be NULL) and that stream is made to wait for the render; flux_ctx_sync waits on the host. */
int flux_render_row_list_into_frame(flux_ctx *ctx, const uint32_t *rows, uint32_t n_rows, flux_frame *frame,
                                    void *cuda_stream);
int flux_ctx_sync(flux_ctx *ctx);
/* Whole frame to host memory (a blocking copy; any process / GPU that holds the frame may call it).  It does not
 * wait for renders still in flight: the caller orders it after them (flux_ctx_sync on every context that rendered
 * into the frame, and its own barrier across processes).  The owner must not close the frame while another process
 * still has it open. */
int flux_frame_read(flux_frame *frame, double *host_rgb /* [image_height][image_width][3] */);
int flux_frame_device_ptr(flux_frame *frame, void **device_ptr);
int flux_frame_close(flux_frame *frame);

/* Scene::hit (scene.rs:156-160) on an explicit ray batch.
 * origin_xyz, dir_xyz: [n][3]; hit_shape_id[n] = shape id or -1; t[n] = hit
 * distance (undefined on miss: written as +inf). */
int flux_trace_rays(flux_ctx *ctx, uint64_t n, const double *origin_xyz, const double *dir_xyz,
                    int32_t *hit_shape_id, double *t);
int flux_trace_rays_device(flux_ctx *ctx, uint64_t n, const double *d_origin_xyz,
                           const double *d_dir_xyz, int32_t *d_hit_shape_id, double *d_t,
                           void *cuda_stream);

/* ---- instrumentation ----------------------------------------------------- */
int flux_enable_counters(flux_ctx *ctx, int enable);
int flux_get_counters(flux_ctx *ctx, flux_counters *out); /* accumulated since last reset */
int flux_reset_counters(flux_ctx *ctx);
/* Device time in milliseconds of the kernels of the most recent render /
 * trace call (CUDA events on the launching stream). */
int flux_last_kernel_ms(flux_ctx *ctx, float *ms);
/* Number of kernel launches issued by this context so far. */
int flux_launch_count(flux_ctx *ctx, uint64_t *n);
/* Force the acceleration mode for closest-hit: 0 = auto, 1 = linear scan,
 * 2 = BVH.  Results are identical by construction; used by parity tests. */
int flux_set_accel_mode(flux_ctx *ctx, int mode);

/* Host-only description of the acceleration structure flux_set_scene would build for `scene` (EXTENSION;
 * needs no device, so the builder is testable on CPU).  Builds the BVH, checks that every primitive box lies
 * inside the box of the leaf slot that holds it and every child box inside its parent slot, and reports:
 * out[0] nodes, out[1] 4-wide levels, out[2] leaf size, out[3] shapes on the linear list, out[4] primitive
 * references in the tree, out[5] containment violations (must be 0), out[6] primitives referenced more or less
 * than once (must be 0), out[7] 1 if flux_set_scene would use the BVH in auto mode else 0. */
int flux_bvh_describe(const flux_scene_flat *scene, uint64_t out[8]);
/* FNV-1a hash of everything the traversal would read for `scene` (nodes, primitive references, linear list, leaf
 * records, depth, leaf size, extent).  Host only.  The builder is deterministic — the same scene gives the same
 * tree whatever the number of host threads that built it — and the tests hold it to that. */
int flux_bvh_hash(const flux_scene_flat *scene, uint64_t *hash);

/* Force the render kernel variant: 0 = auto, 1 = direct (lane group per pixel), 2 = regeneration
 * (warp per pixel with in-warp path regeneration; needs spp >= 64 and a sphere/plane scene),
 * 4 = block-local wavefront (CTA per pixel, path slots in shared memory, conservative FP32 box pre-test,
 * material-sorted shading and regeneration; needs spp >= 256, depth <= 8; the auto choice when it applies);
 * 3 was its first generation, removed in round 2 and refused with FLUX_ERR_INVALID.  All compute the
 * same per-sample radiance; they differ only in the order of the per-pixel sum (last bits).  Used by parity
 * tests and A/B timing. */
int flux_set_kernel_mode(flux_ctx *ctx, int mode);

/* Enable (default) / disable the glossy lobe table: to_unit_hemi(pixel sample, reflect_exponent)
 * (brdf.rs:64, samplers/src/lib.rs:133-142) tabulated per sample and distinct exponent when the sample
 * sets are installed, instead of being evaluated at every glossy bounce.  Same device function, same bits.
 * Takes effect at the next flux_set_scene / flux_set_samples / flux_generate_samples. */
int flux_set_glossy_table(flux_ctx *ctx, int enable);

/* Unfused FP64 issue-rate microbenchmark (the roofline denominator of
 * SURVEY.md §8d / H8): returns 1e9 FP64 instr/s for dependent DADD/DMUL chains
 * over the whole device. */
int flux_measure_fp64_peak(flux_ctx *ctx, double *ginstr_per_s);

/* ---- output file (N2): Image::write, fluxcore/src/image.rs:42-60 --------- */
int flux_write_ppm(const char *path, uint32_t width, uint32_t height, const double *rgb);

#ifdef __cplusplus
}
#endif
#endif /* FLUXB200_H */
