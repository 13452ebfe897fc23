// flux_scene.cuh — device-resident scene, sample sets and kernel parameters.
//
// Layout in HBM (DESIGN.md "data layout"):
//   spheres   SoA of 12 f64 planes [SPH_FIELDS][n_spheres] + u32 meta [2][n]
//   planes    SoA of 6  f64 planes [PLN_FIELDS][n_planes]  + u32 meta [2][n]
//   triangles SoA of 9  f64 planes [TRI_FIELDS][n_tris]    + u32 meta [2][n]  (EXTENSION)
//   materials AoS DevMaterial[n_materials]
//   samples   reference layout: pixel [set][i]{x,y}, disc [set][i]{x,y},
//             hemi [set][depth][i]{x,y,z}   (sampling.rs:5-10)
//   set index u32 [H][W]                     (trace.rs:64,68-69)
#pragma once
#include "../../include/fluxb200.h"
#include "flux_math.cuh"

enum { SPH_CX = 0, SPH_CY, SPH_CZ, SPH_R, SPH_RR, SPH_INV, SPH_C0X, SPH_C0Y, SPH_C0Z, SPH_C1X, SPH_C1Y, SPH_C1Z, SPH_FIELDS };
enum { PLN_PX = 0, PLN_PY, PLN_PZ, PLN_NX, PLN_NY, PLN_NZ, PLN_FIELDS };
enum { TRI_V0X = 0, TRI_V0Y, TRI_V0Z, TRI_E1X, TRI_E1Y, TRI_E1Z, TRI_E2X, TRI_E2Y, TRI_E2Z, TRI_FIELDS };
enum { KIND_SPHERE = 0, KIND_PLANE = 1, KIND_TRI = 2 };

struct DevMaterial {
    uint32_t kind;   // FLUX_MAT_*
    uint32_t gidx;   // glossy: index of this material's exponent in the lobe table (DevSamples::ghemi)
    double c[3];     // matte: (cd*kd)*INV_PI; emissive: color*power; reflective/glossy: cs*ks
    double exp;      // glossy exponent
    double inv_e1;   // 1.0/(exp+1.0), samplers/src/lib.rs:135
};

struct DevCamera {
    V3 eye, u, v, w;
    double aps;        // pixel_size / zoom_factor            trace.rs:60
    double half_w;     // W as f64 * 0.5                      trace.rs:57
    double half_h;     // H as f64 * 0.5                      trace.rs:56
    double factor;     // focal_distance/view_plane_distance  trace.rs:45
    double focal;      // focal_distance
    double lens_radius;
    V3 focal_w;        // focal_distance * w                  trace.rs:50
    double bg[3];
    uint32_t W, H;
    uint32_t max_depth;
};

struct DevScene {
    uint32_t n_spheres, n_planes, n_tris, n_materials;
    const double *sph;          // [SPH_FIELDS][n_spheres]
    const uint32_t *sph_meta;   // [2][n_spheres]: shape_id, material
    const double *pln;
    const uint32_t *pln_meta;
    const double *tri;
    const uint32_t *tri_meta;
    const DevMaterial *materials;
    // BVH over sphere/triangle boxes (EXTENSION, flux_bvh.cuh; unused when the linear scan is selected)
    const void *bvh_nodes;      // BvhNode4[bvh_n_nodes]
    unsigned long long bvh_tex; // the same nodes as a linear texture of uint4 texels (8 per node); 0 = none
    const uint32_t *bvh_prims;  // leaf primitive refs: (kind<<30 | index)
    const void *bvh_sph;        // SphRec[n_spheres]
    const void *bvh_tri;        // TriRec[n_tris]
    const uint32_t *bvh_linear; // spheres kept out of the tree (oversized / non-finite), ascending
    uint32_t bvh_n_linear;
    uint32_t bvh_n_nodes;
    uint32_t use_bvh;
    uint32_t bvh_tree_spheres;  // spheres inside the tree (0: only triangles there; the spheres, if any, are on the linear list)
    double bvh_extent;          // max |coordinate| of the boxes in the tree (scale of the pruning margin)
};

struct DevSamples {
    uint32_t root, n, max_depth, num_sets;
    const double2 *pixel;   // [set][i]
    const double2 *disc;    // [set][i]
    const double *hemi;     // [set][depth][i][3]
    // glossy lobe table: to_unit_hemi(pixel_sets[set][i], exp_k) for the K distinct glossy exponents of the
    // scene, [set][k][i][3]; null when not built (then the kernels evaluate it inline)
    const double *ghemi;
    uint32_t gk;
};

#define FLUX_CULL_MAX 128  // spheres covered by the constant-bank FP32 table (render_wave2.cu)
#ifndef FLUX_WAVE2_LINEAR_MAX
#define FLUX_WAVE2_LINEAR_MAX 128  // up to here render_wave2.cu scans linearly even when a BVH exists (api.cu); 112 in r1, when the two
                                   // met at ~120 spheres; with the primary mask the linear scan leads at 124: 2639 vs 2375 Msamples/s (r2AJ)
#endif

struct RenderParams {
    DevScene scene;
    DevCamera cam;
    DevSamples ss;
    const uint32_t *set_index;  // [H][W]
    const uint32_t *rows;       // [n_rows] image rows to render, ascending
    uint32_t n_rows;
    double *out;                // [n_rows][W][3], or with out_by_row the whole frame [H][W][3] (flux_frame)
    uint32_t out_by_row;        // 1: the pixel of image row r, column c is stored at out[(r * W + c) * 3]
    unsigned long long *counters;  // flux_counters as u64[...] or null
    unsigned int *work_counter;    // dynamic work distribution
    // spheres rounded to f32, {centre x, y, z, radius}, read as constant-bank operands by the conservative slab
    // pre-test of render_wave2.cu; NaN for spheres that must always take the exact test (negative radius,
    // non-finite) and for the unused entries.  cull_cmax >= max_k |centre_k| + radius over the valid spheres.
    float cull[FLUX_CULL_MAX][4];
    float cull_cmax;
    // 1: pixel samples lie in [0,1]^2 and lens samples in the unit disc (always so for device-generated sets; checked on
    // upload for caller-supplied ones) — what the per-pixel primary-ray mask of render_wave2.cu is derived from
    uint32_t primary_mask_ok;
    // progressive passes (flux_progressive_pass, render.cu only): samples [i_begin, i_end) of every pixel; when
    // `accum` is set their radiance sum is added to accum[n_rows][W][3] instead of being averaged into `out`
    uint32_t i_begin, i_end;
    double *accum;
};

// indices into flux_counters viewed as u64[]
enum {
    CN_SAMPLES = 0, CN_SEGMENTS, CN_BBOX_TESTS, CN_BBOX_PASS, CN_DISC_NONNEG, CN_T2, CN_PLANE_TESTS,
    CN_TRI_TESTS, CN_CANDIDATES, CN_HIT_SPHERE, CN_HIT_PLANE, CN_HIT_TRI, CN_EMISSIVE, CN_MATTE,
    CN_SPECULAR, CN_GLOSSY, CN_GLOSSY_FLIP, CN_DEPTH_CUT, CN_MISS, CN_NODES, CN_COUNT
};
