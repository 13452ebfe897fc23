// flux_kernels.cuh — kernel entry points shared between translation units.
#pragma once
#include "flux_scene.cuh"

#define FLUX_MAX_DEPTH_CAP 32  // per-path (f, weight) stack entries
#define FLUX_LINEAR_LIMIT 40   // bounded shapes (spheres + triangles) up to which closest-hit scans linearly; measured r1:
                               // at 67 spheres the BVH kernel renders config 4 at 1.6 Gsamples/s, the linear wavefront at 1.1

// Camera::render for a list of rows (trace.rs:53-97).
void launch_render(const RenderParams &p, bool count, int sm_count, cudaStream_t stream);
void launch_resolve_accum(const double *accum, double *out, uint32_t npix, double inv_count, cudaStream_t stream);

// Same result, scheduled for high sample counts: warp per pixel, in-warp path regeneration,
// scene in shared memory (render_regen.cu).
bool regen_kernel_applicable(const RenderParams &p);
void launch_render_regen(const RenderParams &p, bool count, int sm_count, cudaStream_t stream);

// Block-local wavefront (render_wave2.cu): CTA per pixel, path slots in shared memory, conservative FP32 slab
// pre-test from the constant bank, terminated paths regenerated in the material-sorted item stage, 3 barriers per
// iteration.  (Its first generation, render_wave.cu — compacted (slot, sphere) candidate pairs, 8 barriers — was an
// A/B relic by the end of round 1 and left the library in round 2; kernel mode 3 is refused.)
bool wave2_kernel_applicable(const RenderParams &p);
void launch_render_wave2(const RenderParams &p, bool count, int sm_count, cudaStream_t stream);

// Scene::hit on an explicit ray batch (scene.rs:156-160).
void launch_trace_rays(const DevScene &sc, uint64_t n, const double *o, const double *d, int32_t *hit, double *t,
                       int sm_count, cudaStream_t stream, unsigned long long *counters, unsigned long long *work);

// MasterSampleSets::new on the device (sampling.rs:13-33) + per-row set-index permutations (sampling.rs:35-40).
void launch_generate_samples(uint64_t seed, uint32_t root, uint32_t max_depth, uint32_t num_sets, double2 *pixel,
                             double2 *disc, double *hemi, cudaStream_t stream);
void launch_generate_set_index(uint64_t seed, uint32_t H, uint32_t W, uint32_t num_sets, uint32_t *idx,
                               cudaStream_t stream);

void launch_build_glossy_table(const double2 *pixel, uint32_t n, uint32_t num_sets, uint32_t gk, const double *inv_e1,
                               double *ghemi, int sm_count, cudaStream_t stream);

// FP64 issue-rate microbenchmark; returns total FP64 instructions executed.
double launch_fp64_peak(int sm_count, int iters, double *sink, cudaStream_t stream);
