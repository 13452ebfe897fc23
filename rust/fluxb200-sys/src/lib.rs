//! Raw declarations of `include/fluxb200.h` (the C-ABI of `libfluxb200.so`), one for one.
//!
//! UNCOMPILED: this repository's image has no Rust toolchain.  `tests/test_rust_binding.py` parses the header and
//! this file and checks that every function is declared here with the same number of parameters, and that the
//! `#[repr(C)]` structs list the header's fields in the header's order.
//!
//! The reference enters its hot path at `fluxcore/src/workers.rs:46-64`
//! (`Scene::from_data`, `Camera::new`, `camera.render(&scene, unit)`); `rust/gpu_worker.rs` is the `Worker`
//! (`manager.rs:232-236`) that makes those three calls through the functions below.
#![allow(non_camel_case_types)]

use std::os::raw::{c_char, c_int, c_void};

pub const FLUX_OK: c_int = 0;
pub const FLUX_ERR_INVALID: c_int = 1;
pub const FLUX_ERR_CUDA: c_int = 2;
pub const FLUX_ERR_STATE: c_int = 3;
pub const FLUX_ERR_NO_DEVICE: c_int = 4;

pub const FLUX_MAT_MATTE: u32 = 0;
pub const FLUX_MAT_EMISSIVE: u32 = 1;
pub const FLUX_MAT_REFLECTIVE: u32 = 2;
pub const FLUX_MAT_GLOSSY: u32 = 3;

pub const FLUX_FRAME_HANDLE_BYTES: usize = 64;

#[repr(C)]
pub struct flux_ctx {
    _private: [u8; 0],
}

#[repr(C)]
pub struct flux_frame {
    _private: [u8; 0],
}

/// MaterialData variants, fluxcore/src/shapes.rs:42-83.
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct flux_material {
    pub kind: u32,
    pub _pad: u32,
    pub color: [f64; 3],
    pub k: f64,
    pub exp: f64,
}

/// Flattened SceneData (fluxcore/src/scene.rs:42-66) + the shapes of shapes.rs:18-37 as per-kind arrays.
#[repr(C)]
pub struct flux_scene_flat {
    pub image_width: u32,
    pub image_height: u32,
    pub pixel_size: f64,
    pub background: [f64; 3],
    pub eye: [f64; 3],
    pub look_at: [f64; 3],
    pub up: [f64; 3],
    pub zoom_factor: f64,
    pub view_plane_distance: f64,
    pub focal_distance: f64,
    pub lens_radius: f64,
    pub n_materials: u32,
    pub materials: *const flux_material,
    pub n_spheres: u32,
    pub sphere_center: *const f64,
    pub sphere_radius: *const f64,
    pub sphere_invert: *const u8,
    pub sphere_shape_id: *const u32,
    pub sphere_material: *const u32,
    pub n_planes: u32,
    pub plane_point: *const f64,
    pub plane_normal: *const f64,
    pub plane_shape_id: *const u32,
    pub plane_material: *const u32,
    pub n_triangles: u32,
    pub tri_v0: *const f64,
    pub tri_v1: *const f64,
    pub tri_v2: *const f64,
    pub tri_shape_id: *const u32,
    pub tri_material: *const u32,
}

/// JobConfiguration, fluxcore/src/job.rs:49-53.
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct flux_job_config {
    pub sample_root: u32,
    pub max_trace_depth: u32,
    pub rows_per_work_unit: u32,
}

/// Event counters of the algorithmic-op model.
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct flux_counters {
    pub samples: u64,
    pub segments: u64,
    pub bbox_tests: u64,
    pub bbox_pass: u64,
    pub disc_nonneg: u64,
    pub t2_evals: u64,
    pub plane_tests: u64,
    pub tri_tests: u64,
    pub candidates: u64,
    pub hit_sphere: u64,
    pub hit_plane: u64,
    pub hit_tri: u64,
    pub emissive: u64,
    pub matte: u64,
    pub specular: u64,
    pub glossy: u64,
    pub glossy_flip: u64,
    pub depth_cut: u64,
    pub miss: u64,
    pub nodes_visited: u64,
}

extern "C" {
    // ---- lifecycle ----
    pub fn flux_ctx_create(device: c_int, out: *mut *mut flux_ctx) -> c_int;
    pub fn flux_ctx_destroy(ctx: *mut flux_ctx) -> c_int;
    pub fn flux_last_error(ctx: *const flux_ctx) -> *const c_char;
    pub fn flux_version() -> *const c_char;

    // ---- job setup: Scene::from_data + Camera::new (scene.rs:128-154, trace.rs:26-42, sampling.rs:13-40) ----
    pub fn flux_set_scene(ctx: *mut flux_ctx, scene: *const flux_scene_flat, cfg: *const flux_job_config) -> c_int;
    pub fn flux_set_samples(ctx: *mut flux_ctx, sample_root: u32, max_depth: u32, num_sets: u32, pixel_xy: *const f64,
                            disc_xy: *const f64, hemi_xyz: *const f64) -> c_int;
    pub fn flux_generate_samples(ctx: *mut flux_ctx, seed: u64, num_sets: u32) -> c_int;
    pub fn flux_get_samples(ctx: *mut flux_ctx, pixel_xy: *mut f64, disc_xy: *mut f64, hemi_xyz: *mut f64) -> c_int;
    pub fn flux_get_set_index(ctx: *mut flux_ctx, idx: *mut u32) -> c_int;
    pub fn flux_set_set_index(ctx: *mut flux_ctx, idx: *const u32) -> c_int;

    // ---- the hot path: Camera::render (trace.rs:53-97) ----
    pub fn flux_render_rows(ctx: *mut flux_ctx, row_start: u32, row_end_inclusive: u32, out_rgb: *mut f64) -> c_int;
    pub fn flux_progressive_begin(ctx: *mut flux_ctx, rows: *const u32, n_rows: u32) -> c_int;
    pub fn flux_progressive_pass(ctx: *mut flux_ctx, sample_begin: u32, sample_end: u32, out_rgb: *mut f64) -> c_int;
    pub fn flux_render_rows_device(ctx: *mut flux_ctx, row_start: u32, row_end_inclusive: u32, d_out_rgb: *mut f64,
                                   cuda_stream: *mut c_void) -> c_int;
    pub fn flux_shard_rows(image_height: u32, tile_rows: u32, rank: u32, world: u32, rows: *mut u32, n_rows: *mut u32) -> c_int;
    pub fn flux_render_row_list(ctx: *mut flux_ctx, rows: *const u32, n_rows: u32, out_rgb: *mut f64) -> c_int;
    pub fn flux_render_row_list_device(ctx: *mut flux_ctx, rows: *const u32, n_rows: u32, d_out_rgb: *mut f64,
                                       cuda_stream: *mut c_void) -> c_int;

    // ---- multi-GPU frame assembly over NVLink peer memory (manager.rs:100,156-162 replaced) ----
    pub fn flux_frame_create(ctx: *mut flux_ctx, image_width: u32, image_height: u32, out: *mut *mut flux_frame) -> c_int;
    pub fn flux_frame_export(frame: *mut flux_frame, handle: *mut u8) -> c_int;
    pub fn flux_frame_open_ipc(ctx: *mut flux_ctx, handle: *const u8, image_width: u32, image_height: u32,
                               out: *mut *mut flux_frame) -> c_int;
    pub fn flux_frame_open_peer(ctx: *mut flux_ctx, owner_frame: *mut flux_frame, out: *mut *mut flux_frame) -> c_int;
    pub fn flux_render_row_list_into_frame(ctx: *mut flux_ctx, rows: *const u32, n_rows: u32, frame: *mut flux_frame,
                                           cuda_stream: *mut c_void) -> c_int;
    pub fn flux_ctx_sync(ctx: *mut flux_ctx) -> c_int;
    pub fn flux_frame_read(frame: *mut flux_frame, host_rgb: *mut f64) -> c_int;
    pub fn flux_frame_device_ptr(frame: *mut flux_frame, device_ptr: *mut *mut c_void) -> c_int;
    pub fn flux_frame_close(frame: *mut flux_frame) -> c_int;

    // ---- Scene::hit (scene.rs:156-160) on explicit rays ----
    pub fn flux_trace_rays(ctx: *mut flux_ctx, n: u64, origin_xyz: *const f64, dir_xyz: *const f64, hit_shape_id: *mut i32,
                           t: *mut f64) -> c_int;
    pub fn flux_trace_rays_device(ctx: *mut flux_ctx, n: u64, d_origin_xyz: *const f64, d_dir_xyz: *const f64,
                                  d_hit_shape_id: *mut i32, d_t: *mut f64, cuda_stream: *mut c_void) -> c_int;

    // ---- instrumentation ----
    pub fn flux_enable_counters(ctx: *mut flux_ctx, enable: c_int) -> c_int;
    pub fn flux_get_counters(ctx: *mut flux_ctx, out: *mut flux_counters) -> c_int;
    pub fn flux_reset_counters(ctx: *mut flux_ctx) -> c_int;
    pub fn flux_last_kernel_ms(ctx: *mut flux_ctx, ms: *mut f32) -> c_int;
    pub fn flux_launch_count(ctx: *mut flux_ctx, n: *mut u64) -> c_int;
    pub fn flux_set_accel_mode(ctx: *mut flux_ctx, mode: c_int) -> c_int;
    pub fn flux_bvh_describe(scene: *const flux_scene_flat, out: *mut u64) -> c_int;
    pub fn flux_bvh_hash(scene: *const flux_scene_flat, hash: *mut u64) -> c_int;
    pub fn flux_set_kernel_mode(ctx: *mut flux_ctx, mode: c_int) -> c_int;
    pub fn flux_set_glossy_table(ctx: *mut flux_ctx, enable: c_int) -> c_int;
    pub fn flux_measure_fp64_peak(ctx: *mut flux_ctx, ginstr_per_s: *mut f64) -> c_int;

    // ---- Image::write (image.rs:42-60) ----
    pub fn flux_write_ppm(path: *const c_char, width: u32, height: u32, rgb: *const f64) -> c_int;
}
