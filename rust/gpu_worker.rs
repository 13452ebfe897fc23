//! gpu_worker.rs — `GpuWorker`, a third `Worker` (fluxcore/src/manager.rs:232-236) beside `LocalWorker`
//! (workers.rs:14-104) and `NetworkWorker` (workers.rs:112-260), rendering on one B200 through `fluxb200-sys`.
//!
//! UNCOMPILED: this repository's image has no Rust toolchain.  The same call sequence is compiled and tested in C++
//! (host/fluxhost.cpp: `GpuWorker::{begin_job, render_unit, render_job}`) and in Python (flux_b200/worker.py).
//!
//! Drop it into `fluxcore/src/` as `gpu_worker.rs`, add `pub mod gpu_worker;` to `fluxcore/src/lib.rs`,
//! `fluxb200-sys = { path = ".../rust/fluxb200-sys" }` to `fluxcore/Cargo.toml`, and in `flux/src/main.rs:43-60`
//! push `GpuWorker::new(device, seed)` — one per visible GPU — into `worker_handles` next to (or instead of)
//! `LocalWorker::new(..)`.  Nothing else in flux changes: the manager keeps sending `WorkUnit`s over the shared
//! `bounded(1)` channel (manager.rs:100) and receiving `RenderEvent::RowsReady` (manager.rs:156-162).
//!
//! The thread body below is `LocalWorker`'s (workers.rs:42-78) with its three hot-path calls replaced:
//!
//! | reference (workers.rs)                                   | here                                              |
//! |-----------------------------------------------------------|---------------------------------------------------|
//! | `Scene::from_data(job.scene_data, job.config)`  :46       | `FlatScene::from_data` + `flux_set_scene`         |
//! | `Camera::new(.., job.config, image_width, ..)`  :47-54    | `flux_generate_samples(ctx, seed, image_width)`   |
//! | `camera.render(&scene, unit)`                   :60       | `flux_render_rows(ctx, row_start, row_end, buf)`  |
use std::ffi::CStr;
use std::os::raw::c_int;
use std::thread;

use crossbeam::channel::{unbounded, Receiver, Sender};

use fluxb200_sys::*;

use crate::color::Color;
use crate::manager::{RenderEvent, Worker, WorkerHandle, WorkerInfo, WorkerRequest, WorkUnitResult};
use crate::scene::{SceneData, ShapeData};
use crate::shapes::MaterialData;

/// `SceneData` (scene.rs:42-66) flattened into the per-kind arrays of `flux_scene_flat`.  The position of a shape
/// in `scene_data.shapes` is its shape id: closest-hit ties go to the earlier shape (scene.rs:156-160 +
/// common.rs:17-23), and spheres and planes share that one index space.  The vectors own what `raw` points into.
pub struct FlatScene {
    pub raw: flux_scene_flat,
    _materials: Vec<flux_material>,
    _sphere_center: Vec<f64>,
    _sphere_radius: Vec<f64>,
    _sphere_invert: Vec<u8>,
    _sphere_shape_id: Vec<u32>,
    _sphere_material: Vec<u32>,
    _plane_point: Vec<f64>,
    _plane_normal: Vec<f64>,
    _plane_shape_id: Vec<u32>,
    _plane_material: Vec<u32>,
}

fn material(m: &MaterialData) -> flux_material {
    // shapes.rs:42-83 -> flux_material: colour, one coefficient, one exponent
    match m {
        MaterialData::Matte(d) => flux_material {
            kind: FLUX_MAT_MATTE, _pad: 0, color: [d.diffuse_color.r, d.diffuse_color.g, d.diffuse_color.b],
            k: d.diffuse_coefficient, exp: 0.0,   // ambient_color is never read on the path (materials.rs:19-33)
        },
        MaterialData::Emissive(d) => flux_material {
            kind: FLUX_MAT_EMISSIVE, _pad: 0, color: [d.color.r, d.color.g, d.color.b], k: d.power, exp: 0.0,
        },
        MaterialData::Reflective(d) => flux_material {
            kind: FLUX_MAT_REFLECTIVE, _pad: 0, color: [d.reflect_color.r, d.reflect_color.g, d.reflect_color.b],
            k: d.reflect_amount, exp: 0.0,
        },
        MaterialData::GlossyReflective(d) => flux_material {
            kind: FLUX_MAT_GLOSSY, _pad: 0, color: [d.reflect_color.r, d.reflect_color.g, d.reflect_color.b],
            k: d.reflect_amount, exp: d.reflect_exponent,
        },
    }
}

impl FlatScene {
    pub fn from_data(sd: &SceneData) -> FlatScene {
        let mut materials = Vec::new();
        let (mut sc, mut sr, mut si, mut sid, mut sm) = (Vec::new(), Vec::new(), Vec::new(), Vec::new(), Vec::new());
        let (mut pp, mut pn, mut pid, mut pm) = (Vec::new(), Vec::new(), Vec::new(), Vec::new());
        for (shape_id, shape) in sd.shapes.iter().enumerate() {
            match shape {
                ShapeData::Sphere(s) => {
                    sc.extend_from_slice(&[s.center.x, s.center.y, s.center.z]);
                    sr.push(s.radius);
                    si.push(s.invert as u8);
                    sid.push(shape_id as u32);
                    sm.push(materials.len() as u32);
                    materials.push(material(&s.material));
                }
                ShapeData::Plane(p) => {
                    pp.extend_from_slice(&[p.point.x, p.point.y, p.point.z]);
                    pn.extend_from_slice(&[p.normal.x, p.normal.y, p.normal.z]);   // used as given (shapes.rs:137-147)
                    pid.push(shape_id as u32);
                    pm.push(materials.len() as u32);
                    materials.push(material(&p.material));
                }
            }
        }
        let o = &sd.output_settings;
        let (cs, cd) = (&sd.camera_settings, &sd.camera_data);
        let raw = flux_scene_flat {
            image_width: o.image_width as u32,
            image_height: o.image_height as u32,
            pixel_size: o.pixel_size,
            background: [sd.background.r, sd.background.g, sd.background.b],
            eye: [cs.eye.x, cs.eye.y, cs.eye.z],
            look_at: [cs.look_at.x, cs.look_at.y, cs.look_at.z],
            up: [cs.up.x, cs.up.y, cs.up.z],
            zoom_factor: cd.zoom_factor,
            view_plane_distance: cd.view_plane_distance,
            focal_distance: cd.focal_distance,
            lens_radius: cd.lens_radius,
            n_materials: materials.len() as u32,
            materials: materials.as_ptr(),
            n_spheres: sr.len() as u32,
            sphere_center: sc.as_ptr(),
            sphere_radius: sr.as_ptr(),
            sphere_invert: si.as_ptr(),
            sphere_shape_id: sid.as_ptr(),
            sphere_material: sm.as_ptr(),
            n_planes: pid.len() as u32,
            plane_point: pp.as_ptr(),
            plane_normal: pn.as_ptr(),
            plane_shape_id: pid.as_ptr(),
            plane_material: pm.as_ptr(),
            n_triangles: 0,                       // the reference has only Sphere and Plane (scene.rs:71-74)
            tri_v0: std::ptr::null(),
            tri_v1: std::ptr::null(),
            tri_v2: std::ptr::null(),
            tri_shape_id: std::ptr::null(),
            tri_material: std::ptr::null(),
        };
        // moving a Vec does not move its heap buffer: the pointers in `raw` stay valid for the life of the struct
        FlatScene {
            raw, _materials: materials, _sphere_center: sc, _sphere_radius: sr, _sphere_invert: si, _sphere_shape_id: sid,
            _sphere_material: sm, _plane_point: pp, _plane_normal: pn, _plane_shape_id: pid, _plane_material: pm,
        }
    }
}

/// The reference `unwrap()`s / panics around its hot path (workers.rs:159, manager.rs:161); so does this.
fn check(ctx: *mut flux_ctx, rc: c_int, what: &str) {
    if rc != FLUX_OK {
        let msg = unsafe { CStr::from_ptr(flux_last_error(ctx)) }.to_string_lossy().into_owned();
        panic!("fluxb200: {} failed ({}): {}", what, rc, msg);
    }
}

pub struct GpuWorker {
    sender: Sender<WorkerRequest>,
    thread_handle: thread::JoinHandle<()>,
    worker_info: WorkerInfo,
}

impl GpuWorker {
    /// `device`: CUDA device ordinal.  `seed`: seed of the device-generated sample sets — the reference seeds its
    /// sampler from `thread_rng` (samplers/src/lib.rs:27-33), which no two runs share; here a job is reproducible.
    pub fn new(device: i32, seed: u64) -> Self {
        let (s, r): (Sender<WorkerRequest>, Receiver<WorkerRequest>) = unbounded();

        let handle = thread::Builder::new().name(format!("GpuWorker{}", device)).spawn(move || {
            let mut ctx: *mut flux_ctx = std::ptr::null_mut();
            let rc = unsafe { flux_ctx_create(device as c_int, &mut ctx) };
            if rc != FLUX_OK {
                let msg = unsafe { CStr::from_ptr(flux_last_error(std::ptr::null())) }.to_string_lossy().into_owned();
                panic!("fluxb200: no usable CUDA device {} ({}); the library has no CPU fallback", device, msg);
            }

            'main: while let Ok(Some((job, recv_unit, send_result, wg))) = r.recv() {
                // Scene::from_data (workers.rs:46)
                let flat = FlatScene::from_data(&job.scene_data);
                let cfg = flux_job_config {
                    sample_root: job.config.sample_root as u32,
                    max_trace_depth: job.config.max_trace_depth as u32,
                    rows_per_work_unit: job.config.rows_per_work_unit as u32,
                };
                check(ctx, unsafe { flux_set_scene(ctx, &flat.raw, &cfg) }, "flux_set_scene");
                // Camera::new (workers.rs:47-54): MasterSampleSets::new with num_sets = image_width (workers.rs:50)
                let width = job.scene_data.output_settings.image_width;
                check(ctx, unsafe { flux_generate_samples(ctx, seed, width as u32) }, "flux_generate_samples");

                while let Ok(unit) = recv_unit.recv() {
                    // camera.render(&scene, unit) (workers.rs:60; trace.rs:53-97)
                    let n_rows = unit.row_end - unit.row_start + 1;
                    let mut buf = vec![0f64; n_rows * width * 3];
                    check(ctx, unsafe { flux_render_rows(ctx, unit.row_start as u32, unit.row_end as u32, buf.as_mut_ptr()) },
                          "flux_render_rows");
                    let rows: Vec<Vec<Color>> = buf.chunks(width * 3)
                        .map(|row| row.chunks(3).map(|c| Color { r: c[0], g: c[1], b: c[2] }).collect())
                        .collect();
                    let ev = RenderEvent::RowsReady(WorkUnitResult { work_unit: unit, rows });
                    match send_result.send(Some(ev)) {
                        Ok(()) => (),
                        Err(_) => continue 'main,     // as LocalWorker: advance to the next job (workers.rs:64-69)
                    }
                }

                drop(wg);
            }

            unsafe { flux_ctx_destroy(ctx) };
        }).unwrap();

        Self {
            sender: s,
            thread_handle: handle,
            worker_info: WorkerInfo { num_threads: 1 },   // one render thread = one GPU (manager.rs:221-224)
        }
    }
}

impl Worker for GpuWorker {
    fn handle(&self) -> WorkerHandle {
        WorkerHandle::new(self.sender.clone())
    }

    fn stop(self) {
        self.sender.send(None).ok();
        self.thread_handle.join().ok();
    }

    fn info(&self) -> WorkerInfo {
        self.worker_info
    }
}
