// fluxhost.hpp — C++17 host side above the C-ABI of include/fluxb200.h.
//
// The reference's host is Rust (fluxcore + the flux binary); this image has no Rust toolchain, so the host
// that is built and tested here is C++ and mirrors the reference's interface for the render path by name
// and meaning:
//
//   fluxcore/src/scene.rs:42-74    SceneData, OutputSettings, CameraSettings, CameraData, ShapeData
//   fluxcore/src/shapes.rs:18-83   SphereData, PlaneData, MaterialData::{Matte,Emissive,Reflective,GlossyReflective}
//   fluxcore/src/job.rs:40-88      JobConfiguration, WorkUnit, Job::work_units
//   fluxcore/src/manager.rs:25-33  WorkUnitResult, WorkerInfo;  :232-236 trait Worker -> GpuWorker
//   fluxcore/src/workers.rs:46-64  the LocalWorker loop body   -> GpuWorker::run_job
//   fluxcore/src/trace.rs:26,53    Camera::new, Camera::render
//   fluxcore/src/image.rs:5-60     Image, Image::write (P3 PPM, maxval 65535)
//   flux/src/main.rs:28-29         serde_yaml::from_reader     -> SceneData::from_yaml_file
//
// Triangle / Mesh / Rectangle / Box are EXTENSIONS of ShapeData (the reference has Sphere and Plane only).
// Everything that computes pixels is behind libfluxb200.so; there is no CPU rendering path in this file.
#pragma once
#include <array>
#include <atomic>
#include <cstdint>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <variant>
#include <vector>

#include "../include/fluxb200.h"

namespace flux {

using Vec3 = std::array<double, 3>;

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// ---- MaterialData, shapes.rs:42-83 ----
struct MatteData { Vec3 diffuse_color, ambient_color; double diffuse_coefficient; };
struct EmissiveData { Vec3 color; double power; };
struct ReflectiveData { double reflect_amount; Vec3 reflect_color; };
struct GlossyReflectiveData { double reflect_amount; Vec3 reflect_color; double reflect_exponent; };
using MaterialData = std::variant<MatteData, EmissiveData, ReflectiveData, GlossyReflectiveData>;

// ---- ShapeData, scene.rs:71-74 + shapes.rs:18-37 (+ extensions) ----
struct SphereData { Vec3 center; double radius; MaterialData material; bool invert; };
struct PlaneData { Vec3 point, normal; MaterialData material; };
struct TriangleData { Vec3 v0, v1, v2; MaterialData material; };                                    // EXTENSION
struct MeshData { std::vector<Vec3> vertices; std::vector<std::array<int64_t, 3>> faces; MaterialData material; };  // EXTENSION
struct RectangleData { Vec3 corner, edge_a, edge_b; MaterialData material; };                       // EXTENSION
struct BoxData { Vec3 min, max; MaterialData material; };                                           // EXTENSION
using ShapeData = std::variant<SphereData, PlaneData, TriangleData, MeshData, RectangleData, BoxData>;

struct OutputSettings { uint32_t image_width, image_height; double pixel_size; };   // scene.rs:58-63
struct CameraSettings { Vec3 eye, look_at, up; };                                  // scene.rs:11-16
struct CameraData { double zoom_factor, view_plane_distance, focal_distance, lens_radius; };  // scene.rs:50-56

struct JobConfiguration {          // job.rs:49-53; defaults flux/src/main.rs:20-21,172
    uint32_t sample_root = 1;
    uint32_t max_trace_depth = 5;
    uint32_t rows_per_work_unit = 50;
};

// job.rs:40-44, row_end inclusive.  JobID is (allocator id, serial) in the reference (job.rs:10): job_id is the
// serial, job_allocator_id the allocator's random id; both travel back unchanged in the WorkUnitResult.
struct WorkUnit { uint32_t row_start, row_end; uint64_t job_id; uint64_t job_allocator_id = 0; };
struct WorkUnitResult {                                             // manager.rs:25-28
    WorkUnit work_unit;
    std::vector<double> rows;   // [row_end-row_start+1][W][3], linear RGB, averaged and max_to_one-clamped
};
struct WorkerInfo { std::string name; uint32_t num_threads; };      // manager.rs:221-224 (num_threads) + a display name that never travels

// Owns the arrays a flux_scene_flat points into.
struct FlatScene {
    flux_scene_flat flat{};
    std::vector<flux_material> materials;
    std::vector<double> sphere_center, sphere_radius, plane_point, plane_normal, tri_v0, tri_v1, tri_v2;
    std::vector<uint8_t> sphere_invert;
    std::vector<uint32_t> sphere_shape_id, sphere_material, plane_shape_id, plane_material, tri_shape_id, tri_material;
    uint32_t n_shapes = 0;
    const flux_scene_flat *ptr();
};

struct SceneData {   // scene.rs:42-49
    std::string scene_name;
    OutputSettings output_settings{};
    Vec3 background{};
    std::vector<ShapeData> shapes;
    CameraSettings camera_settings{};
    CameraData camera_data{};

    // serde_yaml::from_reader (flux/src/main.rs:28-29): every field required, unknown keys ignored,
    // anchors/aliases resolved, Vector3/Point3/Color accepted as 3-element sequences.
    static SceneData from_yaml_file(const std::string &path);
    static SceneData from_yaml_string(const std::string &text);
    SceneData with_size(uint32_t width, uint32_t height) const;   // BASELINE config 1 (no CLI override in the reference)
    // plain data -> per-kind arrays with one shared shape-id space (include/fluxb200.h flux_scene_flat)
    std::unique_ptr<FlatScene> flatten() const;
};

// The flattened scene as text, doubles in hex-float: what the tests compare with the Python mirror's flattening.
void dump_flat_text(const std::string &path, FlatScene &f);

// Job::work_units (job.rs:66-88) without its dropped-trailing-row quirk: covers every row.
std::vector<WorkUnit> work_units(uint32_t image_height, uint32_t rows_per_work_unit, uint64_t job_id = 0);

// Image, image.rs:5-60.
struct Image {
    uint32_t width = 0, height = 0;
    std::vector<double> pixels;   // [H][W][3]; rows never set stay black (image.rs:54-58)
    Image(uint32_t w, uint32_t h) : width(w), height(h), pixels((size_t)w * h * 3, 0.0) {}
    void set_rows(const WorkUnitResult &r);          // ImageBuilder: manager.rs:316-325
    void write(const std::string &path) const;       // image.rs:42-60 via flux_write_ppm
};

// RAII flux_ctx; throws flux::Error with flux_last_error's text.
class GpuContext {
  public:
    explicit GpuContext(int device);
    ~GpuContext();
    GpuContext(const GpuContext &) = delete;
    GpuContext &operator=(const GpuContext &) = delete;
    flux_ctx *get() const { return ctx_; }
    int device() const { return device_; }
    void check(int rc, const char *what) const;

  private:
    flux_ctx *ctx_ = nullptr;
    int device_ = 0;
};

// Scene::from_data (scene.rs:128-154): plain data + job configuration, flattened for the device.
struct Scene {
    SceneData data;
    JobConfiguration job_config;
    std::unique_ptr<FlatScene> flat;
    static Scene from_data(const SceneData &sd, const JobConfiguration &cfg);
};

// Camera::new + Camera::render (trace.rs:26-42, 53-97) on one GPU.  num_sets is the number of sample sets
// (the reference passes image_width, workers.rs:50); the sets are generated on the device from `seed`
// because the reference's come from an unseeded RNG.
class Camera {
  public:
    static Camera create(GpuContext &ctx, const Scene &scene, const JobConfiguration &cfg, uint32_t num_sets, uint64_t seed);
    WorkUnitResult render(const Scene &scene, const WorkUnit &unit) const;
    std::vector<double> render_row_list(const std::vector<uint32_t> &rows) const;
    // Progressive refinement of one work unit: passes of `batch` samples per pixel; after each
    // pass `on_pass(samples_done, result_so_far)` is called and returns false to cancel (JobHandle::cancel,
    // manager.rs:66-69, between passes instead of between work units).  Returns the last result.
    WorkUnitResult render_progressive(const WorkUnit &unit, uint32_t sample_root, uint32_t batch,
                                      const std::function<bool(uint32_t, const WorkUnitResult &)> &on_pass) const;
    float last_kernel_ms() const;

  private:
    Camera(GpuContext &ctx, uint32_t w, uint32_t h) : ctx_(&ctx), width_(w), height_(h) {}
    GpuContext *ctx_;
    uint32_t width_, height_;
};

// Third Worker beside LocalWorker / NetworkWorker (manager.rs:232-236).  One GpuWorker drives `devices`
// GPUs of one box: interleaved tiles of `tile_rows` rows per GPU (flux_shard_rows) replace the shared
// bounded(1) work queue (manager.rs:100); each GPU renders its shard on its own host thread and the rows
// are assembled in host memory.
class GpuWorker {
  public:
    GpuWorker(std::vector<int> devices, uint64_t seed, uint32_t tile_rows = 1);
    WorkerInfo info() const;
    // workers.rs:46-64 for one job: Scene::from_data, Camera::new, render every work unit -> image
    Image render_job(const SceneData &sd, const JobConfiguration &cfg, double *render_seconds = nullptr);
    // the single-GPU loop body itself, unit by unit (what a manager would drive).  `cancel` mirrors
    // JobHandle::cancel / CancellableIterator (manager.rs:66-69,365-393): once set, no further unit is issued; the
    // unit in flight finishes; the rows rendered so far are returned (missing rows stay black in an Image).
    std::vector<WorkUnitResult> run_job(const SceneData &sd, const JobConfiguration &cfg, const std::atomic<bool> *cancel = nullptr);
    // the same loop body driven from outside, the way a manager drives a worker over its unit channel
    // (workers.rs:46-71; flux-node/src/main.rs:60-75): begin_job = Scene::from_data + Camera::new on every GPU,
    // render_unit = camera.render(&scene, unit) with the unit's rows sharded over the GPUs in interleaved tiles.
    // Progressive refinement of the whole frame: passes of `batch` samples per pixel on every
    // GPU's row shard; after each pass on_pass(samples_done, image_so_far) is called and returns false to cancel
    // (the preview's Esc -> JobHandle::cancel, flux/src/main.rs:288-293).  Returns the last image.
    Image render_job_progressive(const SceneData &sd, const JobConfiguration &cfg, uint32_t batch,
                                 const std::function<bool(uint32_t, const Image &)> &on_pass);
    void begin_job(const SceneData &sd, const JobConfiguration &cfg);
    bool has_job() const { return job_ != nullptr; }
    WorkUnitResult render_unit(const WorkUnit &unit);
    uint32_t job_image_width() const { return job_ ? job_->width : 0; }

  private:
    struct ActiveJob {
        Scene scene;
        std::vector<Camera> cameras;   // one per GPU, same seed: identical sample sets everywhere
        uint32_t width = 0, height = 0;
    };
    std::unique_ptr<ActiveJob> job_;
    std::vector<int> devices_;
    std::vector<std::unique_ptr<GpuContext>> contexts_;   // created with the worker, like LocalWorker::new builds its pool
    uint64_t seed_;
    uint32_t tile_rows_;
};

}  // namespace flux
