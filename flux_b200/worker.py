"""Host-side mirror of the reference's render-path interface over the C-ABI.

Reference shape (fluxcore/src/workers.rs:46-64), which ``GpuWorker`` follows::

    let scene  = Scene::from_data(job.scene_data, job.config);
    let camera = Camera::new(settings, basis, job.config, image_width, ...);
    loop { let r = camera.render(&scene, unit); send(RowsReady(r)) }

``Scene.from_data`` and ``Camera.new`` keep the reference's names and argument
meaning; ``Camera.render(scene, WorkUnit) -> WorkUnitResult`` returns the same
rows of linear RGB the reference returns (trace.rs:53-97, manager.rs:25-28).

Differences forced by the device and by reproducibility: the
sample sets are either generated on the GPU from an explicit ``seed`` or
supplied explicitly, because the reference's are drawn from an unseeded RNG.

No CPU fallback: constructing a ``GpuContext`` without the built library or
without a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List, Optional

import numpy as np

from . import _capi
from .scene import JobConfiguration, SceneData, WorkUnit, WorkUnitResult, work_units


class FluxError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"fluxb200 error {code}: {message}")
        self.code = code


class GpuContext:
    """Owns one ``flux_ctx`` (one GPU)."""

    def __init__(self, device: int = 0):
        self._lib = _capi.lib()
        self._ctx = C.c_void_p()
        rc = self._lib.flux_ctx_create(device, C.byref(self._ctx))
        if rc != 0:
            raise FluxError(rc, (self._lib.flux_last_error(None) or b"").decode())
        self.device = device

    def _ck(self, rc: int):
        if rc != 0:
            raise FluxError(rc, (self._lib.flux_last_error(self._ctx) or b"").decode())

    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.flux_ctx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- thin wrappers -----------------------------------------------------
    def set_scene(self, flat, cfg: JobConfiguration):
        jc = _capi.flux_job_config(cfg.sample_root, cfg.max_trace_depth, cfg.rows_per_work_unit)
        self._ck(self._lib.flux_set_scene(self._ctx, flat.ptr(), C.byref(jc)))

    def set_samples(self, root, max_depth, num_sets, pixel, disc, hemi):
        pixel = np.ascontiguousarray(pixel, np.float64)
        disc = np.ascontiguousarray(disc, np.float64)
        hemi = np.ascontiguousarray(hemi, np.float64)
        n = root * root
        if pixel.size != num_sets * n * 2 or disc.size != num_sets * n * 2 or hemi.size != num_sets * max_depth * n * 3:
            raise ValueError("sample arrays do not match (root, max_depth, num_sets)")
        self._ck(self._lib.flux_set_samples(self._ctx, root, max_depth, num_sets, _capi.as_dp(pixel),
                                            _capi.as_dp(disc), _capi.as_dp(hemi) if hemi.size else None))

    def generate_samples(self, seed: int, num_sets: int):
        self._ck(self._lib.flux_generate_samples(self._ctx, seed, num_sets))

    def get_samples(self, root, max_depth, num_sets):
        n = root * root
        pixel = np.empty((num_sets, n, 2)); disc = np.empty((num_sets, n, 2)); hemi = np.empty((num_sets, max_depth, n, 3))
        self._ck(self._lib.flux_get_samples(self._ctx, _capi.as_dp(pixel), _capi.as_dp(disc), _capi.as_dp(hemi)))
        return pixel, disc, hemi

    def set_set_index(self, idx):
        idx = np.ascontiguousarray(idx, np.uint32)
        self._ck(self._lib.flux_set_set_index(self._ctx, _capi.as_u32p(idx)))

    def get_set_index(self, height, width):
        idx = np.empty((height, width), np.uint32)
        self._ck(self._lib.flux_get_set_index(self._ctx, _capi.as_u32p(idx)))
        return idx

    def render_rows(self, row_start: int, row_end_inclusive: int, width: int) -> np.ndarray:
        if row_end_inclusive < row_start:
            raise FluxError(_capi.FLUX_ERR_INVALID, "render: row_end < row_start")
        out = np.empty((row_end_inclusive - row_start + 1, width, 3), np.float64)
        self._ck(self._lib.flux_render_rows(self._ctx, row_start, row_end_inclusive, _capi.as_dp(out)))
        return out

    def render_row_list(self, rows, width: int, out: Optional[np.ndarray] = None) -> np.ndarray:
        rows = np.ascontiguousarray(rows, np.uint32)
        if out is None:
            out = np.empty((rows.shape[0], width, 3), np.float64)
        self._ck(self._lib.flux_render_row_list(self._ctx, _capi.as_u32p(rows), rows.shape[0], _capi.as_dp(out)))
        return out

    def render_row_list_device(self, rows, d_out_ptr: int, stream_ptr: int = 0):
        rows = np.ascontiguousarray(rows, np.uint32)
        self._ck(self._lib.flux_render_row_list_device(self._ctx, _capi.as_u32p(rows), rows.shape[0],
                                                       C.c_void_p(d_out_ptr), C.c_void_p(stream_ptr)))

    # ---- multi-GPU frame assembly over NVLink peer memory (include/fluxb200.h, flux_frame_*) ----
    def frame_create(self, width: int, height: int) -> "Frame":
        h = C.c_void_p()
        self._ck(self._lib.flux_frame_create(self._ctx, width, height, C.byref(h)))
        return Frame(self._lib, h, width, height, owner=True)

    def frame_open_ipc(self, handle: bytes, width: int, height: int) -> "Frame":
        if len(handle) != _capi.FLUX_FRAME_HANDLE_BYTES:
            raise ValueError("frame handle must be 64 bytes")
        h = C.c_void_p()
        self._ck(self._lib.flux_frame_open_ipc(self._ctx, handle, width, height, C.byref(h)))
        return Frame(self._lib, h, width, height, owner=False)

    def frame_open_peer(self, owner: "Frame") -> "Frame":
        h = C.c_void_p()
        self._ck(self._lib.flux_frame_open_peer(self._ctx, owner._h, C.byref(h)))
        return Frame(self._lib, h, owner.width, owner.height, owner=False)

    def render_row_list_into_frame(self, rows, frame: "Frame", stream_ptr: int = 0):
        rows = np.ascontiguousarray(rows, np.uint32)
        self._ck(self._lib.flux_render_row_list_into_frame(self._ctx, _capi.as_u32p(rows), rows.shape[0], frame._h,
                                                           C.c_void_p(stream_ptr)))

    def sync(self):
        self._ck(self._lib.flux_ctx_sync(self._ctx))

    def progressive_begin(self, rows):
        rows = np.ascontiguousarray(rows, np.uint32)
        self._prog_shape = rows.shape[0]
        self._ck(self._lib.flux_progressive_begin(self._ctx, _capi.as_u32p(rows), rows.shape[0]))

    def progressive_pass(self, sample_begin: int, sample_end: int, width: int, want_image: bool = True):
        """Adds samples [sample_begin, sample_end) of every pixel; returns the image so far (or None)."""
        out = np.empty((self._prog_shape, width, 3), np.float64) if want_image else None
        self._ck(self._lib.flux_progressive_pass(self._ctx, sample_begin, sample_end, _capi.as_dp(out) if want_image else None))
        return out

    def trace_rays(self, origins, dirs):
        o = np.ascontiguousarray(origins, np.float64).reshape(-1, 3)
        d = np.ascontiguousarray(dirs, np.float64).reshape(-1, 3)
        if o.shape != d.shape:
            raise ValueError("origins and directions differ in shape")
        n = o.shape[0]
        hit = np.empty(n, np.int32); t = np.empty(n, np.float64)
        self._ck(self._lib.flux_trace_rays(self._ctx, n, _capi.as_dp(o), _capi.as_dp(d), _capi.as_i32p(hit), _capi.as_dp(t)))
        return hit, t

    def trace_rays_device(self, n, d_o, d_d, d_hit, d_t, stream_ptr: int = 0):
        self._ck(self._lib.flux_trace_rays_device(self._ctx, n, C.c_void_p(d_o), C.c_void_p(d_d), C.c_void_p(d_hit),
                                                  C.c_void_p(d_t), C.c_void_p(stream_ptr)))

    def enable_counters(self, on: bool):
        self._ck(self._lib.flux_enable_counters(self._ctx, 1 if on else 0))

    def counters(self) -> dict:
        cn = _capi.flux_counters()
        self._ck(self._lib.flux_get_counters(self._ctx, C.byref(cn)))
        return cn.as_dict()

    def reset_counters(self):
        self._ck(self._lib.flux_reset_counters(self._ctx))

    def last_kernel_ms(self) -> float:
        ms = C.c_float()
        self._ck(self._lib.flux_last_kernel_ms(self._ctx, C.byref(ms)))
        return ms.value

    def launch_count(self) -> int:
        n = C.c_uint64()
        self._ck(self._lib.flux_launch_count(self._ctx, C.byref(n)))
        return n.value

    def set_accel_mode(self, mode: int):
        self._ck(self._lib.flux_set_accel_mode(self._ctx, mode))

    def set_kernel_mode(self, mode: int):
        self._ck(self._lib.flux_set_kernel_mode(self._ctx, mode))

    def set_glossy_table(self, on: bool):
        self._ck(self._lib.flux_set_glossy_table(self._ctx, 1 if on else 0))

    def measure_fp64_peak(self) -> float:
        v = C.c_double()
        self._ck(self._lib.flux_measure_fp64_peak(self._ctx, C.byref(v)))
        return v.value


class Frame:
    """One framebuffer [H][W][3] f64 on its owner's GPU that every GPU of the box renders its rows into through
    peer memory (flux_frame_*; replaces the RowsReady stream into the manager's ImageBuilder, manager.rs:100,156-162)."""

    def __init__(self, lib, handle, width: int, height: int, owner: bool):
        self._lib, self._h, self.width, self.height, self.owner = lib, handle, width, height, owner

    def export(self) -> bytes:
        buf = C.create_string_buffer(_capi.FLUX_FRAME_HANDLE_BYTES)
        rc = self._lib.flux_frame_export(self._h, buf)
        if rc != 0:
            raise FluxError(rc, (self._lib.flux_last_error(None) or b"").decode())
        return buf.raw

    def device_ptr(self) -> int:
        p = C.c_void_p()
        rc = self._lib.flux_frame_device_ptr(self._h, C.byref(p))
        if rc != 0:
            raise FluxError(rc, "flux_frame_device_ptr")
        return int(p.value)

    def read(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        if out is None:
            out = np.empty((self.height, self.width, 3), np.float64)
        rc = self._lib.flux_frame_read(self._h, _capi.as_dp(out))
        if rc != 0:
            raise FluxError(rc, (self._lib.flux_last_error(None) or b"").decode())
        return out

    def close(self):
        if self._h:
            self._lib.flux_frame_close(self._h)
            self._h = None


def shard_rows(image_height: int, tile_rows: int, rank: int, world: int) -> np.ndarray:
    """Rows r with (r // tile_rows) % world == rank — flux_shard_rows (replaces the
    bounded(1) work queue, manager.rs:100).  Pure integer logic, mirrored here so
    the host side can plan gathers without a device."""
    if tile_rows <= 0 or world <= 0 or not (0 <= rank < world):
        raise ValueError("bad shard arguments")
    r = np.arange(image_height, dtype=np.uint32)
    return r[(r // tile_rows) % world == rank]


class Scene:
    """Scene::from_data (scene.rs:128-154): plain data + job configuration, flattened for the device."""

    def __init__(self, data: SceneData, config: JobConfiguration):
        self.data = data
        self.job_config = config
        self.output_settings = data.output_settings
        self.flat = data.flatten()

    @staticmethod
    def from_data(sd: SceneData, config: JobConfiguration) -> "Scene":
        return Scene(sd, config)


class Camera:
    """Camera::new + Camera::render (trace.rs:26-42, 53-97) on one GPU.

    ``num_sets`` is the number of sample sets (the reference passes image_width,
    workers.rs:50).  ``seed`` selects the device-generated sample sets; pass
    ``samples=(pixel, disc, hemi, set_index)`` to supply them explicitly.
    """

    def __init__(self, scene: Scene, config: JobConfiguration, num_sets: int, seed: int = 1,
                 samples=None, device: int = 0, ctx: Optional[GpuContext] = None):
        self.config = config
        self.ctx = ctx or GpuContext(device)
        self.width = scene.output_settings.image_width
        self.height = scene.output_settings.image_height
        self.ctx.set_scene(scene.flat, config)
        if samples is None:
            self.ctx.generate_samples(seed, num_sets)
        else:
            pixel, disc, hemi, set_index = samples
            self.ctx.set_samples(config.sample_root, config.max_trace_depth, num_sets, pixel, disc, hemi)
            self.ctx.set_set_index(set_index)

    @staticmethod
    def new(scene: Scene, config: JobConfiguration, num_sets: int, **kw) -> "Camera":
        return Camera(scene, config, num_sets, **kw)

    def render(self, scene: Scene, work: WorkUnit) -> WorkUnitResult:
        rows = self.ctx.render_rows(work.row_start, work.row_end, self.width)
        return WorkUnitResult(work, rows)

    def render_progressive(self, scene: Scene, work: WorkUnit, batch: int, cancel=None):
        """Progressive refinement of one work unit: yields ``(samples_done, WorkUnitResult)``
        after every pass of ``batch`` samples per pixel; the last one holds Camera::render's result (up to the
        order of the per-pixel sum).  ``cancel()`` returning True stops before the next pass, like
        JobHandle::cancel stops before the next work unit (manager.rs:66-69)."""
        if batch < 1:
            raise ValueError("batch must be >= 1")
        n = self.config.sample_root ** 2
        self.ctx.progressive_begin(np.arange(work.row_start, work.row_end + 1, dtype=np.uint32))
        done = 0
        while done < n:
            if cancel is not None and cancel():
                return
            end = min(n, done + batch)
            yield end, WorkUnitResult(work, self.ctx.progressive_pass(done, end, self.width))
            done = end


class GpuWorker:
    """Third ``Worker`` beside LocalWorker/NetworkWorker (manager.rs:232-236): given a job
    (scene data + configuration) and a stream of work units, yields RowsReady results."""

    def __init__(self, device: int = 0, seed: int = 1):
        self.ctx = GpuContext(device)
        self.seed = seed

    def info(self) -> dict:  # WorkerInfo, manager.rs:221-224 (num_threads) + a display name
        return {"name": f"gpu{self.ctx.device}", "num_threads": 1}

    def run_job(self, scene_data: SceneData, config: JobConfiguration, units: Optional[Iterable[WorkUnit]] = None,
                cancel=None):
        """Yields one WorkUnitResult per work unit.  ``cancel`` (a callable returning bool) mirrors
        JobHandle::cancel / CancellableIterator (manager.rs:66-69,365-393): once it returns True no further unit
        is issued; units already rendered have been yielded."""
        scene = Scene.from_data(scene_data, config)
        camera = Camera.new(scene, config, scene.output_settings.image_width, seed=self.seed, ctx=self.ctx)
        if units is None:
            units = work_units(scene.output_settings.image_height, config.rows_per_work_unit)
        for unit in units:
            if cancel is not None and cancel():
                return
            yield camera.render(scene, unit)

    def render_image(self, scene_data: SceneData, config: JobConfiguration) -> np.ndarray:
        h, w = scene_data.output_settings.image_height, scene_data.output_settings.image_width
        img = np.empty((h, w, 3), np.float64)
        for res in self.run_job(scene_data, config):
            img[res.work_unit.row_start:res.work_unit.row_end + 1] = res.rows
        return img

    def stop(self):
        self.ctx.close()
