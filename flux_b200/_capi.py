"""ctypes binding of include/fluxb200.h (libfluxb200.so).

This is the same thin binding a Rust ``fluxb200-sys`` crate would make
(INTEGRATION.md); nothing here computes.  The library is loaded lazily and
loading fails loudly: there is no CPU fallback for the render path.
"""
from __future__ import annotations

import ctypes as C
import os

FLUX_OK, FLUX_ERR_INVALID, FLUX_ERR_CUDA, FLUX_ERR_STATE, FLUX_ERR_NO_DEVICE = range(5)
FLUX_MAT_MATTE, FLUX_MAT_EMISSIVE, FLUX_MAT_REFLECTIVE, FLUX_MAT_GLOSSY = range(4)

_dp = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)
_i32p = C.POINTER(C.c_int32)


class flux_material(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("_pad", C.c_uint32), ("color", C.c_double * 3),
                ("k", C.c_double), ("exp", C.c_double)]


class flux_scene_flat(C.Structure):
    _fields_ = [
        ("image_width", C.c_uint32), ("image_height", C.c_uint32), ("pixel_size", C.c_double),
        ("background", C.c_double * 3),
        ("eye", C.c_double * 3), ("look_at", C.c_double * 3), ("up", C.c_double * 3),
        ("zoom_factor", C.c_double), ("view_plane_distance", C.c_double),
        ("focal_distance", C.c_double), ("lens_radius", C.c_double),
        ("n_materials", C.c_uint32), ("materials", C.POINTER(flux_material)),
        ("n_spheres", C.c_uint32), ("sphere_center", _dp), ("sphere_radius", _dp),
        ("sphere_invert", _u8p), ("sphere_shape_id", _u32p), ("sphere_material", _u32p),
        ("n_planes", C.c_uint32), ("plane_point", _dp), ("plane_normal", _dp),
        ("plane_shape_id", _u32p), ("plane_material", _u32p),
        ("n_triangles", C.c_uint32), ("tri_v0", _dp), ("tri_v1", _dp), ("tri_v2", _dp),
        ("tri_shape_id", _u32p), ("tri_material", _u32p),
    ]


class flux_job_config(C.Structure):
    _fields_ = [("sample_root", C.c_uint32), ("max_trace_depth", C.c_uint32),
                ("rows_per_work_unit", C.c_uint32)]


COUNTER_FIELDS = ["samples", "segments", "bbox_tests", "bbox_pass", "disc_nonneg", "t2_evals",
                  "plane_tests", "tri_tests", "candidates", "hit_sphere", "hit_plane", "hit_tri",
                  "emissive", "matte", "specular", "glossy", "glossy_flip", "depth_cut", "miss",
                  "nodes_visited"]


class flux_counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in COUNTER_FIELDS]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n in COUNTER_FIELDS}


# Every symbol include/fluxb200.h declares: name -> (restype, argtypes)
_ctx = C.c_void_p
PROTOTYPES = {
    "flux_ctx_create": (C.c_int, [C.c_int, C.POINTER(_ctx)]),
    "flux_ctx_destroy": (C.c_int, [_ctx]),
    "flux_last_error": (C.c_char_p, [_ctx]),
    "flux_version": (C.c_char_p, []),
    "flux_set_scene": (C.c_int, [_ctx, C.POINTER(flux_scene_flat), C.POINTER(flux_job_config)]),
    "flux_set_samples": (C.c_int, [_ctx, C.c_uint32, C.c_uint32, C.c_uint32, _dp, _dp, _dp]),
    "flux_generate_samples": (C.c_int, [_ctx, C.c_uint64, C.c_uint32]),
    "flux_get_samples": (C.c_int, [_ctx, _dp, _dp, _dp]),
    "flux_get_set_index": (C.c_int, [_ctx, _u32p]),
    "flux_set_set_index": (C.c_int, [_ctx, _u32p]),
    "flux_render_rows": (C.c_int, [_ctx, C.c_uint32, C.c_uint32, _dp]),
    "flux_render_rows_device": (C.c_int, [_ctx, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]),
    "flux_shard_rows": (C.c_int, [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, _u32p, _u32p]),
    "flux_render_row_list": (C.c_int, [_ctx, _u32p, C.c_uint32, _dp]),
    "flux_render_row_list_device": (C.c_int, [_ctx, _u32p, C.c_uint32, C.c_void_p, C.c_void_p]),
    "flux_trace_rays": (C.c_int, [_ctx, C.c_uint64, _dp, _dp, _i32p, _dp]),
    "flux_trace_rays_device": (C.c_int, [_ctx, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p]),
    "flux_enable_counters": (C.c_int, [_ctx, C.c_int]),
    "flux_get_counters": (C.c_int, [_ctx, C.POINTER(flux_counters)]),
    "flux_reset_counters": (C.c_int, [_ctx]),
    "flux_last_kernel_ms": (C.c_int, [_ctx, C.POINTER(C.c_float)]),
    "flux_launch_count": (C.c_int, [_ctx, C.POINTER(C.c_uint64)]),
    "flux_set_accel_mode": (C.c_int, [_ctx, C.c_int]),
    "flux_bvh_describe": (C.c_int, [C.POINTER(flux_scene_flat), C.POINTER(C.c_uint64)]),
    "flux_bvh_hash": (C.c_int, [C.POINTER(flux_scene_flat), C.POINTER(C.c_uint64)]),
    "flux_set_kernel_mode": (C.c_int, [_ctx, C.c_int]),
    "flux_set_glossy_table": (C.c_int, [_ctx, C.c_int]),
    "flux_measure_fp64_peak": (C.c_int, [_ctx, _dp]),
    "flux_write_ppm": (C.c_int, [C.c_char_p, C.c_uint32, C.c_uint32, _dp]),
    "flux_progressive_begin": (C.c_int, [_ctx, C.POINTER(C.c_uint32), C.c_uint32]),
    "flux_progressive_pass": (C.c_int, [_ctx, C.c_uint32, C.c_uint32, _dp]),
    # multi-GPU frame assembly over peer memory
    "flux_frame_create": (C.c_int, [_ctx, C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p)]),
    "flux_frame_export": (C.c_int, [C.c_void_p, C.c_char_p]),
    "flux_frame_open_ipc": (C.c_int, [_ctx, C.c_char_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p)]),
    "flux_frame_open_peer": (C.c_int, [_ctx, C.c_void_p, C.POINTER(C.c_void_p)]),
    "flux_render_row_list_into_frame": (C.c_int, [_ctx, _u32p, C.c_uint32, C.c_void_p, C.c_void_p]),
    "flux_ctx_sync": (C.c_int, [_ctx]),
    "flux_frame_read": (C.c_int, [C.c_void_p, _dp]),
    "flux_frame_device_ptr": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "flux_frame_close": (C.c_int, [C.c_void_p]),
}
FLUX_FRAME_HANDLE_BYTES = 64

# FLUXB200_LIB selects an alternative build of the same library (A/B of compile-time kernel variants)
LIB_PATH = os.environ.get("FLUXB200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libfluxb200.so")
_lib = None


class FluxLibraryMissing(RuntimeError):
    pass


def lib():
    """Load libfluxb200.so (once) and set prototypes.  Raises loudly if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FluxLibraryMissing(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(flux_b200 has no CPU fallback for the render path)")
    l = C.CDLL(LIB_PATH, mode=C.RTLD_LOCAL)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(l, name)  # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = l
    return l


def as_dp(a):
    return a.ctypes.data_as(_dp)


def as_u32p(a):
    return a.ctypes.data_as(_u32p)


def as_i32p(a):
    return a.ctypes.data_as(_i32p)
