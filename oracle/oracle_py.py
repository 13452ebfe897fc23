"""ctypes binding of oracle/liboracle.so — TEST INFRASTRUCTURE ONLY.

Importable only from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  flux_b200/ never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from flux_b200 import _capi
from flux_b200.scene import FlatScene, JobConfiguration

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
_dp, _u32p, _i32p = C.POINTER(C.c_double), C.POINTER(C.c_uint32), C.POINTER(C.c_int32)
_lib = None


def build():
    subprocess.run(["make", "-C", _HERE, "-s"], check=True)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    src = os.path.join(_HERE, "flux_oracle.cpp")
    if not os.path.exists(LIB_PATH) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(LIB_PATH)):
        build()
    l = C.CDLL(LIB_PATH)
    S, J, K = C.POINTER(_capi.flux_scene_flat), C.POINTER(_capi.flux_job_config), C.POINTER(_capi.flux_counters)
    l.oracle_version.restype = C.c_char_p
    l.oracle_render_rows.argtypes = [S, J, C.c_uint32, _dp, _dp, _dp, _u32p, C.c_uint32, C.c_uint32, _dp, K, C.c_int]
    l.oracle_render_row_list.argtypes = [S, J, C.c_uint32, _dp, _dp, _dp, _u32p, _u32p, C.c_uint32, _dp, K, C.c_int]
    l.oracle_trace_rays.argtypes = [S, C.c_uint64, _dp, _dp, _i32p, _dp, C.c_int]
    l.oracle_hit_record.argtypes = [S, _dp, _dp, _dp, _dp, _dp, _u32p]
    l.oracle_generate_samples.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, _dp, _dp, _dp]
    l.oracle_generate_samples.restype = None
    l.oracle_mj_grid.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, _dp]
    l.oracle_mj_grid.restype = None
    l.oracle_generate_set_index.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, _u32p]
    l.oracle_generate_set_index.restype = None
    l.oracle_camera_basis.argtypes = [_dp] * 6
    l.oracle_camera_basis.restype = None
    l.oracle_bbox_hit.argtypes = [_dp] * 4
    l.oracle_sphere_hit.argtypes = [_dp, C.c_double, C.c_int, _dp, _dp, _dp, _dp, _dp]
    l.oracle_plane_hit.argtypes = [_dp] * 7
    l.oracle_triangle_hit.argtypes = [_dp] * 8
    l.oracle_lambertian_sample_f.argtypes = [_dp, _dp, _dp, C.c_double, _dp, _dp, _dp]
    l.oracle_lambertian_sample_f.restype = None
    l.oracle_specular_sample_f.argtypes = [_dp, _dp, _dp, C.c_double, _dp, _dp, _dp]
    l.oracle_specular_sample_f.restype = None
    l.oracle_glossy_sample_f.argtypes = [_dp, _dp, _dp, _dp, C.c_double, C.c_double, _dp, _dp, _dp]
    l.oracle_to_unit_hemi.argtypes = [C.c_double, C.c_double, C.c_double, _dp]
    l.oracle_to_unit_hemi.restype = None
    l.oracle_to_poisson_disc.argtypes = [C.c_double, C.c_double, _dp]
    l.oracle_to_poisson_disc.restype = None
    l.oracle_max_to_one.argtypes = [_dp]
    l.oracle_max_to_one.restype = None
    l.oracle_primary_ray.argtypes = [S, C.c_uint32, C.c_uint32, C.c_double, C.c_double, C.c_double, C.c_double, _dp, _dp]
    l.oracle_ppm_quantize.argtypes = [_dp, C.c_uint64, C.POINTER(C.c_uint16)]
    l.oracle_ppm_quantize.restype = None
    _lib = l
    return l


def _v(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64))


def _p(a):
    return a.ctypes.data_as(_dp)


class SampleSets:
    """MasterSampleSets (sampling.rs:5-10) in reference layout + the set-index map."""

    def __init__(self, root, max_depth, num_sets, pixel, disc, hemi, set_index=None):
        self.root, self.max_depth, self.num_sets = root, max_depth, num_sets
        self.pixel, self.disc, self.hemi, self.set_index = pixel, disc, hemi, set_index


def generate_samples(seed, root, max_depth, num_sets) -> SampleSets:
    n = root * root
    pixel = np.empty((num_sets, n, 2), np.float64)
    disc = np.empty((num_sets, n, 2), np.float64)
    hemi = np.empty((num_sets, max_depth, n, 3), np.float64)
    lib().oracle_generate_samples(seed, root, max_depth, num_sets, _p(pixel), _p(disc), _p(hemi))
    return SampleSets(root, max_depth, num_sets, pixel, disc, hemi)


def generate_set_index(seed, height, width, num_sets) -> np.ndarray:
    idx = np.empty((height, width), np.uint32)
    lib().oracle_generate_set_index(seed, height, width, num_sets, idx.ctypes.data_as(_u32p))
    return idx


def mj_grid(seed, root, set_, grid, correlated) -> np.ndarray:
    out = np.empty((root * root, 2), np.float64)
    lib().oracle_mj_grid(seed, root, set_, grid, 1 if correlated else 0, _p(out))
    return out


def render_row_list(flat: FlatScene, cfg: JobConfiguration, ss: SampleSets, rows, counters=False, threads=0):
    rows = np.ascontiguousarray(np.asarray(rows, dtype=np.uint32))
    W = flat.struct.image_width
    out = np.empty((rows.shape[0], W, 3), np.float64)
    jc = _capi.flux_job_config(cfg.sample_root, cfg.max_trace_depth, cfg.rows_per_work_unit)
    cn = _capi.flux_counters() if counters else None
    idx = np.ascontiguousarray(ss.set_index, dtype=np.uint32)
    rc = lib().oracle_render_row_list(flat.ptr(), C.byref(jc), ss.num_sets, _p(ss.pixel), _p(ss.disc), _p(ss.hemi),
                                      idx.ctypes.data_as(_u32p), rows.ctypes.data_as(_u32p), rows.shape[0],
                                      _p(out), C.byref(cn) if counters else None, threads)
    if rc != 0:
        raise RuntimeError(f"oracle_render_row_list failed: {rc}")
    return (out, cn.as_dict()) if counters else out


def render_rows(flat, cfg, ss, row_start, row_end_inclusive, counters=False, threads=0):
    return render_row_list(flat, cfg, ss, np.arange(row_start, row_end_inclusive + 1), counters, threads)


def trace_rays(flat: FlatScene, origins, dirs, threads=0):
    o, d = _v(origins), _v(dirs)
    n = o.shape[0]
    hit = np.empty(n, np.int32)
    t = np.empty(n, np.float64)
    rc = lib().oracle_trace_rays(flat.ptr(), n, _p(o), _p(d), hit.ctypes.data_as(_i32p), _p(t), threads)
    if rc != 0:
        raise RuntimeError(f"oracle_trace_rays failed: {rc}")
    return hit, t


def hit_record(flat: FlatScene, o, d):
    o, d = _v(o), _v(d)
    t = C.c_double()
    normal, point = np.empty(3), np.empty(3)
    mat = C.c_uint32()
    sid = lib().oracle_hit_record(flat.ptr(), _p(o), _p(d), C.byref(t), _p(normal), _p(point), C.byref(mat))
    if sid < 0:
        return None
    return dict(shape_id=sid, t=t.value, normal=normal, point=point, material=mat.value)


def camera_basis(eye, look_at, up):
    u, v, w = np.empty(3), np.empty(3), np.empty(3)
    lib().oracle_camera_basis(_p(_v(eye)), _p(_v(look_at)), _p(_v(up)), _p(u), _p(v), _p(w))
    return u, v, w


def bbox_hit(c0, c1, o, d) -> bool:
    return bool(lib().oracle_bbox_hit(_p(_v(c0)), _p(_v(c1)), _p(_v(o)), _p(_v(d))))


def _hit3(fn, *args):
    t = C.c_double()
    normal, point = np.empty(3), np.empty(3)
    ok = fn(*args, C.byref(t), _p(normal), _p(point))
    return dict(t=t.value, normal=normal, point=point) if ok else None


def sphere_hit(center, radius, invert, o, d):
    return _hit3(lib().oracle_sphere_hit, _p(_v(center)), float(radius), 1 if invert else 0, _p(_v(o)), _p(_v(d)))


def plane_hit(point, normal, o, d):
    return _hit3(lib().oracle_plane_hit, _p(_v(point)), _p(_v(normal)), _p(_v(o)), _p(_v(d)))


def triangle_hit(v0, v1, v2, o, d):
    return _hit3(lib().oracle_triangle_hit, _p(_v(v0)), _p(_v(v1)), _p(_v(v2)), _p(_v(o)), _p(_v(d)))


def lambertian_sample_f(normal, hemi, cd, kd):
    wi, f, pdf = np.empty(3), np.empty(3), C.c_double()
    lib().oracle_lambertian_sample_f(_p(_v(normal)), _p(_v(hemi)), _p(_v(cd)), kd, _p(wi), C.byref(pdf), _p(f))
    return wi, pdf.value, f


def specular_sample_f(normal, wo, cr, kr):
    wi, f, pdf = np.empty(3), np.empty(3), C.c_double()
    lib().oracle_specular_sample_f(_p(_v(normal)), _p(_v(wo)), _p(_v(cr)), kr, _p(wi), C.byref(pdf), _p(f))
    return wi, pdf.value, f


def glossy_sample_f(normal, wo, sq, cs, ks, exp):
    wi, f, pdf = np.empty(3), np.empty(3), C.c_double()
    flipped = lib().oracle_glossy_sample_f(_p(_v(normal)), _p(_v(wo)), _p(_v(sq)), _p(_v(cs)), ks, exp,
                                           _p(wi), C.byref(pdf), _p(f))
    return wi, pdf.value, f, bool(flipped)


def to_unit_hemi(px, py, e):
    out = np.empty(3)
    lib().oracle_to_unit_hemi(px, py, e, _p(out))
    return out


def to_poisson_disc(px, py):
    out = np.empty(2)
    lib().oracle_to_poisson_disc(px, py, _p(out))
    return out


def max_to_one(rgb):
    a = _v(rgb).copy()
    lib().oracle_max_to_one(_p(a))
    return a


def primary_ray(flat: FlatScene, row, col, spx, spy, ldx, ldy):
    o, d = np.empty(3), np.empty(3)
    rc = lib().oracle_primary_ray(flat.ptr(), row, col, spx, spy, ldx, ldy, _p(o), _p(d))
    if rc != 0:
        raise RuntimeError("oracle_primary_ray failed")
    return o, d


def ppm_quantize(rgb):
    a = _v(rgb).ravel()
    out = np.empty(a.shape[0], np.uint16)
    lib().oracle_ppm_quantize(_p(a), a.shape[0], out.ctypes.data_as(C.POINTER(C.c_uint16)))
    return out


def num_threads() -> int:
    return int(lib().oracle_num_threads())
