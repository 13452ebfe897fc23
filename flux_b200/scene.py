"""Plain-data scene types of the reference and their flattening for the C-ABI.

Host-side mirror of the reference's data model for the render path (same names,
same field meaning, same YAML format):

* ``SceneData``/``OutputSettings``/``CameraSettings``/``CameraData`` —
  fluxcore/src/scene.rs:11-66
* ``SphereData``/``PlaneData`` and the ``MaterialData`` variants ``Matte``,
  ``Emissive``, ``Reflective``, ``GlossyReflective`` — fluxcore/src/shapes.rs:15-83
* ``JobConfiguration``/``WorkUnit`` — fluxcore/src/job.rs:40-53
* ``WorkUnitResult`` — fluxcore/src/manager.rs:25-28

``scenes/*.yml`` are serde-YAML of ``SceneData`` (flux/src/main.rs:28-29):
externally tagged enums (``- Sphere: {...}``, ``material: {Matte: {...}}``),
vectors and colours as 3-element sequences, YAML anchors/aliases, unknown keys
ignored, every field required (no ``#[serde(default)]`` anywhere).

``TriangleData``/``MeshData`` are EXTENSIONS (the reference has only Sphere and
Plane, scene.rs:71-74); see DESIGN.md.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Sequence, Union

import numpy as np
import yaml

from . import _capi

Vec3 = Sequence[float]


def _vec3(v, what: str) -> tuple:
    if isinstance(v, dict):  # serde also accepts a map for Color {r,g,b}
        try:
            v = [v["r"], v["g"], v["b"]]
        except KeyError as e:
            raise ValueError(f"{what}: missing field {e}") from None
    if not isinstance(v, (list, tuple)) or len(v) != 3:
        raise ValueError(f"{what}: expected a sequence of 3 numbers, got {v!r}")
    return (float(v[0]), float(v[1]), float(v[2]))


def _req(d: dict, key: str, what: str):
    if not isinstance(d, dict) or key not in d:
        raise ValueError(f"{what}: missing field `{key}`")
    return d[key]


# ---- MaterialData, shapes.rs:42-83 -------------------------------------------
@dataclass(frozen=True)
class Matte:
    diffuse_color: tuple
    ambient_color: tuple
    diffuse_coefficient: float


@dataclass(frozen=True)
class Emissive:
    color: tuple
    power: float


@dataclass(frozen=True)
class Reflective:
    reflect_amount: float
    reflect_color: tuple


@dataclass(frozen=True)
class GlossyReflective:
    reflect_amount: float
    reflect_color: tuple
    reflect_exponent: float


MaterialData = Union[Matte, Emissive, Reflective, GlossyReflective]


def material_from_yaml(d) -> MaterialData:
    if not isinstance(d, dict) or len(d) != 1:
        raise ValueError(f"material: expected a single-key map (externally tagged enum), got {d!r}")
    (tag, m), = d.items()
    if tag == "Matte":
        return Matte(_vec3(_req(m, "diffuse_color", tag), "diffuse_color"),
                     _vec3(_req(m, "ambient_color", tag), "ambient_color"),
                     float(_req(m, "diffuse_coefficient", tag)))
    if tag == "Emissive":
        return Emissive(_vec3(_req(m, "color", tag), "color"), float(_req(m, "power", tag)))
    if tag == "Reflective":
        return Reflective(float(_req(m, "reflect_amount", tag)),
                          _vec3(_req(m, "reflect_color", tag), "reflect_color"))
    if tag == "GlossyReflective":
        return GlossyReflective(float(_req(m, "reflect_amount", tag)),
                                _vec3(_req(m, "reflect_color", tag), "reflect_color"),
                                float(_req(m, "reflect_exponent", tag)))
    raise ValueError(f"unknown variant `{tag}`, expected one of "
                     "`Matte`, `Emissive`, `Reflective`, `GlossyReflective`")


def material_to_flat(m: MaterialData):
    """MaterialData -> (kind, color, k, exp); see include/fluxb200.h flux_material."""
    if isinstance(m, Matte):
        return (_capi.FLUX_MAT_MATTE, m.diffuse_color, m.diffuse_coefficient, 0.0)
    if isinstance(m, Emissive):
        return (_capi.FLUX_MAT_EMISSIVE, m.color, m.power, 0.0)
    if isinstance(m, Reflective):
        return (_capi.FLUX_MAT_REFLECTIVE, m.reflect_color, m.reflect_amount, 0.0)
    if isinstance(m, GlossyReflective):
        return (_capi.FLUX_MAT_GLOSSY, m.reflect_color, m.reflect_amount, m.reflect_exponent)
    raise TypeError(f"not a MaterialData: {m!r}")


# ---- ShapeData, scene.rs:71-74 + shapes.rs:15-37 ------------------------------
@dataclass(frozen=True)
class SphereData:
    center: tuple
    radius: float
    material: MaterialData
    invert: bool


@dataclass(frozen=True)
class PlaneData:
    point: tuple
    normal: tuple
    material: MaterialData


@dataclass(frozen=True)
class TriangleData:  # EXTENSION
    v0: tuple
    v1: tuple
    v2: tuple
    material: MaterialData


@dataclass
class MeshData:  # EXTENSION: many triangles sharing one material
    vertices: np.ndarray  # [nv][3] f64
    faces: np.ndarray     # [nf][3] u32/i64
    material: MaterialData


@dataclass(frozen=True)
class RectangleData:  # EXTENSION (TODO.md:2 "Quad (for area light)"): corner + two edge vectors, two triangles
    corner: tuple
    edge_a: tuple
    edge_b: tuple
    material: MaterialData

    def triangles(self):
        """(v0, v1, v2) triples; both wound so that the normal is edge_a x edge_b."""
        c = np.asarray(self.corner, np.float64)
        pa = c + np.asarray(self.edge_a, np.float64)
        pab = pa + np.asarray(self.edge_b, np.float64)
        pb = c + np.asarray(self.edge_b, np.float64)
        return [(c, pa, pab), (c, pab, pb)]


@dataclass(frozen=True)
class BoxData:  # EXTENSION: axis-aligned box [min, max], six outward-facing rectangles = twelve triangles
    min: tuple
    max: tuple
    material: MaterialData

    def rectangles(self):
        (x0, y0, z0), (x1, y1, z1) = self.min, self.max
        dx, dy, dz = x1 - x0, y1 - y0, z1 - z0
        m = self.material
        return [
            RectangleData((x0, y0, z0), (0.0, dy, 0.0), (dx, 0.0, 0.0), m),   # z = z0, normal -z
            RectangleData((x0, y0, z1), (dx, 0.0, 0.0), (0.0, dy, 0.0), m),   # z = z1, normal +z
            RectangleData((x0, y0, z0), (0.0, 0.0, dz), (0.0, dy, 0.0), m),   # x = x0, normal -x
            RectangleData((x1, y0, z0), (0.0, dy, 0.0), (0.0, 0.0, dz), m),   # x = x1, normal +x
            RectangleData((x0, y0, z0), (dx, 0.0, 0.0), (0.0, 0.0, dz), m),   # y = y0, normal -y
            RectangleData((x0, y1, z0), (0.0, 0.0, dz), (dx, 0.0, 0.0), m),   # y = y1, normal +y
        ]


ShapeData = Union[SphereData, PlaneData, TriangleData, MeshData, RectangleData, BoxData]


def shape_from_yaml(d) -> ShapeData:
    if not isinstance(d, dict) or len(d) != 1:
        raise ValueError(f"shape: expected a single-key map (externally tagged enum), got {d!r}")
    (tag, s), = d.items()
    if tag == "Sphere":
        inv = _req(s, "invert", tag)
        if not isinstance(inv, bool):
            raise ValueError(f"Sphere.invert: expected a boolean, got {inv!r}")
        return SphereData(_vec3(_req(s, "center", tag), "center"), float(_req(s, "radius", tag)),
                          material_from_yaml(_req(s, "material", tag)), inv)
    if tag == "Plane":
        return PlaneData(_vec3(_req(s, "point", tag), "point"), _vec3(_req(s, "normal", tag), "normal"),
                         material_from_yaml(_req(s, "material", tag)))
    if tag == "Triangle":  # EXTENSION
        return TriangleData(_vec3(_req(s, "v0", tag), "v0"), _vec3(_req(s, "v1", tag), "v1"),
                            _vec3(_req(s, "v2", tag), "v2"), material_from_yaml(_req(s, "material", tag)))
    if tag == "Mesh":  # EXTENSION
        v = np.asarray(_req(s, "vertices", tag), dtype=np.float64).reshape(-1, 3)
        f = np.asarray(_req(s, "faces", tag), dtype=np.int64).reshape(-1, 3)
        return MeshData(v, f, material_from_yaml(_req(s, "material", tag)))
    if tag == "Rectangle":  # EXTENSION
        return RectangleData(_vec3(_req(s, "corner", tag), "corner"), _vec3(_req(s, "edge_a", tag), "edge_a"),
                             _vec3(_req(s, "edge_b", tag), "edge_b"), material_from_yaml(_req(s, "material", tag)))
    if tag == "Box":  # EXTENSION
        return BoxData(_vec3(_req(s, "min", tag), "min"), _vec3(_req(s, "max", tag), "max"),
                       material_from_yaml(_req(s, "material", tag)))
    raise ValueError(f"unknown variant `{tag}`, expected `Sphere` or `Plane`")


@dataclass(frozen=True)
class OutputSettings:  # scene.rs:58-63
    image_width: int
    image_height: int
    pixel_size: float


@dataclass(frozen=True)
class CameraSettings:  # scene.rs:11-16
    eye: tuple
    look_at: tuple
    up: tuple


@dataclass(frozen=True)
class CameraData:  # scene.rs:50-56
    zoom_factor: float
    view_plane_distance: float
    focal_distance: float
    lens_radius: float


@dataclass(frozen=True)
class JobConfiguration:  # job.rs:49-53
    sample_root: int
    max_trace_depth: int = 5      # flux/src/main.rs:21
    rows_per_work_unit: int = 50  # flux/src/main.rs:172


@dataclass(frozen=True)
class WorkUnit:  # job.rs:40-44 (row_end inclusive)
    row_start: int
    row_end: int
    job_id: tuple = (0, 0)


@dataclass
class WorkUnitResult:  # manager.rs:25-28; rows: [n_rows][W][3] f64
    work_unit: WorkUnit
    rows: np.ndarray


def work_units(image_height: int, rows_per_work_unit: int) -> List[WorkUnit]:
    """Row bands covering the whole image.

    Follows Job::work_units (job.rs:66-88) except for its loop condition
    ``i < image_height - 1``, which silently drops a trailing single row (and
    yields nothing for a 1-row image); here every row is covered.
    """
    if rows_per_work_unit == 0:
        raise ValueError(f"Job row per work unit count invalid: {rows_per_work_unit}")
    us, i = [], 0
    while i < image_height:
        n = min(rows_per_work_unit, image_height - i)
        us.append(WorkUnit(i, i + n - 1))
        i += n
    return us


class FlatScene:
    """``flux_scene_flat`` plus the numpy arrays that back its pointers."""

    def __init__(self, struct, keep):
        self.struct = struct
        self._keep = keep

    def ptr(self):
        return C.byref(self.struct)


@dataclass
class SceneData:  # scene.rs:42-49
    scene_name: str
    output_settings: OutputSettings
    background: tuple
    shapes: List[ShapeData]
    camera_settings: CameraSettings
    camera_data: CameraData

    # -- serde_yaml::from_reader (flux/src/main.rs:28-29) --
    @staticmethod
    def from_dict(d: dict) -> "SceneData":
        what = "SceneData"
        os_ = _req(d, "output_settings", what)
        cs = _req(d, "camera_settings", what)
        cd = _req(d, "camera_data", what)
        shapes = _req(d, "shapes", what)
        if not isinstance(shapes, list):
            raise ValueError("shapes: expected a sequence")
        w, h = _req(os_, "image_width", "output_settings"), _req(os_, "image_height", "output_settings")
        for n, v in (("image_width", w), ("image_height", h)):
            if not isinstance(v, int) or isinstance(v, bool) or v < 0:
                raise ValueError(f"output_settings.{n}: expected an unsigned integer, got {v!r}")
        return SceneData(
            scene_name=str(_req(d, "scene_name", what)),
            output_settings=OutputSettings(w, h, float(_req(os_, "pixel_size", "output_settings"))),
            background=_vec3(_req(d, "background", what), "background"),
            shapes=[shape_from_yaml(s) for s in shapes],
            camera_settings=CameraSettings(_vec3(_req(cs, "eye", "camera_settings"), "eye"),
                                           _vec3(_req(cs, "look_at", "camera_settings"), "look_at"),
                                           _vec3(_req(cs, "up", "camera_settings"), "up")),
            camera_data=CameraData(float(_req(cd, "zoom_factor", "camera_data")),
                                   float(_req(cd, "view_plane_distance", "camera_data")),
                                   float(_req(cd, "focal_distance", "camera_data")),
                                   float(_req(cd, "lens_radius", "camera_data"))),
        )

    @staticmethod
    def from_yaml(path: str) -> "SceneData":
        with open(path, "r") as f:
            return SceneData.from_dict(yaml.safe_load(f))

    @staticmethod
    def from_yaml_string(text: str) -> "SceneData":
        return SceneData.from_dict(yaml.safe_load(text))

    # -- the way back: a scene file serde_yaml (and both loaders here) reads; doubles survive exactly --
    def to_yaml(self, path: str) -> None:
        """Writes the scene as YAML in the reference's format (externally tagged enums, 3-sequences for vectors and
        colours; the extension shapes under their own tags).  Floats are written with repr() and always carry a
        '.', so YAML 1.1 and 1.2 readers both take them as floats and get the same double back."""
        def f(x):
            x = float(x)
            if x != x:
                return ".nan"
            if x in (float("inf"), float("-inf")):
                return ".inf" if x > 0 else "-.inf"
            r = repr(x)
            if "e" in r and "." not in r.split("e")[0]:
                m, e = r.split("e")
                r = m + ".0e" + e
            return r

        def v(p):
            return "[" + ", ".join(f(c) for c in p) + "]"

        def mat(m):
            if isinstance(m, Matte):
                return (f"Matte: {{diffuse_color: {v(m.diffuse_color)}, ambient_color: {v(m.ambient_color)}, "
                        f"diffuse_coefficient: {f(m.diffuse_coefficient)}}}")
            if isinstance(m, Emissive):
                return f"Emissive: {{color: {v(m.color)}, power: {f(m.power)}}}"
            if isinstance(m, Reflective):
                return f"Reflective: {{reflect_amount: {f(m.reflect_amount)}, reflect_color: {v(m.reflect_color)}}}"
            return (f"GlossyReflective: {{reflect_amount: {f(m.reflect_amount)}, reflect_color: {v(m.reflect_color)}, "
                    f"reflect_exponent: {f(m.reflect_exponent)}}}")

        import json
        o, cs, cd = self.output_settings, self.camera_settings, self.camera_data
        with open(path, "w") as out:
            w = out.write
            w(f"scene_name: {json.dumps(self.scene_name)}\n")
            w(f"camera_settings:\n  eye: {v(cs.eye)}\n  look_at: {v(cs.look_at)}\n  up: {v(cs.up)}\n")
            w(f"camera_data:\n  zoom_factor: {f(cd.zoom_factor)}\n  view_plane_distance: {f(cd.view_plane_distance)}\n"
              f"  focal_distance: {f(cd.focal_distance)}\n  lens_radius: {f(cd.lens_radius)}\n")
            w(f"output_settings:\n  image_width: {int(o.image_width)}\n  image_height: {int(o.image_height)}\n"
              f"  pixel_size: {f(o.pixel_size)}\n")
            w(f"background: {v(self.background)}\nshapes:\n")
            for sh in self.shapes:
                if isinstance(sh, SphereData):
                    w(f"- Sphere:\n    center: {v(sh.center)}\n    radius: {f(sh.radius)}\n    material:\n      {mat(sh.material)}\n"
                      f"    invert: {'true' if sh.invert else 'false'}\n")
                elif isinstance(sh, PlaneData):
                    w(f"- Plane:\n    point: {v(sh.point)}\n    normal: {v(sh.normal)}\n    material:\n      {mat(sh.material)}\n")
                elif isinstance(sh, TriangleData):
                    w(f"- Triangle:\n    v0: {v(sh.v0)}\n    v1: {v(sh.v1)}\n    v2: {v(sh.v2)}\n    material:\n      {mat(sh.material)}\n")
                elif isinstance(sh, RectangleData):
                    w(f"- Rectangle:\n    corner: {v(sh.corner)}\n    edge_a: {v(sh.edge_a)}\n    edge_b: {v(sh.edge_b)}\n"
                      f"    material:\n      {mat(sh.material)}\n")
                elif isinstance(sh, BoxData):
                    w(f"- Box:\n    min: {v(sh.min)}\n    max: {v(sh.max)}\n    material:\n      {mat(sh.material)}\n")
                elif isinstance(sh, MeshData):
                    w("- Mesh:\n    vertices:\n")
                    for p in np.asarray(sh.vertices, np.float64):
                        w(f"      - {v(p)}\n")
                    w("    faces:\n")
                    for t in np.asarray(sh.faces, np.int64):
                        w(f"      - [{int(t[0])}, {int(t[1])}, {int(t[2])}]\n")
                    w(f"    material:\n      {mat(sh.material)}\n")
                else:
                    raise TypeError(f"not a ShapeData: {sh!r}")

    def with_size(self, width: int, height: int) -> "SceneData":
        """Same scene at another resolution (BASELINE config 1: demo1 at 512x512;
        the reference has no CLI override)."""
        return SceneData(self.scene_name, OutputSettings(width, height, self.output_settings.pixel_size),
                         self.background, self.shapes, self.camera_settings, self.camera_data)

    # -- flattening for the C-ABI (include/fluxb200.h flux_scene_flat) --
    def flatten(self) -> FlatScene:
        mats: List[tuple] = []
        mat_index = {}

        def mat_id(m):
            key = material_to_flat(m)
            if key not in mat_index:
                mat_index[key] = len(mats)
                mats.append(key)
            return mat_index[key]

        sc, sr, si, sid, sm = [], [], [], [], []
        pp, pn, pid, pm = [], [], [], []
        t0, t1, t2, tid, tm = [], [], [], [], []
        shape_id = 0
        for sh in self.shapes:
            if isinstance(sh, SphereData):
                sc.append(sh.center); sr.append(sh.radius); si.append(1 if sh.invert else 0)
                sid.append(shape_id); sm.append(mat_id(sh.material)); shape_id += 1
            elif isinstance(sh, PlaneData):
                pp.append(sh.point); pn.append(sh.normal); pid.append(shape_id)
                pm.append(mat_id(sh.material)); shape_id += 1
            elif isinstance(sh, TriangleData):
                t0.append(np.asarray([sh.v0], np.float64)); t1.append(np.asarray([sh.v1], np.float64))
                t2.append(np.asarray([sh.v2], np.float64))
                tid.append(np.asarray([shape_id], np.uint32)); tm.append(np.asarray([mat_id(sh.material)], np.uint32))
                shape_id += 1
            elif isinstance(sh, (RectangleData, BoxData)):
                rects = [sh] if isinstance(sh, RectangleData) else sh.rectangles()
                for rc in rects:
                    for (a, b, c) in rc.triangles():
                        t0.append(np.asarray([a], np.float64)); t1.append(np.asarray([b], np.float64))
                        t2.append(np.asarray([c], np.float64))
                        tid.append(np.asarray([shape_id], np.uint32)); tm.append(np.asarray([mat_id(rc.material)], np.uint32))
                        shape_id += 1
            elif isinstance(sh, MeshData):
                f = np.asarray(sh.faces, dtype=np.int64)
                v = np.asarray(sh.vertices, dtype=np.float64)
                nf = f.shape[0]
                # np.take: the same gather as v[f[:, k]] (bounds checked alike) at half the time on a million faces
                t0.append(np.take(v, f[:, 0], axis=0)); t1.append(np.take(v, f[:, 1], axis=0)); t2.append(np.take(v, f[:, 2], axis=0))
                tid.append(np.arange(shape_id, shape_id + nf, dtype=np.uint32))
                tm.append(np.full(nf, mat_id(sh.material), np.uint32))
                shape_id += nf
            else:
                raise TypeError(f"not a ShapeData: {sh!r}")

        def arr(x, dtype, shape):
            a = np.ascontiguousarray(np.asarray(x, dtype=dtype).reshape(shape))
            return a

        def cat(xs, dtype, shape):
            if not xs:
                return np.zeros((0,) + tuple(shape[1:]), dtype)
            if len(xs) == 1 and xs[0].dtype == dtype and xs[0].flags.c_contiguous:   # one mesh: no second copy of its 72 MB
                return xs[0].reshape(shape)
            return np.ascontiguousarray(np.concatenate(xs).astype(dtype, copy=False).reshape(shape))

        mat_arr = (_capi.flux_material * max(1, len(mats)))()
        for i, (kind, color, k, ex) in enumerate(mats):
            mat_arr[i].kind = kind
            mat_arr[i].color[:] = color
            mat_arr[i].k = k
            mat_arr[i].exp = ex

        a_sc = arr(sc, np.float64, (-1, 3)); a_sr = arr(sr, np.float64, (-1,))
        a_si = arr(si, np.uint8, (-1,)); a_sid = arr(sid, np.uint32, (-1,)); a_sm = arr(sm, np.uint32, (-1,))
        a_pp = arr(pp, np.float64, (-1, 3)); a_pn = arr(pn, np.float64, (-1, 3))
        a_pid = arr(pid, np.uint32, (-1,)); a_pm = arr(pm, np.uint32, (-1,))
        a_t0 = cat(t0, np.float64, (-1, 3)); a_t1 = cat(t1, np.float64, (-1, 3)); a_t2 = cat(t2, np.float64, (-1, 3))
        a_tid = cat(tid, np.uint32, (-1,)); a_tm = cat(tm, np.uint32, (-1,))

        s = _capi.flux_scene_flat()
        s.image_width = self.output_settings.image_width
        s.image_height = self.output_settings.image_height
        s.pixel_size = self.output_settings.pixel_size
        s.background[:] = self.background
        s.eye[:] = self.camera_settings.eye
        s.look_at[:] = self.camera_settings.look_at
        s.up[:] = self.camera_settings.up
        s.zoom_factor = self.camera_data.zoom_factor
        s.view_plane_distance = self.camera_data.view_plane_distance
        s.focal_distance = self.camera_data.focal_distance
        s.lens_radius = self.camera_data.lens_radius
        s.n_materials = len(mats)
        s.materials = C.cast(mat_arr, C.POINTER(_capi.flux_material))
        dp, u8p, u32p = C.POINTER(C.c_double), C.POINTER(C.c_uint8), C.POINTER(C.c_uint32)
        s.n_spheres = a_sr.shape[0]
        s.sphere_center = a_sc.ctypes.data_as(dp); s.sphere_radius = a_sr.ctypes.data_as(dp)
        s.sphere_invert = a_si.ctypes.data_as(u8p); s.sphere_shape_id = a_sid.ctypes.data_as(u32p)
        s.sphere_material = a_sm.ctypes.data_as(u32p)
        s.n_planes = a_pid.shape[0]
        s.plane_point = a_pp.ctypes.data_as(dp); s.plane_normal = a_pn.ctypes.data_as(dp)
        s.plane_shape_id = a_pid.ctypes.data_as(u32p); s.plane_material = a_pm.ctypes.data_as(u32p)
        s.n_triangles = a_tid.shape[0]
        s.tri_v0 = a_t0.ctypes.data_as(dp); s.tri_v1 = a_t1.ctypes.data_as(dp); s.tri_v2 = a_t2.ctypes.data_as(dp)
        s.tri_shape_id = a_tid.ctypes.data_as(u32p); s.tri_material = a_tm.ctypes.data_as(u32p)
        keep = [mat_arr, a_sc, a_sr, a_si, a_sid, a_sm, a_pp, a_pn, a_pid, a_pm, a_t0, a_t1, a_t2, a_tid, a_tm]
        fs = FlatScene(s, keep)
        fs.n_shapes = shape_id
        fs.materials = mats
        return fs
