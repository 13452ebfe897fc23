"""The reference's network-rendering protocol, Python side.

Reference: ``NetworkWorkerRequest {SetJob(Box<Job>), WorkUnit(WorkUnit), Done}`` (fluxcore/src/workers.rs:106-110),
the manager's end ``NetworkWorker`` (workers.rs:118-245), the node's end (flux-node/src/main.rs:21-94),
``RenderEvent::RowsReady(WorkUnitResult)`` (manager.rs:16-28), ``WorkerInfo`` (manager.rs:221-224).  Messages are
CBOR items written back to back on one TCP connection by serde_cbor 0.9 (``to_writer`` / ``StreamDeserializer``).

This module is written independently of ``host/cbor.cpp`` / ``host/fluxnet.cpp`` (which serve the GPUs as
``fluxb200-node``): a small RFC 7049 codec, the serde tree of the reference's data types, and a ``NetworkWorker``
that drives any node speaking the protocol.  The tests pit the two implementations against each other and against
hand-derived byte vectors; the Rust reference itself cannot be built in this image, so serde_cbor's choices (structs
as text-keyed maps in declaration order, unit variants as text, f64 narrowed to f32 when exact and to f16 for NaN /
infinities) are taken from its documented behaviour.  Newtype enum variants have two forms: the 2-array
``[name, content]`` that serde_cbor < 0.10 writes (the reference pins 0.9.0) and the one-entry map
``{name: content}`` of 0.10 and later, which also read the older form.  ``form="array"`` is therefore the default for
everything sent; both are accepted on input.  Nothing here renders.
"""
from __future__ import annotations

import math
import socket
import struct
from typing import Any, List, Optional, Tuple

import numpy as np

from .scene import (BoxData, Emissive, GlossyReflective, JobConfiguration, Matte, MeshData, PlaneData, RectangleData,
                    Reflective, SceneData, SphereData, TriangleData, WorkUnit, WorkUnitResult, work_units)

DEFAULT_PORT = 2000  # constants.rs:6


class CborError(ValueError):
    pass


class Break:  # the 0xff stop code of an indefinite-length item
    pass


# ---------------------------------------------------------------------------------------------------------------
# encoder
# ---------------------------------------------------------------------------------------------------------------
def _head(major: int, v: int) -> bytes:
    m = major << 5
    if v < 24:
        return bytes([m | v])
    if v <= 0xFF:
        return bytes([m | 24, v])
    if v <= 0xFFFF:
        return bytes([m | 25]) + struct.pack(">H", v)
    if v <= 0xFFFFFFFF:
        return bytes([m | 26]) + struct.pack(">I", v)
    return bytes([m | 27]) + struct.pack(">Q", v)


def _float(v: float) -> bytes:
    """serde_cbor's serialize_f64: NaN / infinities as f16, f32 when the value survives the round trip, else f64."""
    if math.isnan(v):
        return b"\xf9\x7e\x00"
    if math.isinf(v):
        return b"\xf9\x7c\x00" if v > 0 else b"\xf9\xfc\x00"
    try:
        f32 = struct.pack(">f", v)
        if struct.unpack(">f", f32)[0] == v:
            return b"\xfa" + f32
    except OverflowError:
        pass
    return b"\xfb" + struct.pack(">d", v)


def dumps(x: Any) -> bytes:
    """bool / int >= 0 / float / str / list (array) / dict with str keys (map, insertion order)."""
    if isinstance(x, bool):
        return b"\xf5" if x else b"\xf4"
    if isinstance(x, (int, np.integer)):
        x = int(x)
        return _head(0, x) if x >= 0 else _head(1, -1 - x)
    if isinstance(x, (float, np.floating)):
        return _float(float(x))
    if isinstance(x, str):
        b = x.encode()
        return _head(3, len(b)) + b
    if isinstance(x, (list, tuple)):
        return _head(4, len(x)) + b"".join(dumps(e) for e in x)
    if isinstance(x, dict):
        return _head(5, len(x)) + b"".join(dumps(k) + dumps(v) for k, v in x.items())
    if x is None:
        return b"\xf6"
    raise TypeError(type(x))


# ---------------------------------------------------------------------------------------------------------------
# decoder (streaming: pulls bytes through `read(n)`, which returns exactly n bytes or raises EOFError)
# ---------------------------------------------------------------------------------------------------------------
def _half(h: int) -> float:
    return struct.unpack(">e", struct.pack(">H", h))[0]


def load(read, _depth: int = 0) -> Any:
    if _depth > 64:
        raise CborError("nesting too deep")
    first = read(1)[0]
    major, info = first >> 5, first & 31

    def arg():
        if info < 24:
            return info
        if info > 27:
            raise CborError("reserved additional information")
        return int.from_bytes(read(1 << (info - 24)), "big")

    if major == 0:
        return arg()
    if major == 1:
        return -1 - arg()
    if major in (2, 3):
        if info == 31:
            out = b""
            while True:
                c = read(1)[0]
                if c == 0xFF:
                    break
                if c >> 5 != major or c & 31 == 31:
                    raise CborError("bad chunk")
                n = c & 31
                if n >= 24:
                    n = int.from_bytes(read(1 << (n - 24)), "big")
                out += read(n)
        else:
            out = read(arg())
        return out.decode() if major == 3 else out
    if major == 4:
        if info == 31:
            items = []
            while True:
                e = load(read, _depth + 1)
                if e is Break:
                    return items
                items.append(e)
        return [load(read, _depth + 1) for _ in range(arg())]
    if major == 5:
        d = {}
        n = None if info == 31 else arg()
        while n is None or len(d) < n:
            k = load(read, _depth + 1)
            if k is Break:
                if n is None:
                    break
                raise CborError("break inside a definite-length map")
            d[k] = load(read, _depth + 1)
        return d
    if major == 6:
        arg()
        return load(read, _depth + 1)
    if info == 20:
        return False
    if info == 21:
        return True
    if info in (22, 23):
        return None
    if info == 25:
        return _half(arg())
    if info == 26:
        return struct.unpack(">f", read(4))[0]
    if info == 27:
        return struct.unpack(">d", read(8))[0]
    if info == 31:
        return Break
    raise CborError("unsupported simple value")


def loads(data: bytes, offset: int = 0) -> Tuple[Any, int]:
    """First item of `data` from `offset`; returns (value, offset past it)."""
    pos = [offset]

    def read(n):
        if pos[0] + n > len(data):
            raise EOFError("truncated CBOR item")
        b = data[pos[0]:pos[0] + n]
        pos[0] += n
        return b

    v = load(read)
    if v is Break:
        raise CborError("break outside an indefinite-length item")
    return v, pos[0]


# ---------------------------------------------------------------------------------------------------------------
# serde trees of the reference's types (field names and order as declared there)
# ---------------------------------------------------------------------------------------------------------------
def _v3(v) -> list:  # nalgebra Vector3 / Point3
    return [float(v[0]), float(v[1]), float(v[2])]


def _color(c) -> dict:  # color.rs:12-16
    return {"r": float(c[0]), "g": float(c[1]), "b": float(c[2])}


def variant(name: str, content, form: str = "array"):
    """A newtype enum variant in serde_cbor's array form (< 0.10) or map form (>= 0.10)."""
    if form == "array":
        return [name, content]
    if form == "map":
        return {name: content}
    raise ValueError("form must be 'array' or 'map'")


def variant_parts(x):
    """(name, content) of an enum value in either form; a unit variant has content None."""
    if isinstance(x, str):
        return x, None
    if isinstance(x, dict) and len(x) == 1:
        (k, v), = x.items()
        return k, v
    if isinstance(x, list) and 1 <= len(x) <= 2 and isinstance(x[0], str):
        return x[0], (x[1] if len(x) == 2 else None)
    raise CborError(f"not an enum value: {str(x)[:80]}")


def material_tree(m, form: str = "array"):  # shapes.rs:42-82
    V = lambda name, content: variant(name, content, form)
    if isinstance(m, Matte):
        return V("Matte", {"diffuse_color": _color(m.diffuse_color), "ambient_color": _color(m.ambient_color),
                          "diffuse_coefficient": float(m.diffuse_coefficient)})
    if isinstance(m, Emissive):
        return V("Emissive", {"color": _color(m.color), "power": float(m.power)})
    if isinstance(m, Reflective):
        return V("Reflective", {"reflect_amount": float(m.reflect_amount), "reflect_color": _color(m.reflect_color)})
    if isinstance(m, GlossyReflective):
        return V("GlossyReflective", {"reflect_amount": float(m.reflect_amount), "reflect_color": _color(m.reflect_color),
                                      "reflect_exponent": float(m.reflect_exponent)})
    raise TypeError(type(m))


def shape_tree(s, form: str = "array"):  # scene.rs:71-74, shapes.rs:18-37 (+ extensions)
    V = lambda name, content: variant(name, content, form)
    M = lambda m: material_tree(m, form)
    if isinstance(s, SphereData):
        return V("Sphere", {"center": _v3(s.center), "radius": float(s.radius), "material": M(s.material), "invert": bool(s.invert)})
    if isinstance(s, PlaneData):
        return V("Plane", {"point": _v3(s.point), "normal": _v3(s.normal), "material": M(s.material)})
    if isinstance(s, TriangleData):
        return V("Triangle", {"v0": _v3(s.v0), "v1": _v3(s.v1), "v2": _v3(s.v2), "material": M(s.material)})
    if isinstance(s, MeshData):
        return V("Mesh", {"vertices": [_v3(v) for v in np.asarray(s.vertices)],
                          "faces": [[int(i) for i in f] for f in np.asarray(s.faces)], "material": M(s.material)})
    if isinstance(s, RectangleData):
        return V("Rectangle", {"corner": _v3(s.corner), "edge_a": _v3(s.edge_a), "edge_b": _v3(s.edge_b), "material": M(s.material)})
    if isinstance(s, BoxData):
        return V("Box", {"min": _v3(s.min), "max": _v3(s.max), "material": M(s.material)})
    raise TypeError(type(s))


def scene_tree(sd: SceneData, form: str = "array") -> dict:  # scene.rs:42-49
    o, cs, cd = sd.output_settings, sd.camera_settings, sd.camera_data
    return {"scene_name": sd.scene_name,
            "output_settings": {"image_width": int(o.image_width), "image_height": int(o.image_height),
                                "pixel_size": float(o.pixel_size)},
            "background": _color(sd.background),
            "shapes": [shape_tree(s, form) for s in sd.shapes],
            "camera_settings": {"eye": _v3(cs.eye), "look_at": _v3(cs.look_at), "up": _v3(cs.up)},
            "camera_data": {"zoom_factor": float(cd.zoom_factor), "view_plane_distance": float(cd.view_plane_distance),
                            "focal_distance": float(cd.focal_distance), "lens_radius": float(cd.lens_radius)}}


def work_unit_tree(u: WorkUnit) -> dict:  # job.rs:40-44; JobID(usize, usize) is a 2-tuple
    return {"row_start": int(u.row_start), "row_end": int(u.row_end), "job_id": [int(u.job_id[0]), int(u.job_id[1])]}


def job_tree(job_id: Tuple[int, int], sd: SceneData, cfg: JobConfiguration, form: str = "array") -> dict:  # job.rs:58-62
    return {"id": [int(job_id[0]), int(job_id[1])], "scene_data": scene_tree(sd, form),
            "config": {"sample_root": int(cfg.sample_root), "max_trace_depth": int(cfg.max_trace_depth),
                       "rows_per_work_unit": int(cfg.rows_per_work_unit)}}


def set_job(job_id: Tuple[int, int], sd: SceneData, cfg: JobConfiguration, form: str = "array") -> bytes:
    return dumps(variant("SetJob", job_tree(job_id, sd, cfg, form), form))


def work_unit(u: WorkUnit, form: str = "array") -> bytes:
    return dumps(variant("WorkUnit", work_unit_tree(u), form))


def done() -> bytes:
    return dumps("Done")


def worker_info(num_threads: int) -> bytes:
    return dumps({"num_threads": int(num_threads)})


def rows_ready_from_tree(ev) -> WorkUnitResult:
    """RenderEvent::RowsReady({work_unit, rows: [[{r,g,b}]]}) -> WorkUnitResult with rows [n][W][3] f64."""
    name, r = variant_parts(ev)
    if name != "RowsReady" or r is None:
        raise CborError(f"expected RenderEvent::RowsReady, got {str(ev)[:80]}")
    wu = r["work_unit"]
    rows = np.array([[[c["r"], c["g"], c["b"]] for c in row] for row in r["rows"]], dtype=np.float64)
    return WorkUnitResult(WorkUnit(wu["row_start"], wu["row_end"], tuple(wu["job_id"])), rows)


# ---------------------------------------------------------------------------------------------------------------
# the manager's end
# ---------------------------------------------------------------------------------------------------------------
class NetworkWorker:
    """workers.rs:118-245: connect, read the node's WorkerInfo, then per job SetJob, two units in flight, one
    RenderEvent per unit, Done.  A node ends the connection after Done, so one NetworkWorker serves one job."""

    def __init__(self, endpoint: str, timeout: Optional[float] = 600.0, form: str = "array"):
        self.form = form
        host, _, port = endpoint.rpartition(":") if ":" in endpoint else (endpoint, "", "")
        self.sock = socket.create_connection((host, int(port) if port else DEFAULT_PORT), timeout=timeout)
        self.sock.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
        self._rf = self.sock.makefile("rb")
        info = load(self._read)
        self.num_threads = int(info["num_threads"])

    def _read(self, n: int) -> bytes:
        b = self._rf.read(n)
        if b is None or len(b) != n:
            raise EOFError("connection closed inside a CBOR item")
        return b

    def info(self) -> dict:
        return {"num_threads": self.num_threads}

    def render_job(self, sd: SceneData, cfg: JobConfiguration, job_id: Tuple[int, int] = (1, 0),
                   units: Optional[List[WorkUnit]] = None) -> np.ndarray:
        h, w = sd.output_settings.image_height, sd.output_settings.image_width
        if units is None:
            units = [WorkUnit(u.row_start, u.row_end, tuple(job_id)) for u in work_units(h, cfg.rows_per_work_unit)]
        img = np.zeros((h, w, 3), np.float64)
        self.sock.sendall(set_job(job_id, sd, cfg, self.form))

        def collect():
            res = rows_ready_from_tree(load(self._read))
            if tuple(res.work_unit.job_id) != tuple(job_id):
                raise CborError("rows of another job")
            img[res.work_unit.row_start:res.work_unit.row_end + 1] = res.rows

        sent = received = 0
        while sent < len(units) and sent < 2:
            self.sock.sendall(work_unit(units[sent], self.form)); sent += 1
        while sent < len(units):
            self.sock.sendall(work_unit(units[sent], self.form)); sent += 1
            collect(); received += 1
        while received < len(units):
            collect(); received += 1
        self.sock.sendall(done())
        self.close()
        return img

    def close(self):
        try:
            self._rf.close()
            self.sock.close()
        except OSError:
            pass
