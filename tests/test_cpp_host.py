"""The C++ host (host/fluxhost.cpp: YAML loader, SceneData model, flattening) against the Python mirror:
both must hand the C-ABI the very same flux_scene_flat for the same scene file.  No GPU needed (--dump-flat)."""
import os
import subprocess

import numpy as np
import pytest

from flux_b200 import SceneData

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "host", "fluxb200")


@pytest.fixture(scope="module")
def fluxb200():
    subprocess.run(["make", "-C", os.path.join(ROOT, "host"), "-s"], check=True)
    assert os.path.exists(BIN)
    return BIN


def _hex(a):
    return [float(x).hex() for x in np.asarray(a, np.float64).ravel()]


def _parse_dump(path):
    out, mats = {}, []
    for line in open(path):
        f = line.split()
        if f[0] == "material":
            mats.append((int(f[1]),) + tuple(float.fromhex(x) for x in f[2:]))
        elif f[0] in ("image", "materials"):
            out[f[0]] = [int(x) for x in f[1:]]
        else:
            n = int(f[1])
            assert len(f) == n + 2, f[0]
            out[f[0]] = f[2:]
    out["material_list"] = mats
    return out


def _check_same(dump, sd):
    flat = sd.flatten()
    s = flat.struct
    assert dump["image"] == [s.image_width, s.image_height]
    as_hex = lambda xs: [float.fromhex(x).hex() for x in xs]
    assert as_hex(dump["scalars"]) == _hex([s.pixel_size, s.zoom_factor, s.view_plane_distance, s.focal_distance, s.lens_radius])
    for k in ("background", "eye", "look_at", "up"):
        assert as_hex(dump[k]) == _hex(list(getattr(s, k))), k
    assert dump["materials"] == [s.n_materials]
    for i, m in enumerate(dump["material_list"]):
        pm = s.materials[i]
        assert m[0] == pm.kind
        assert [x.hex() for x in m[1:]] == _hex([pm.color[0], pm.color[1], pm.color[2], pm.k, pm.exp])
    def arr(ptr, n, dtype=np.float64):
        return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype) if n else np.zeros(0, dtype)
    ns, npl, nt = s.n_spheres, s.n_planes, s.n_triangles
    assert as_hex(dump["sphere_center"]) == _hex(arr(s.sphere_center, 3 * ns))
    assert as_hex(dump["sphere_radius"]) == _hex(arr(s.sphere_radius, ns))
    assert [int(x) for x in dump["sphere_invert"]] == list(arr(s.sphere_invert, ns, np.int64))
    assert [int(x) for x in dump["sphere_shape_id"]] == list(arr(s.sphere_shape_id, ns, np.int64))
    assert [int(x) for x in dump["sphere_material"]] == list(arr(s.sphere_material, ns, np.int64))
    assert as_hex(dump["plane_point"]) == _hex(arr(s.plane_point, 3 * npl))
    assert as_hex(dump["plane_normal"]) == _hex(arr(s.plane_normal, 3 * npl))
    assert [int(x) for x in dump["plane_shape_id"]] == list(arr(s.plane_shape_id, npl, np.int64))
    assert [int(x) for x in dump["plane_material"]] == list(arr(s.plane_material, npl, np.int64))
    for k, p in (("tri_v0", s.tri_v0), ("tri_v1", s.tri_v1), ("tri_v2", s.tri_v2)):
        assert as_hex(dump[k]) == _hex(arr(p, 3 * nt)), k
    assert [int(x) for x in dump["tri_shape_id"]] == list(arr(s.tri_shape_id, nt, np.int64))
    assert [int(x) for x in dump["tri_material"]] == list(arr(s.tri_material, nt, np.int64))


@pytest.mark.parametrize("name", ["demo1", "demo2"])
def test_cpp_loader_flattens_shipped_scenes_like_python(fluxb200, tmp_path, name):
    path = os.path.join(ROOT, "scenes", f"{name}.yml")
    out = tmp_path / "flat.txt"
    subprocess.run([fluxb200, path, "--dump-flat", str(out)], check=True)
    _check_same(_parse_dump(out), SceneData.from_yaml(path))


EXT_YAML = """
# anchors, aliases, flow maps, a sequence at the key's own column, every extension shape
m: &grey
  Matte: {diffuse_color: [0.5, 0.5, 0.5], ambient_color: {r: 1, g: 1, b: 1}, diffuse_coefficient: 0.75}
scene_name: "ext scene"   # quoted
unknown_key: [1, 2, [3, 4]]
camera_settings: {eye: [0, 3, -9.5], look_at: [0, 1, 0], up: [0, 1, 0]}
camera_data:
  zoom_factor: 2
  view_plane_distance: 500.0
  focal_distance: 1.0e+1
  lens_radius: 0.05
output_settings:
  image_width: 96
  image_height: 64
  pixel_size: 0.5
background: [0.1, 0.2, 0.3]
shapes:
- Sphere:
    center: [0, 0, 0]
    radius: 100.0
    material:
      Emissive:
        color: [1, 0.9686, 0.8588]
        power: 0.5
    invert: true
- Plane: {point: [0, 0, 0], normal: [0, 1, 0], material: *grey}
- Triangle:
    v0: [-1, 0.5, 0]
    v1: [1, 0.5, 0]
    v2: [0, 2.5, 0.25]
    material:
      Reflective: {reflect_amount: 0.9, reflect_color: [0.9, 0.9, 1.0]}
- Rectangle:
    corner: [-3.1, 4.0, -1.0]
    edge_a: [0.7, 0, 0]
    edge_b: [0, 0, 0.9]
    material:
      Emissive: {color: [1, 1, 1], power: 12.5}
- Box:
    min: [1.5, 0.0, -0.5]
    max: [2.7, 1.3, 0.6]
    material: *grey
- Mesh:
    vertices: [[0, 0, 2], [1, 0, 2], [1, 1, 2.5],
               [0, 1, 2]]
    faces: [[0, 1, 2], [0, 2, 3]]
    material:
      GlossyReflective: {reflect_amount: 0.5, reflect_color: [1, 0.9, 0.9], reflect_exponent: 100.0}
"""


def test_cpp_loader_extension_shapes_and_yaml_features(fluxb200, tmp_path):
    src = tmp_path / "ext.yml"
    src.write_text(EXT_YAML)
    out = tmp_path / "flat.txt"
    subprocess.run([fluxb200, str(src), "--dump-flat", str(out)], check=True)
    sd = SceneData.from_yaml(str(src))
    assert sd.scene_name == "ext scene"
    d = _parse_dump(out)
    _check_same(d, sd)
    assert len(d["tri_shape_id"]) == 1 + 2 + 12 + 2
    assert [int(x) for x in d["tri_shape_id"]] == list(range(2, 19))   # one shape-id space across kinds


@pytest.mark.parametrize("text,msg", [
    ("scene_name: x\n", "missing field `output_settings`"),
    (EXT_YAML.replace("    invert: true\n", ""), "missing field `invert`"),
    (EXT_YAML.replace("Emissive:\n        color", "Glowing:\n        color"), "unknown variant `Glowing`"),
    (EXT_YAML.replace("image_width: 96", "image_width: -96"), "image_width"),
    (EXT_YAML.replace("material: *grey", "material: *nope", 1), "unknown alias"),
])
def test_cpp_loader_errors_mirror_serde(fluxb200, tmp_path, text, msg):
    """serde_yaml errors make the reference panic (flux/src/main.rs:28-29 unwrap); the driver exits 101 with the text."""
    src = tmp_path / "bad.yml"
    src.write_text(text)
    r = subprocess.run([fluxb200, str(src), "--dump-flat", str(tmp_path / "o.txt")], capture_output=True, text=True)
    assert r.returncode == 101
    assert msg in r.stderr
    with pytest.raises(Exception):
        SceneData.from_yaml(str(src))


def test_cli_flags_mirror_the_reference(fluxb200):
    r = subprocess.run([fluxb200], capture_output=True, text=True)
    assert r.returncode == 2 and "Scene filename is required" in r.stderr
    r = subprocess.run([fluxb200, "x.yml", "-r", "abc"], capture_output=True, text=True)
    assert r.returncode == 2
    r = subprocess.run([fluxb200, "--help"], capture_output=True, text=True)
    assert r.returncode == 0 and "--root" in r.stderr and "--depth" in r.stderr and "--rows" in r.stderr


@pytest.mark.gpu
def test_cli_renders_the_same_ppm_as_the_python_mirror(fluxb200, tmp_path):
    """fluxb200 scenes/demo2.yml -r 3 (64x48 override) writes byte-for-byte the PPM the Python GpuWorker path writes:
    same C-ABI, same device-generated sample sets (seed 1), Image::write quantisation (image.rs:42-60)."""
    import ctypes
    from flux_b200 import JobConfiguration, _capi
    from flux_b200.worker import GpuWorker
    scene = os.path.join(ROOT, "scenes", "demo2.yml")
    out = tmp_path / "cli.ppm"
    r = subprocess.run([fluxb200, scene, "-r", "3", "-d", "4", "--width", "64", "--height", "48", "-o", str(out)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "flux render (demo2, 9 samples per pixel, max depth 4)" in r.stdout
    assert "rendering finished, total time" in r.stdout
    sd = SceneData.from_yaml(scene).with_size(64, 48)
    w = GpuWorker(0, seed=1)
    img = w.render_image(sd, JobConfiguration(3, 4, 50))
    w.stop()
    ref = tmp_path / "py.ppm"
    assert _capi.lib().flux_write_ppm(str(ref).encode(), 64, 48, img.ctypes.data_as(ctypes.POINTER(ctypes.c_double))) == 0
    assert out.read_bytes() == ref.read_bytes()
    assert out.read_text().startswith("P3\n64 48\n65535\n")


@pytest.mark.gpu
def test_cli_on_two_gpus_writes_the_one_gpu_frame(fluxb200, tmp_path):
    """fluxb200 -G 2: one host thread per GPU, both kernels storing their rows into one frame on GPU 0 through
    cudaDeviceEnablePeerAccess (flux_frame_open_peer) — the PPM must be the 1-GPU PPM byte for byte.  Needs two GPUs."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    scene = os.path.join(ROOT, "scenes", "demo2.yml")
    outs = []
    for g in ("1", "2"):
        out = tmp_path / f"g{g}.ppm"
        r = subprocess.run([fluxb200, scene, "-r", "16", "--width", "96", "--height", "70", "-G", g, "-o", str(out)],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        outs.append(out.read_bytes())
    assert outs[0] == outs[1]


def _scenes_to_write():
    from flux_b200 import (CameraData, CameraSettings, Emissive, Matte, OutputSettings, PlaneData, SphereData, synth)
    from tests import helpers as Hp
    from tests.test_extension_shapes import room_scene
    yield "deterministic", Hp.deterministic_scene()
    yield "room", room_scene()                                            # rectangles, boxes
    yield "glossy", synth.glossy_scene()                                   # config 4: all material kinds
    yield "mesh", synth.mesh_scene(200, 100, seed=3, width=40, height=30)  # config 3 at 1/25 size: 40 000 triangles
    odd = [SphereData((5e-324, -0.0, 1e300), 2.2250738585072014e-308, Emissive((float("inf"), 0.1, 1 / 3), 1e-5), True),
           PlaneData((1e21, -1e-7, 123456789.125), (0.0, 1.0, float("nan")), Matte((0.1, 0.2, 0.3), (1, 1, 1), 1.0))]
    yield "odd floats", SceneData('name with "quotes" and: colons', OutputSettings(3, 2, 0.5), (0.0, 0.0, 0.0), odd,
                                  CameraSettings((0, 0, -5), (0, 0, 0), (0, 1, 0)), CameraData(1.0, 500.0, 10.0, 0.0))


@pytest.mark.parametrize("name,sd", list(_scenes_to_write()), ids=[n for n, _ in _scenes_to_write()])
def test_written_scene_files_load_back_exactly_in_both_loaders(fluxb200, tmp_path, name, sd):
    """SceneData.to_yaml writes what serde_yaml reads; the Python loader and the C++ loader get every double back
    bit for bit (NaN, infinities, -0.0, denormals and exponents without a fraction included)."""
    path = tmp_path / "scene.yml"
    sd.to_yaml(str(path))
    back = SceneData.from_yaml(str(path))
    assert back.scene_name == sd.scene_name and len(back.shapes) == len(sd.shapes)
    out = tmp_path / "flat.txt"
    subprocess.run([fluxb200, str(path), "--dump-flat", str(out)], check=True)
    d = _parse_dump(out)
    _check_same(d, sd)        # C++ loader of the written file == the original scene, flattened
    _check_same(d, back)      # ... == the Python loader of the written file
