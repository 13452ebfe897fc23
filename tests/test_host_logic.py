"""Host-side logic that needs no GPU: scene YAML loading (the reference's scenes load unchanged),
flattening, work units, sharding, the op model."""
import os

import numpy as np
import pytest

from flux_b200 import (Emissive, GlossyReflective, JobConfiguration, Matte, PlaneData, SceneData, SphereData,
                       work_units)
from flux_b200 import _capi
from flux_b200.opsmodel import algorithmic_ops, ops_per_sample
from flux_b200.worker import shard_rows


def test_demo1_loads_unchanged(demo1):
    assert demo1.scene_name == "demo1"
    os_ = demo1.output_settings
    assert (os_.image_width, os_.image_height, os_.pixel_size) == (800, 600, 0.5)
    assert len(demo1.shapes) == 6  # 5 spheres + 1 plane; the commented-out area light is absent
    assert isinstance(demo1.shapes[0], SphereData) and demo1.shapes[0].invert is True
    assert isinstance(demo1.shapes[0].material, Emissive) and demo1.shapes[0].material.power == 1.0
    assert isinstance(demo1.shapes[5], PlaneData) and demo1.shapes[5].normal == (0.0, 1.0, 0.0)
    assert demo1.shapes[3].material == GlossyReflective(0.9, (0.9, 1.0, 0.9), 100000.0)
    assert demo1.camera_settings.eye == (2.5, 1.5, -9.0) and demo1.camera_data.lens_radius == 0.0


def test_demo2_anchors_aliases_and_unknown_keys(demo2):
    # demo2.yml defines materials under ignored top-level keys mat1..3 and aliases them (demo2.yml:1-15,54)
    assert len(demo2.shapes) == 13
    mats = [s.material for s in demo2.shapes[2:12]]
    assert mats[0] == GlossyReflective(0.5, (0.8, 0.6, 1.0), 10000.0)
    assert mats[0] == mats[3] == mats[6] == mats[9]
    assert mats[1] == GlossyReflective(0.5, (0.9, 1.0, 0.7), 100.0) and mats[2].reflect_exponent == 10.0
    assert demo2.camera_data.lens_radius == 0.09
    flat = demo2.flatten()
    assert flat.struct.n_spheres == 12 and flat.struct.n_planes == 1 and flat.struct.n_materials == 6
    assert list(np.ctypeslib.as_array(flat.struct.sphere_shape_id, (12,))) == list(range(12))
    assert flat.struct.plane_shape_id[0] == 12


def test_yaml_errors_mirror_serde():
    base = open(os.path.join(os.path.dirname(__file__), "..", "scenes", "demo1.yml")).read()
    with pytest.raises(ValueError, match="missing field `invert`"):
        SceneData.from_yaml_string(base.replace("      invert: true\n", "", 1))
    with pytest.raises(ValueError, match="unknown variant `Cube`"):
        SceneData.from_yaml_string(base.replace("  - Plane:", "  - Cube:"))
    with pytest.raises(ValueError, match="missing field `background`"):
        SceneData.from_yaml_string(base.replace("background: [0, 0, 0]\n", ""))
    with pytest.raises(ValueError, match="unknown variant `Shiny`"):
        SceneData.from_yaml_string(base.replace("        Emissive:", "        Shiny:", 1))
    with pytest.raises(ValueError, match="sequence of 3"):
        SceneData.from_yaml_string(base.replace("eye: [2.5, 1.5, -9.0]", "eye: [2.5, 1.5]"))
    sd = SceneData.from_yaml_string(base + "\nsome_unknown_key: 42\n")  # unknown keys are ignored
    assert sd.scene_name == "demo1"


def test_flatten_shape_ids_share_one_index_space():
    m = Matte((1, 1, 1), (1, 1, 1), 1.0)
    sd = SceneData.from_yaml_string(open(os.path.join(os.path.dirname(__file__), "..", "scenes", "demo1.yml")).read())
    sd.shapes = [PlaneData((0, 0, 0), (0, 1, 0), m), SphereData((0, 1, 0), 1.0, m, False),
                 PlaneData((0, 5, 0), (0, -1, 0), m), SphereData((3, 1, 0), 1.0, m, True)]
    f = sd.flatten().struct
    assert [f.plane_shape_id[i] for i in range(2)] == [0, 2]
    assert [f.sphere_shape_id[i] for i in range(2)] == [1, 3]
    assert [f.sphere_invert[i] for i in range(2)] == [0, 1]
    assert f.n_materials == 1  # identical materials are shared


def test_with_size_only_changes_resolution(demo1):
    sd = demo1.with_size(512, 512)  # BASELINE config 1
    assert (sd.output_settings.image_width, sd.output_settings.image_height) == (512, 512)
    assert sd.shapes == demo1.shapes and sd.camera_data == demo1.camera_data


def test_work_units_cover_all_rows():
    us = work_units(600, 50)
    assert len(us) == 12 and us[0].row_start == 0 and us[-1].row_end == 599
    assert all(u.row_end - u.row_start == 49 for u in us)
    us = work_units(601, 50)  # the reference drops this trailing row (job.rs:76 `i < h - 1`); we do not
    assert us[-1].row_start == 600 and us[-1].row_end == 600
    assert [(u.row_start, u.row_end) for u in work_units(1, 50)] == [(0, 0)]
    with pytest.raises(ValueError):
        work_units(10, 0)  # job.rs:67-70 panics


@pytest.mark.parametrize("height,tile,world", [(600, 4, 1), (600, 4, 8), (601, 4, 8), (7, 4, 8), (512, 8, 3), (1080, 4, 8)])
def test_shard_rows_partition(height, tile, world):
    parts = [shard_rows(height, tile, r, world) for r in range(world)]
    allr = np.sort(np.concatenate(parts))
    assert np.array_equal(allr, np.arange(height))
    for r, p in enumerate(parts):
        assert np.all(np.diff(p) > 0) and np.all((p // tile) % world == r)
    assert max(map(len, parts)) - min(map(len, parts)) <= tile


def test_shard_rows_bad_arguments():
    for args in [(10, 0, 0, 1), (10, 4, 1, 1), (10, 4, 0, 0)]:
        with pytest.raises(ValueError):
            shard_rows(*args)


def test_ops_model_matches_survey_magnitude(demo2):
    """demo2 needs ~920 algorithmic FP64 ops per sample (oracle counters x op table)."""
    from oracle import oracle_py as O
    cfg = JobConfiguration(2, 5, 50)
    ss = O.generate_samples(1, 2, 5, 800)
    ss.set_index = O.generate_set_index(1, 600, 800, 800)
    rows = np.arange(0, 600, 20)
    _, cn = O.render_row_list(demo2.flatten(), cfg, ss, rows, counters=True)
    assert cn["samples"] == 30 * 800 * 4
    assert 880 < ops_per_sample(cn) < 960
    assert algorithmic_ops(cn) == pytest.approx(ops_per_sample(cn) * cn["samples"])
    assert cn["segments"] == cn["emissive"] + cn["matte"] + cn["glossy"] + cn["specular"] + cn["miss"]
    assert cn["hit_sphere"] + cn["hit_plane"] + cn["miss"] == cn["segments"]
    assert set(_capi.COUNTER_FIELDS) == set(cn)


# ---- BVH builder (EXTENSION; host code inside libfluxb200.so, no device needed) ----
def _describe(flat):
    import ctypes
    from flux_b200 import _capi
    out = (ctypes.c_uint64 * 8)()
    assert _capi.lib().flux_bvh_describe(flat.ptr(), out) == 0
    return dict(zip(("nodes", "levels", "leaf", "linear", "refs", "violations", "miscount", "auto"), list(out)))


def test_bvh_builder_invariants_sphere_cloud():
    from flux_b200 import synth
    flat = synth.sphere_cloud_scene(10_000, seed=5).flatten()
    d = _describe(flat)
    assert d["violations"] == 0 and d["miscount"] == 0
    assert d["refs"] + d["linear"] == 10_000 and d["linear"] == 0
    assert d["auto"] == 1 and 3 * d["levels"] + 1 <= 32 and d["leaf"] == 1   # leaves of one primitive since r2 (+14 % on config 5)
    assert d["nodes"] < 10_000


def test_bvh_builder_keeps_oversized_spheres_linear_and_handles_meshes():
    from flux_b200 import synth
    flat = synth.mesh_scene(100, 60, seed=3, width=64, height=48).flatten()
    d = _describe(flat)
    assert d["violations"] == 0 and d["miscount"] == 0
    assert d["linear"] == 2 and d["refs"] == 2 * 100 * 60      # environment sphere + light stay out of the tree
    assert d["auto"] == 1


def test_bvh_builder_small_and_degenerate_scenes(demo2):
    d = _describe(demo2.flatten())
    assert d["violations"] == 0 and d["miscount"] == 0 and d["auto"] == 0   # 12 spheres: linear scan in auto mode
    from flux_b200.scene import SceneData, SphereData, Matte
    m = Matte((0.5, 0.5, 0.5), (1, 1, 1), 1.0)
    same = [SphereData((1.0, 2.0, 3.0), 0.5, m, False) for _ in range(200)]   # identical centroids: median split must terminate
    sd = SceneData("same", demo2.output_settings, (0, 0, 0), same, demo2.camera_settings, demo2.camera_data)
    d = _describe(sd.flatten())
    assert d["violations"] == 0 and d["miscount"] == 0 and d["refs"] == 200
    one = SceneData("one", demo2.output_settings, (0, 0, 0), same[:1], demo2.camera_settings, demo2.camera_data)
    d = _describe(one.flatten())
    assert d["nodes"] == 1 and d["refs"] == 1 and d["miscount"] == 0


def test_bvh_builder_is_deterministic_and_plans_its_depth():
    """The median splits run on several host threads; the tree must not depend on that (flux_bvh_hash covers every
    byte the traversal reads), and the leaf size is chosen from the depth the builder plans arithmetically, which it
    checks against the depth it then builds (an error otherwise).  2.5 M spheres need leaves of 4."""
    import ctypes as C
    from flux_b200 import _capi, synth

    def hash_of(ptr):
        out = C.c_uint64()
        assert _capi.lib().flux_bvh_hash(ptr, C.byref(out)) == 0
        return out.value

    for sd in (synth.mesh_scene(300, 200, seed=3), synth.sphere_cloud_scene(100_000, seed=6)):   # 120 K triangles: threads in play
        flat = sd.flatten()
        assert len({hash_of(flat.ptr()) for _ in range(4)}) == 1
    a, b = synth.sphere_cloud_scene(5000, seed=1).flatten(), synth.sphere_cloud_scene(5000, seed=2).flatten()
    assert hash_of(a.ptr()) != hash_of(b.ptr())
    # leaves of 4: straight into a flux_scene_flat, without two and a half million Python objects
    n = 2_500_000
    rng = np.random.default_rng(7)
    c = np.ascontiguousarray(rng.uniform(-50, 50, (n, 3)))
    r = np.ascontiguousarray(rng.uniform(0.01, 0.05, n))
    inv, sid, mat = np.zeros(n, np.uint8), np.arange(n, dtype=np.uint32), np.zeros(n, np.uint32)
    s = _capi.flux_scene_flat()
    m = (_capi.flux_material * 1)()
    s.n_spheres, s.sphere_center, s.sphere_radius = n, _capi.as_dp(c), _capi.as_dp(r)
    s.sphere_invert, s.sphere_shape_id, s.sphere_material = inv.ctypes.data_as(C.POINTER(C.c_uint8)), _capi.as_u32p(sid), _capi.as_u32p(mat)
    s.n_materials, s.materials = 1, m
    d = (C.c_uint64 * 8)()
    assert _capi.lib().flux_bvh_describe(C.byref(s), d) == 0
    nodes, levels, leaf, linear, prims, violations, miscounted, used = list(d)
    assert (levels, leaf, prims, violations, miscounted) == (10, 4, n, 0, 0)


def test_bvh_builder_writes_every_byte_it_hands_over():
    """The builder's large arrays are not zeroed before the slices fill them (flux_raw_vector, host_slices.h).  glibc's
    MALLOC_PERTURB_ fills every allocation with a byte pattern instead of leaving the fresh zero pages of a large
    mmap: a byte of the node, reference or record arrays that no slice wrote would change the hash."""
    import subprocess
    import sys
    code = ("import ctypes as C\nfrom flux_b200 import _capi, synth\n"
            "for sd in (synth.mesh_scene(300, 200, seed=3), synth.sphere_cloud_scene(20_000, seed=6)):\n"
            "    out = C.c_uint64(); flat = sd.flatten()\n"
            "    assert _capi.lib().flux_bvh_hash(flat.ptr(), C.byref(out)) == 0\n"
            "    print(out.value)\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    seen = set()
    for perturb in ("0", "165", "77"):
        env = dict(os.environ, MALLOC_PERTURB_=perturb, PYTHONPATH=root)
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        seen.add(r.stdout)
    assert len(seen) == 1


def test_bvh_builder_names_the_first_triangle_with_a_non_finite_vertex():
    """NaN and infinite vertices are refused by the builder (a triangle the linear scan can never hit has no box to
    put in a tree), and the message names the lowest such triangle whichever host thread met it."""
    import ctypes as C
    from flux_b200 import _capi, synth
    from flux_b200 import Matte, SceneData, TriangleData
    sd = synth.mesh_scene(300, 200, seed=3)           # 120 000 triangles: the boxes are made in slices on several threads
    m = Matte((0.5, 0.5, 0.5), (1, 1, 1), 1.0)
    bad = [TriangleData((0, 0, 0), (1, float("nan"), 0), (0, 1, 0), m), TriangleData((0, 0, 0), (1, float("inf"), 0), (0, 1, 0), m)]
    out = (C.c_uint64 * 8)()
    for extra, first in ((bad, 120000), (bad[::-1], 120000), (bad[1:], 120000)):
        flat = SceneData("bad", sd.output_settings, sd.background, list(sd.shapes) + extra, sd.camera_settings, sd.camera_data).flatten()
        assert _capi.lib().flux_bvh_describe(flat.ptr(), out) != 0
        assert _capi.lib().flux_last_error(None).decode() == f"bvh: triangle {first} has a non-finite vertex"


def test_reference_arm_under_torchrun_prints_one_line_from_rank_zero():
    """The driver launches `bench.py --impl reference` the way it launches our arm — under torchrun for N > 1.  Rank 0
    alone runs the CPU oracle and prints the one JSON line; the other ranks exit 0 without work.  Both arms print the
    same `config` object; what this arm rendered per step is under `ran`."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29577", os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0", "--cpu-root", "3"], capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["unit"] == "Msamples/s" and d["higher_is_better"] is True
    assert d["ran"]["sample_root"] == 3 and d["ran"]["spp"] == 9
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    sys.path.insert(0, root)
    import bench
    assert d["config"] == bench.workload_config(2)   # the very object our arm prints at 2 GPUs
