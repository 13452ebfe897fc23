"""EXTENSION parity (BASELINE configs 3 and 5): the BVH must return exactly what the
reference's linear scan (scene.rs:156-160) returns — same shape id, same bits of t — and triangle meshes
must render like the oracle's linear scan over the same triangles."""
import numpy as np
import pytest

from flux_b200 import JobConfiguration, synth
from flux_b200.scene import Emissive, Matte, MeshData, PlaneData, SceneData, SphereData, TriangleData
from oracle import oracle_py as O
from tests import helpers as Hp

pytestmark = pytest.mark.gpu


def _trace(ctx, flat, o, d, mode):
    ctx.set_accel_mode(mode)
    try:
        ctx.set_scene(flat, JobConfiguration(1))
        return ctx.trace_rays(o, d)
    finally:
        ctx.set_accel_mode(0)


def _same(a, b):
    assert np.array_equal(a[0], b[0])
    assert np.array_equal(a[1].view(np.uint64), b[1].view(np.uint64))


def test_bvh_matches_linear_scan_10k_spheres_4m_rays(gpu_ctx):
    """Config 5 at 4 M rays: BVH vs the GPU linear scan (bitwise), and vs the CPU oracle on a prefix."""
    flat = synth.sphere_cloud_scene(10_000, seed=5).flatten()
    o, d = synth.random_rays(4_000_000, seed=5)
    lin = _trace(gpu_ctx, flat, o, d, 1)
    bvh = _trace(gpu_ctx, flat, o, d, 2)
    _same(bvh, lin)
    ho, to = O.trace_rays(flat, o[:100_000], d[:100_000])
    assert np.array_equal(bvh[0][:100_000], ho)
    assert np.array_equal(bvh[1][:100_000].view(np.uint64), to.view(np.uint64))
    assert 0.02 < (bvh[0] >= 0).mean() < 0.98


def test_ray_batch_kernel_without_node_texture_is_bitwise_the_same(gpu_ctx, monkeypatch):
    """The ray-batch kernel fetches four of a node's seven 16-byte pieces through the texture unit (flux_bvh.cuh
    TRACE_TEX_MASK); a tree that cannot be bound as a linear texture runs the LSU-only instantiation.  Both return the
    same ids and the same bits of t."""
    flat = synth.sphere_cloud_scene(10_000, seed=7).flatten()
    o, d = synth.random_rays(1_000_000, seed=7)
    with_tex = _trace(gpu_ctx, flat, o, d, 2)
    monkeypatch.setenv("FLUXB200_NO_NODE_TEXTURE", "1")
    without = _trace(gpu_ctx, flat, o, d, 2)
    monkeypatch.delenv("FLUXB200_NO_NODE_TEXTURE")
    _same(with_tex, without)
    assert 0.02 < (with_tex[0] >= 0).mean() < 0.98


def test_bvh_ieee_corner_rays(gpu_ctx):
    """Axis-parallel rays (1/0 = inf and 0*inf = NaN slabs), rays starting inside spheres, on box faces, zero and
    denormal directions, unnormalised directions: NaN slabs must never cull."""
    rng = np.random.default_rng(11)
    sd = Hp.random_sphere_scene(rng, 3000, extent=8.0, rmin=0.2, rmax=0.9)
    flat = sd.flatten()
    c = np.array([s.center for s in sd.shapes])
    r = np.array([s.radius for s in sd.shapes])
    axes = np.concatenate([np.eye(3), -np.eye(3)])
    o_list, d_list = [], []
    for k in range(400):
        i = k % len(r)
        for ax in axes:
            o_list.append(c[i] - ax * 20.0); d_list.append(ax)                 # through the centre
            o_list.append(c[i] - ax * 20.0 + np.roll(ax, 1) * r[i]); d_list.append(ax)  # along a box face / tangent
            o_list.append(c[i]); d_list.append(ax)                              # from the centre (t2 branch)
            o_list.append(c[i] - r[i]); d_list.append(ax * 3.5)                 # from a box corner, unnormalised
    o = np.array(o_list); d = np.array(d_list) + 0.0
    d[::7][d[::7] == 0.0] = -0.0                                                # negative zeros flip the slab order
    extra_o = rng.uniform(-9, 9, (2000, 3)); extra_d = rng.standard_normal((2000, 3))
    extra_d[:500, 0] = 0.0; extra_d[500:1000, 1] = 1e-310; extra_d[1000:1010] = 0.0
    o = np.concatenate([o, extra_o]); d = np.concatenate([d, extra_d])
    lin = _trace(gpu_ctx, flat, o, d, 1)
    bvh = _trace(gpu_ctx, flat, o, d, 2)
    _same(bvh, lin)
    ho, to = O.trace_rays(flat, o, d)
    assert np.array_equal(lin[0], ho)
    assert np.array_equal(lin[1].view(np.uint64), to.view(np.uint64))


def test_bvh_ties_go_to_lower_shape_id(gpu_ctx):
    """Coincident and overlapping spheres: equal t must resolve to the earlier shape whatever the traversal order."""
    rng = np.random.default_rng(12)
    base = Hp.random_sphere_scene(rng, 500, extent=6.0, rmin=0.3, rmax=1.0)
    shapes = list(base.shapes) + list(base.shapes[:250]) + list(base.shapes[100:200])   # exact duplicates, later ids
    sd = SceneData("dups", base.output_settings, base.background, shapes, base.camera_settings, base.camera_data)
    flat = sd.flatten()
    o, d = Hp.random_rays(rng, 300_000, extent=8.0)
    lin = _trace(gpu_ctx, flat, o, d, 1)
    bvh = _trace(gpu_ctx, flat, o, d, 2)
    _same(bvh, lin)
    ho, _ = O.trace_rays(flat, o[:50_000], d[:50_000])
    assert np.array_equal(bvh[0][:50_000], ho)
    assert (bvh[0][bvh[0] >= 0] < 500).all()  # a duplicate never wins


def test_bvh_mixed_spheres_triangles_planes_and_oversized(gpu_ctx):
    """Triangles + spheres in one tree, planes and the environment sphere on the linear list."""
    rng = np.random.default_rng(13)
    mesh = synth.heightfield_mesh(60, 40, seed=3, extent=10.0)
    m = Matte((0.5, 0.5, 0.5), (1, 1, 1), 1.0)
    shapes = [SphereData((0, 0, 0), 100.0, Emissive((1, 1, 1), 0.3), True), PlaneData((0, -1.0, 0), (0, 1, 0), m), mesh]
    cs = rng.uniform(-9, 9, (300, 3)); cs[:, 1] = rng.uniform(0.0, 3.0, 300)
    shapes += [SphereData(tuple(cs[i]), float(rng.uniform(0.1, 0.6)), m, False) for i in range(300)]
    shapes += [TriangleData(tuple(rng.uniform(-9, 9, 3)), tuple(rng.uniform(-9, 9, 3)), tuple(rng.uniform(-9, 9, 3)), m)
               for _ in range(50)]
    sd = SceneData("mixed_bvh", Hp.deterministic_scene().output_settings, (0, 0, 0), shapes,
                   Hp.deterministic_scene().camera_settings, Hp.deterministic_scene().camera_data)
    flat = sd.flatten()
    o, d = Hp.random_rays(rng, 400_000, extent=11.0)
    # rays along mesh edges / through vertices: the box padding must cover Moller-Trumbore's edge decisions
    v = mesh.vertices[rng.integers(0, len(mesh.vertices), 20_000)]
    o2 = v + np.array([0.0, 7.0, 0.0]); d2 = np.tile([0.0, -1.0, 0.0], (len(v), 1))
    o = np.concatenate([o, o2]); d = np.concatenate([d, d2])
    lin = _trace(gpu_ctx, flat, o, d, 1)
    bvh = _trace(gpu_ctx, flat, o, d, 2)
    _same(bvh, lin)
    ho, to = O.trace_rays(flat, o[:20_000], d[:20_000])
    assert np.array_equal(bvh[0][:20_000], ho)
    assert np.array_equal(bvh[1][:20_000].view(np.uint64), to.view(np.uint64))
    assert (bvh[0] >= 0).all()  # inside the environment sphere


def test_bvh_forced_on_demo2_renders_bitwise_like_linear(gpu_ctx, demo2):
    """Same per-sample arithmetic either way: the direct kernel with BVH = the direct kernel with the linear scan."""
    sd = demo2.with_size(96, 72)
    cfg = JobConfiguration(4, 5, 50)
    ss = Hp.oracle_samples(3, cfg, 96, 72)
    flat = sd.flatten()
    gpu_ctx.set_kernel_mode(1)
    try:
        Hp.upload(gpu_ctx, flat, cfg, ss)
        a = gpu_ctx.render_rows(0, 71, 96)
        gpu_ctx.set_accel_mode(2)
        Hp.upload(gpu_ctx, flat, cfg, ss)
        b = gpu_ctx.render_rows(0, 71, 96)
    finally:
        gpu_ctx.set_accel_mode(0)
        gpu_ctx.set_kernel_mode(0)
    assert np.array_equal(a.view(np.uint64), b.view(np.uint64))


def test_mesh_scene_renders_like_the_oracle(gpu_ctx):
    """Config 3 shape at a size the oracle's linear scan finishes: 2 x 48 x 32 = 3072 triangles (BVH on the GPU,
    linear scan in the oracle), matte mesh under emissive spheres, 64x48 at 16 spp.  No transcendental on the
    path, so radiance agrees to the pixel-sum order (1e-13)."""
    sd = synth.mesh_scene(48, 32, seed=3, width=64, height=48)
    cfg = JobConfiguration(4, 5, 50)
    ss = Hp.oracle_samples(5, cfg, 64, 48)
    flat = sd.flatten()
    Hp.upload(gpu_ctx, flat, cfg, ss)
    gpu_ctx.enable_counters(True)
    gpu_ctx.reset_counters()
    img = gpu_ctx.render_rows(0, 47, 64)
    cn = gpu_ctx.counters()
    gpu_ctx.enable_counters(False)
    ref, ocn = O.render_rows(flat, cfg, ss, 0, 47, counters=True)
    assert Hp.rel_err(img, ref) <= 1e-12
    for k in ("samples", "segments", "hit_tri", "hit_sphere", "emissive", "matte", "miss", "depth_cut"):
        assert cn[k] == ocn[k], k
    assert cn["hit_tri"] > 0 and cn["nodes_visited"] > 0
    assert cn["tri_tests"] < ocn["tri_tests"] / 20   # the point of the tree


def test_full_size_mesh_bvh_vs_linear_on_gpu(gpu_ctx):
    """Config 3 at full size: 1,000,000 triangles.  BVH vs the GPU's own linear scan on 20 K rays (2e10 triangle
    tests), bitwise."""
    sd = synth.mesh_scene(1000, 500, seed=3)
    flat = sd.flatten()
    assert flat.struct.n_triangles == 1_000_000
    rng = np.random.default_rng(14)
    o = np.stack([rng.uniform(-10, 10, 20_000), rng.uniform(1.0, 6.0, 20_000), rng.uniform(-10, 10, 20_000)], axis=1)
    d = rng.standard_normal((20_000, 3)); d[:, 1] = -np.abs(d[:, 1]) - 0.05
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    lin = _trace(gpu_ctx, flat, o, d, 1)
    bvh = _trace(gpu_ctx, flat, o, d, 2)
    _same(bvh, lin)
    assert (bvh[0] >= 2).mean() > 0.4   # a good share of the rays land on the mesh
    assert (bvh[0] >= 0).all()          # the rest on the environment sphere


def test_mesh_scene_regeneration_kernel_with_bvh(gpu_ctx):
    """spp >= 64 on a BVH scene selects the regeneration kernel with BVH traversal (warp per pixel, in-warp path
    regeneration): must agree with the oracle's linear scan and with the direct BVH kernel's counters."""
    sd = synth.mesh_scene(40, 24, seed=3, width=32, height=24)
    cfg = JobConfiguration(8, 5, 50)
    ss = Hp.oracle_samples(9, cfg, 32, 24)
    flat = sd.flatten()
    Hp.upload(gpu_ctx, flat, cfg, ss)
    imgs, cns = {}, {}
    try:
        for mode in (1, 2):   # direct, regeneration — both through the BVH (1922 shapes)
            gpu_ctx.set_kernel_mode(mode)
            gpu_ctx.enable_counters(True)
            gpu_ctx.reset_counters()
            imgs[mode] = gpu_ctx.render_rows(0, 23, 32)
            cns[mode] = gpu_ctx.counters()
            gpu_ctx.enable_counters(False)
    finally:
        gpu_ctx.set_kernel_mode(0)
        gpu_ctx.enable_counters(False)
    ref = O.render_rows(flat, cfg, ss, 0, 23)
    assert Hp.rel_err(imgs[1], ref) <= 1e-12
    assert Hp.rel_err(imgs[2], ref) <= 1e-12
    assert cns[1] == cns[2] and cns[2]["nodes_visited"] > 0 and cns[2]["hit_tri"] > 0
    gpu_ctx.set_kernel_mode(0)
    auto = gpu_ctx.render_rows(0, 23, 32)
    assert np.array_equal(auto.view(np.uint64), imgs[2].view(np.uint64))   # auto mode = regeneration kernel here


def test_mesh_scene_wavefront_kernel_with_bvh(gpu_ctx):
    """spp >= 256 on a BVH scene selects the wavefront kernel whose owner stage traverses the BVH: must agree with
    the oracle's linear scan (matte mesh, emissive spheres: 1e-12) and reproduce the regeneration kernel's counters."""
    sd = synth.mesh_scene(40, 24, seed=3, width=24, height=16)
    cfg = JobConfiguration(16, 5, 50)
    ss = Hp.oracle_samples(10, cfg, 24, 16)
    flat = sd.flatten()
    Hp.upload(gpu_ctx, flat, cfg, ss)
    imgs, cns = {}, {}
    try:
        for mode in (2, 4):
            gpu_ctx.set_kernel_mode(mode)
            gpu_ctx.enable_counters(True)
            gpu_ctx.reset_counters()
            imgs[mode] = gpu_ctx.render_rows(0, 15, 24)
            cns[mode] = gpu_ctx.counters()
            gpu_ctx.enable_counters(False)
    finally:
        gpu_ctx.set_kernel_mode(0)
        gpu_ctx.enable_counters(False)
    ref = O.render_rows(flat, cfg, ss, 0, 15)
    assert Hp.rel_err(imgs[4], ref) <= 1e-12
    assert Hp.rel_err(imgs[2], ref) <= 1e-12
    assert cns[2] == cns[4] and cns[4]["hit_tri"] > 0 and cns[4]["nodes_visited"] > 0
    auto = gpu_ctx.render_rows(0, 15, 24)
    assert np.array_equal(auto.view(np.uint64), imgs[4].view(np.uint64))   # auto mode = wavefront kernel here


def test_glossy_67_sphere_scene_wavefront_bvh_matches_oracle(gpu_ctx):
    """Config 4 shape: all four shading kinds, 67 spheres.  The automatic choice is the wavefront kernel's FP32-culled
    linear scan (faster than its BVH owner stage up to ~120 spheres); asking for the BVH runs the BVH owner stage.
    Both against the oracle, and bit-identical to each other."""
    sd = synth.glossy_scene(24, 14, seed=4)
    cfg = JobConfiguration(16, 5, 50)
    flat = sd.flatten()
    ss = Hp.oracle_samples(24, cfg, 24, 14)
    ref, cn_o = O.render_rows(flat, cfg, ss, 0, 13, counters=True)
    imgs = {}
    try:
        for accel in (0, 2):
            gpu_ctx.set_accel_mode(accel)
            Hp.upload(gpu_ctx, flat, cfg, ss)
            gpu_ctx.enable_counters(True)
            gpu_ctx.reset_counters()
            imgs[accel] = gpu_ctx.render_rows(0, 13, 24)
            cn = gpu_ctx.counters()
            gpu_ctx.enable_counters(False)
            assert Hp.rel_err(imgs[accel], ref) <= 1e-6
            assert (cn["nodes_visited"] > 0) == (accel == 2)
            for k in ("samples", "segments", "hit_sphere", "hit_plane", "emissive", "matte", "specular", "glossy", "miss", "depth_cut"):
                assert abs(cn[k] - cn_o[k]) <= max(2, 1e-6 * cn_o[k]), (accel, k, cn[k], cn_o[k])
    finally:
        gpu_ctx.set_accel_mode(0)
        gpu_ctx.enable_counters(False)
    assert np.array_equal(imgs[0].view(np.uint64), imgs[2].view(np.uint64))
