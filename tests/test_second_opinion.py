"""The two CPU restatements of the reference's path — oracle/flux_oracle.cpp (C++, what the GPU is checked against) and
oracle/second_opinion.py (scalar Python, written a second time straight from the Rust sources) — must agree bit for
bit.  CPU only; sizes are a few thousand paths (the Python one costs ~50 microseconds per path)."""
import os

import numpy as np
import pytest

from flux_b200 import JobConfiguration, SceneData
from oracle import oracle_py as O
from oracle import second_opinion as S2
from tests import helpers as Hp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _both(sd, root, depth, rows, seed=1, num_sets=None):
    cfg = JobConfiguration(root, depth, 50)
    W, H = sd.output_settings.image_width, sd.output_settings.image_height
    ss = Hp.oracle_samples(seed, cfg, W, H, num_sets)
    a = O.render_row_list(sd.flatten(), cfg, ss, rows)
    b = np.asarray(S2.render_rows(sd, root, depth, ss.pixel, ss.disc, ss.hemi, ss.set_index, list(rows)), np.float64)
    return a, b


def _assert_same_bits(a, b):
    assert a.shape == b.shape
    same = a.view(np.uint64) == b.view(np.uint64)
    both_nan = np.isnan(a) & np.isnan(b)        # a NaN's payload and sign are not the reference's business
    assert np.all(same | both_nan), f"{np.count_nonzero(~(same | both_nan))} of {a.size} values differ; max |a-b| = {np.nanmax(np.abs(a - b))}"


@pytest.mark.parametrize("scene", ["demo1.yml", "demo2.yml"])
def test_shipped_scenes_bit_for_bit(scene):
    """The reference's own scenes (matte, emissive and glossy materials, thin lens), depth 5, 9 samples per pixel."""
    sd = SceneData.from_yaml(os.path.join(ROOT, "scenes", scene)).with_size(40, 30)
    a, b = _both(sd, 4, 5, range(30))
    _assert_same_bits(a, b)
    assert np.count_nonzero(a) > a.size // 2


def test_all_four_materials_bit_for_bit():
    """Perfect specular (in no shipped scene) beside the other three, mirror plane included, lens blur on."""
    sd = Hp.mixed_material_scene(28, 20)
    a, b = _both(sd, 3, 5, range(20), seed=7)
    _assert_same_bits(a, b)


def test_depth_cut_background_and_fewer_sets_than_columns():
    sd = SceneData.from_yaml(os.path.join(ROOT, "scenes", "demo2.yml")).with_size(20, 10)
    for depth in (0, 1, 2):
        a, b = _both(sd, 2, depth, [0, 4, 9], seed=3, num_sets=7)
        _assert_same_bits(a, b)


def test_closest_hit_ids_and_distances_bit_for_bit():
    """Scene::hit alone: 3000 random rays against 300 random spheres — bounding-box rejections, second roots from
    inside, ties impossible to tell apart from order — ids equal, distances the same bits."""
    rng = np.random.default_rng(11)
    sd = Hp.random_sphere_scene(rng, 300, extent=8.0, rmin=0.2, rmax=1.5)
    o, d = Hp.random_rays(rng, 3000, extent=9.0)
    # IEEE corners of the slab test: zero and negative-zero direction components (1 / d = +-inf, 0 * inf = NaN on a slab
    # face), axis-parallel rays through sphere centres, origins on a centre
    c = np.array([sh.center for sh in sd.shapes[:200]])
    d[:200] = 0.0
    d[np.arange(200), np.arange(200) % 3] = np.where(np.arange(200) % 2 == 0, 1.0, -1.0)
    d[100:200][d[100:200] == 0.0] = -0.0
    o[:200] = c
    o[:150, :] -= 5.0 * d[:150]
    o[50:100, 1] += np.array([sh.radius for sh in sd.shapes[50:100]])          # grazing: on the box face
    ids, t = O.trace_rays(sd.flatten(), o, d)
    got = S2.trace(sd, o, d)
    ids2 = np.array([g[0] for g in got], np.int32)
    t2 = np.array([g[1] for g in got], np.float64)
    assert np.array_equal(ids, ids2)
    hit = ids >= 0
    assert hit.sum() > 500
    assert np.array_equal(t[hit].view(np.uint64), t2[hit].view(np.uint64))


def test_sample_maps_bit_for_bit():
    """The two maps between sample domains — unit square to unit disc (Shirley, samplers/src/lib.rs:144-182) and to the
    hemisphere with exponent e (lib.rs:133-142) — on random points, the branch boundaries of the disc map and the
    exponents the scenes use."""
    rng = np.random.default_rng(5)
    pts = rng.random((2000, 2))
    pts[:8] = [(0.5, 0.5), (0.0, 0.0), (1.0, 1.0), (0.25, 0.75), (0.75, 0.25), (0.5, 0.0), (0.0, 0.5), (0.5, 1.0)]
    for x, y in pts:
        a, b = O.to_poisson_disc(float(x), float(y)), np.array(S2.to_poisson_disc(float(x), float(y)))
        assert np.array_equal(a.view(np.uint64), b.view(np.uint64)), (x, y, a, b)
    for e in (0.0, 1.0, 10.0, 50.0, 5000.0, 100000.0):
        for x, y in pts[:400]:
            a, b = O.to_unit_hemi(float(x), float(y), e), np.array(S2.to_unit_hemi(float(x), float(y), e))
            assert np.array_equal(a.view(np.uint64), b.view(np.uint64)), (x, y, e, a, b)


@pytest.mark.parametrize("correlated", [False, True])
def test_multi_jittered_structure_bit_for_bit(correlated):
    """grid_multi_jittered / grid_correlated_multi_jittered (samplers/src/lib.rs:46-126) written with the reference's
    shuffles and transposes, against the C++ oracle's closed form, on the same keyed random source: which base point's x
    and which base point's y end up in cell (i, j)."""
    for root in (1, 2, 3, 5, 8, 13):
        for set_, grid in ((0, 0), (3, 1), (7, 4)):
            a = O.mj_grid(9, root, set_, grid, correlated)
            b = np.array(S2.mj_grid(9, root, set_, grid, correlated), np.float64).reshape(-1, 2)
            assert np.array_equal(a.view(np.uint64), b.view(np.uint64)), (root, set_, grid)
