"""Shared helpers for parity tests (tests only)."""
import numpy as np

from flux_b200 import (Emissive, GlossyReflective, JobConfiguration, Matte, PlaneData, Reflective, SceneData,
                       SphereData, OutputSettings, CameraSettings, CameraData)
from oracle import oracle_py as O


def oracle_samples(seed, cfg, width, height, num_sets=None):
    num_sets = num_sets or width
    ss = O.generate_samples(seed, cfg.sample_root, cfg.max_trace_depth, num_sets)
    ss.set_index = O.generate_set_index(seed, height, width, num_sets)
    return ss


def upload(ctx, flat, cfg, ss):
    ctx.set_scene(flat, cfg)
    ctx.set_samples(ss.root, ss.max_depth, ss.num_sets, ss.pixel, ss.disc, ss.hemi)
    ctx.set_set_index(ss.set_index)


def rel_err(a, b):
    """max |a-b| / max(|b|, tiny) with NaNs required to coincide."""
    a, b = np.asarray(a), np.asarray(b)
    nan_a, nan_b = np.isnan(a), np.isnan(b)
    assert np.array_equal(nan_a, nan_b), "NaN pattern differs"
    m = ~nan_a
    denom = np.maximum(np.abs(b[m]), 1e-300)
    return float(np.max(np.abs(a[m] - b[m]) / denom)) if m.any() else 0.0


def random_sphere_scene(rng, n_spheres, extent=50.0, rmin=0.1, rmax=0.5, width=64, height=48):
    """BASELINE config 5 shape: n spheres, centres uniform in [-extent, extent]^3, radii uniform [rmin, rmax]."""
    mats = [Matte((0.5, 0.5, 0.5), (1, 1, 1), 1.0), Emissive((1, 1, 1), 1.0)]
    c = rng.uniform(-extent, extent, (n_spheres, 3))
    r = rng.uniform(rmin, rmax, n_spheres)
    shapes = [SphereData(tuple(c[i]), float(r[i]), mats[i % 2], False) for i in range(n_spheres)]
    return SceneData("random_spheres", OutputSettings(width, height, 0.5), (0.0, 0.0, 0.0), shapes,
                     CameraSettings((0.0, 0.0, -120.0), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0)),
                     CameraData(1.0, 500.0, 100.0, 0.0))


def random_rays(rng, n, extent=60.0):
    o = rng.uniform(-extent, extent, (n, 3))
    d = rng.standard_normal((n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return o, d


def mixed_material_scene(width=96, height=64):
    """All four material kinds incl. perfect specular (no shipped scene uses Reflective), with lens blur."""
    shapes = [
        SphereData((0, 0, 0), 100.0, Emissive((1, 0.9686, 0.8588), 0.4), True),
        SphereData((-6.0, 6.0, 4.0), 3.0, Emissive((1, 1, 1), 8.0), False),
        SphereData((-2.2, 1.0, 0.0), 1.0, Reflective(0.9, (0.9, 0.9, 1.0)), False),
        SphereData((0.0, 1.0, 1.5), 1.0, GlossyReflective(0.7, (1.0, 0.8, 0.6), 50.0), False),
        SphereData((2.2, 1.0, 0.0), 1.0, Matte((0.2, 0.7, 0.3), (1, 1, 1), 0.9), False),
        SphereData((0.0, 0.5, -2.0), 0.5, GlossyReflective(0.8, (0.9, 0.9, 0.9), 5000.0), False),
        PlaneData((0, 0, 0), (0, 1, 0), Matte((0.5, 0.5, 0.5), (1, 1, 1), 1.0)),
        PlaneData((0, 0, 12.0), (0, 0, -1), Reflective(0.8, (0.8, 0.9, 0.8))),
    ]
    return SceneData("mixed", OutputSettings(width, height, 0.5), (0.05, 0.05, 0.1), shapes,
                     CameraSettings((0.0, 3.0, -9.0), (0.0, 1.0, 0.0), (0.0, 1.0, 0.0)),
                     CameraData(6.0 * width / 800.0, 500.0, 10.0, 0.05))


def deterministic_scene(width=96, height=64):
    """Matte / Emissive / perfect-specular only: no transcendental on the path, so the
    device result must be bit-identical per sample (only the pixel sum order differs)."""
    shapes = [
        SphereData((0, 0, 0), 100.0, Emissive((1, 0.9686, 0.8588), 0.6), True),
        SphereData((-2.2, 1.0, 0.0), 1.0, Reflective(0.9, (0.9, 0.9, 1.0)), False),
        SphereData((2.2, 1.0, 0.0), 1.0, Matte((0.2, 0.7, 0.3), (1, 1, 1), 0.9), False),
        SphereData((0.0, 1.0, 2.0), 1.0, Matte((0.8, 0.3, 0.3), (1, 1, 1), 1.0), False),
        PlaneData((0, 0, 0), (0, 1, 0), Matte((0.5, 0.5, 0.5), (1, 1, 1), 1.0)),
    ]
    return SceneData("deterministic", OutputSettings(width, height, 0.5), (0.0, 0.0, 0.0), shapes,
                     CameraSettings((0.0, 3.0, -9.0), (0.0, 1.0, 0.0), (0.0, 1.0, 0.0)),
                     CameraData(6.0 * width / 800.0, 500.0, 10.0, 0.0))
