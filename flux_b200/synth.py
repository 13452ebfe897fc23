"""Synthetic scenes of the shapes BASELINE.json's configs name.

The reference ships two scenes (scenes/demo1.yml, demo2.yml: 6 and 13 shapes).  Configs 3-5 ask for
scene shapes that do not exist in the reference — a 1M-triangle mesh, a divergent glossy multi-bounce
scene, a 10K-sphere ray microbench — so they are generated here from seeded numpy PRNGs; every
generator is a pure function of its arguments (tests and bench call the same code).
"""
from __future__ import annotations

import numpy as np

from .scene import (CameraData, CameraSettings, Emissive, GlossyReflective, Matte, MeshData, OutputSettings, PlaneData,
                    Reflective, SceneData, SphereData)

_ENV_COLOR = (1.0, 0.9686, 0.8588)   # demo2.yml:39,47


def _hash_noise(ix: np.ndarray, iz: np.ndarray, seed: int) -> np.ndarray:
    """Deterministic per-vertex noise in [0, 1): splitmix64 finaliser of (ix, iz, seed)."""
    with np.errstate(over="ignore"):
        x = (ix.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15) + iz.astype(np.uint64) * np.uint64(0xC2B2AE3D27D4EB4F)
             + np.uint64(seed) * np.uint64(0x165667B19E3779F9))
        x ^= x >> np.uint64(30)
        x *= np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(27)
        x *= np.uint64(0x94D049BB133111EB)
        x ^= x >> np.uint64(31)
    return (x >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def heightfield_mesh(nx: int = 1000, nz: int = 500, seed: int = 3, extent: float = 10.0) -> MeshData:
    """Config 3 geometry: height field over [-extent, extent]^2 with nx x nz quads = 2*nx*nz triangles,
    y = 0.5 sin(0.7 x) cos(0.9 z) + 0.1 hash_noise(ix, iz, seed); Matte (0.5, 0.5, 0.5)."""
    xs = np.linspace(-extent, extent, nx + 1)
    zs = np.linspace(-extent, extent, nz + 1)
    ix, iz = np.meshgrid(np.arange(nx + 1), np.arange(nz + 1), indexing="ij")
    x, z = xs[ix], zs[iz]
    y = 0.5 * np.sin(0.7 * x) * np.cos(0.9 * z) + 0.1 * _hash_noise(ix, iz, seed)
    verts = np.stack([x, y, z], axis=-1).reshape(-1, 3)
    vid = (ix * (nz + 1) + iz)
    a, b = vid[:-1, :-1].ravel(), vid[1:, :-1].ravel()
    c, d = vid[1:, 1:].ravel(), vid[:-1, 1:].ravel()
    # wound so that cross(e1, e2) points up (+y): the normal is used as wound, never flipped
    faces = np.concatenate([np.stack([a, c, b], axis=1), np.stack([a, d, c], axis=1)], axis=1).reshape(-1, 3)
    return MeshData(verts, faces.astype(np.int64), Matte((0.5, 0.5, 0.5), (1.0, 1.0, 1.0), 1.0))


def mesh_scene(nx: int = 1000, nz: int = 500, seed: int = 3, width: int = 800, height: int = 600) -> SceneData:
    """Config 3: the height-field mesh under demo2's environment sphere, one emissive sphere light, demo2's
    camera (800x600); rendered at sample_root 32 (1024 spp), depth 5."""
    shapes = [
        SphereData((0.0, 0.0, 0.0), 100.0, Emissive(_ENV_COLOR, 0.3), True),
        SphereData((-9.0, 7.0, 8.0), 5.0, Emissive(_ENV_COLOR, 10.0), False),
        heightfield_mesh(nx, nz, seed),
    ]
    return SceneData("mesh1m", OutputSettings(width, height, 0.5 * 800.0 / width), (0.0, 0.0, 0.0), shapes,
                     CameraSettings((0.0, 5.5, -9.0), (0.0, 1.0, 0.0), (0.0, 1.0, 0.0)),
                     CameraData(1.0, 500.0, 10.0, 0.09))


def glossy_scene(width: int = 1920, height: int = 1080, seed: int = 4, grid: int = 8) -> SceneData:
    """Config 4: area-light + reflective/glossy multi-bounce scene with divergent shading: an 8x8 jittered grid of
    unit spheres whose materials cycle Matte / Glossy(10) / Glossy(100) / Glossy(10000) / Reflective, two emissive
    sphere lights (power 10, r = 3), an inverted environment sphere (power 0.3), a matte floor; lens radius 0.09.
    Rendered at sample_root 64 (4096 spp), depth 5."""
    rng = np.random.default_rng(seed)
    mats = [
        lambda c: Matte(c, (1.0, 1.0, 1.0), 0.9),
        lambda c: GlossyReflective(0.6, c, 10.0),
        lambda c: GlossyReflective(0.6, c, 100.0),
        lambda c: GlossyReflective(0.6, c, 10000.0),
        lambda c: Reflective(0.8, c),
    ]
    shapes = [
        SphereData((0.0, 0.0, 0.0), 200.0, Emissive(_ENV_COLOR, 0.3), True),
        SphereData((-14.0, 12.0, 6.0), 3.0, Emissive((1.0, 1.0, 1.0), 10.0), False),
        SphereData((14.0, 12.0, 10.0), 3.0, Emissive(_ENV_COLOR, 10.0), False),
    ]
    k = 0
    for gx in range(grid):   # config 4 is the 8 x 8 grid; other sizes serve the linear-scan / BVH break-even measurements
        for gz in range(grid):
            jx, jz = rng.uniform(-0.35, 0.35, 2)
            col = tuple(float(v) for v in rng.uniform(0.45, 1.0, 3))
            shapes.append(SphereData((float((gx - (grid - 1) / 2) * 2.9 + jx), 1.0, float(gz * 2.9 + jz - 2.0)), 1.0, mats[k % 5](col), False))
            k += 1
    shapes.append(PlaneData((0.0, 0.0, 0.0), (0.0, 1.0, 0.0), Matte((0.5, 0.5, 0.5), (1.0, 1.0, 1.0), 1.0)))
    return SceneData("glossy64", OutputSettings(width, height, 0.5 * 1920.0 / width), (0.0, 0.0, 0.0), shapes,
                     CameraSettings((0.0, 9.0, -16.0), (0.0, 1.0, 8.0), (0.0, 1.0, 0.0)),
                     CameraData(1.6, 500.0, 22.0, 0.09))


def sphere_cloud_scene(n_spheres: int = 10_000, seed: int = 5, extent: float = 50.0, rmin: float = 0.1,
                       rmax: float = 0.5) -> SceneData:
    """Config 5 scene: n spheres, centres uniform in [-extent, extent]^3, radii uniform [rmin, rmax], invert = false."""
    rng = np.random.default_rng(seed)
    c = rng.uniform(-extent, extent, (n_spheres, 3))
    r = rng.uniform(rmin, rmax, n_spheres)
    mats = [Matte((0.5, 0.5, 0.5), (1.0, 1.0, 1.0), 1.0), Emissive((1.0, 1.0, 1.0), 1.0)]
    shapes = [SphereData((float(c[i, 0]), float(c[i, 1]), float(c[i, 2])), float(r[i]), mats[i & 1], False)
              for i in range(n_spheres)]
    return SceneData("spheres10k", OutputSettings(64, 48, 0.5), (0.0, 0.0, 0.0), shapes,
                     CameraSettings((0.0, 0.0, -120.0), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0)),
                     CameraData(1.0, 500.0, 100.0, 0.0))


def random_rays(n: int, seed: int = 5, extent: float = 60.0, chunk_offset: int = 0):
    """Config 5 rays: origins uniform in [-extent, extent]^3, directions = normalised standard normals.
    `chunk_offset` selects an independent stream so that 100 M rays can be made in pieces."""
    rng = np.random.default_rng([seed, 0x52415953, chunk_offset])
    o = rng.uniform(-extent, extent, (n, 3))
    d = rng.standard_normal((n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return o, d
