"""Algorithmic FP64 operation model of the render path (normative).

Every add/sub/mul/div/sqrt/compare/select/pow/sin/cos counts 1; negation and x*(+-1.0) are
free.  Traversal terms are the minimum that reproduces the reference bit-for-bit with ray-
invariant terms hoisted; shading terms are as written per event.  Inputs are the event
counters of ``flux_counters`` (device) or the oracle's identical counters.
"""
from __future__ import annotations

OPS = {
    "samples": 48 + 3,     # ray generation (trace.rs:72-80, 44-51) + colour accumulate (trace.rs:82)
    "segments": 13,        # 3 reciprocals + 3 sign tests + d.d (5) + 2a + 4a
    "bbox_tests": 18,      # 6 sub + 6 mul + 4 select + 2 cmp (shapes.rs:98-133)
    "bbox_pass": 19,       # temp 3 + B 6 + C 6 + disc 3 + cmp 1 (shapes.rs:176-183)
    "disc_nonneg": 4,      # sqrt, sub, div, cmp (shapes.rs:186-190)
    "t2_evals": 3,         # add, div, cmp (shapes.rs:200-201)
    "plane_tests": 15,     # shapes.rs:137-139
    "tri_tests": 50,       # extension: Moller-Trumbore as written in the oracle
    "candidates": 1,       # Hit::compare (common.rs:17-23)
    "hit_sphere": 15,      # normal + point (shapes.rs:195-196)
    "hit_plane": 6,        # point (shapes.rs:145)
    "hit_tri": 24,         # extension: point + normalize(cross(e1,e2))
    "emissive": 12,        # materials.rs:44-48
    "matte": 78,           # materials.rs:19-33 + brdf.rs:20-30
    "specular": 40,        # materials.rs:57-71 + brdf.rs:39-45
    "glossy": 122,         # materials.rs:57-71 + brdf.rs:55-78 + to_unit_hemi
    "glossy_flip": 15,     # brdf.rs:67-71
}


def algorithmic_ops(counters: dict) -> float:
    """Total algorithmic FP64 operations for a set of event counters."""
    return float(sum(OPS[k] * counters.get(k, 0) for k in OPS))


def ops_per_sample(counters: dict) -> float:
    s = counters.get("samples", 0)
    return algorithmic_ops(counters) / s if s else 0.0
