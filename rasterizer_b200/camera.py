"""Camera matrices and occluder ordering for the frame loop (application side of the hot path).

Restates what the reference's demo shell feeds into the rasterizer each frame
(Main.cpp:172-178: viewProj = LookToLH x PerspectiveFovLH, 16 floats, row-vector
convention; Main.cpp:185-190: front-to-back sort by squared distance of the occluder
centre to the camera).  All arithmetic is float32.  The matrices are produced once and
handed to both the CUDA path and the oracle, so DirectXMath bit-compatibility is not needed.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32

# default cameras, Main.cpp:28-40
CASTLE_CAMERA = dict(pos=(27.0, 2.0, 47.0), dir=(0.142582759, 0.0611068942, -0.987894833), up=(0.0, 1.0, 0.0), fov=0.628)
SPONZA_CAMERA = dict(pos=(0.0, 0.0, 0.0), dir=(1.0, 0.0, 0.0), up=(0.0, 0.0, 1.0), fov=1.04)


def _normalize(v):
    v = np.asarray(v, dtype=f32)
    return (v / f32(np.sqrt(f32(np.dot(v, v))))).astype(f32)


def look_to_lh(pos, direction, up) -> np.ndarray:
    """Left-handed look-to view matrix (row-vector convention), float32 4x4."""
    pos = np.asarray(pos, dtype=f32)
    r2 = _normalize(direction)
    r0 = _normalize(np.cross(np.asarray(up, dtype=f32), r2).astype(f32))
    r1 = np.cross(r2, r0).astype(f32)
    neg = (-pos).astype(f32)
    m = np.zeros((4, 4), dtype=f32)
    m[:3, 0], m[:3, 1], m[:3, 2] = r0, r1, r2
    m[3, 0], m[3, 1], m[3, 2] = f32(np.dot(r0, neg)), f32(np.dot(r1, neg)), f32(np.dot(r2, neg))
    m[3, 3] = f32(1.0)
    return m


def perspective_fov_lh(fov: float, aspect: float, zn: float, zf: float) -> np.ndarray:
    """Left-handed perspective projection with depth in [0, 1] (row-vector convention)."""
    h = f32(np.cos(f32(0.5) * f32(fov))) / f32(np.sin(f32(0.5) * f32(fov)))
    w = f32(h / f32(aspect))
    rng = f32(f32(zf) / (f32(zf) - f32(zn)))
    m = np.zeros((4, 4), dtype=f32)
    m[0, 0], m[1, 1], m[2, 2], m[2, 3], m[3, 2] = w, h, rng, f32(1.0), f32(-rng * f32(zn))
    return m


def view_projection(pos, direction, up, fov, width, height, zn=1.0, zf=5000.0) -> np.ndarray:
    """The 16 floats Main.cpp:172-178 hands to setModelViewProjection (view x proj, row-major)."""
    v = look_to_lh(pos, direction, up)
    p = perspective_fov_lh(fov, f32(width) / f32(height), zn, zf)
    return (v @ p).astype(f32).reshape(16)


def front_to_back_order(centers: np.ndarray, cam_pos) -> np.ndarray:
    """Occluder order of Main.cpp:185-190: ascending dp(c-p, c-p) with the dpps 0x7f sum order
    (x*x + y*y) + (z*z + 0); stable, ties by index (the app's std::sort leaves ties unspecified)."""
    d = (np.asarray(centers, dtype=f32)[:, :3] - np.asarray(cam_pos, dtype=f32)[None, :3]).astype(f32)
    sq = (d * d).astype(f32)
    key = ((sq[:, 0] + sq[:, 1]).astype(f32) + sq[:, 2]).astype(f32)
    return np.argsort(key, kind="stable").astype(np.uint32)
