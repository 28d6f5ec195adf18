"""Workload definitions for the BASELINE.json configs: prepared scenes, camera sets, occludee boxes.

A *prepared scene* is what Main.cpp:56-113 produces before baking: batches of quads (4 float4
vertices per quad, quad count a multiple of 8 per batch) plus one reference AABB.  `prepare_mesh`
runs those steps through the product's own preparation (orz_quad_decompose, orz_generate_batches,
bit-identical with the reference's QuadDecomposition / SurfaceAreaHeuristic); for Castle/Sponza the
prepared scene is stored once under scenes/_prepared/ (git-ignored: Castle is under the Intel Code
Samples License; the raw meshes only exist where /root/reference does).  Synthetic scenes are
generated here from a seed.
"""
from __future__ import annotations

import os
import struct
from dataclasses import dataclass, field

import numpy as np

from . import camera as cam

f32 = np.float32
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PREPARED_DIR = os.path.join(ROOT, "scenes", "_prepared")
MAGIC = b"ORZSCN1\0"


@dataclass
class PreparedScene:
    name: str
    batches: list            # list of float32 [nQuads*4, 4]
    ref_min: np.ndarray      # float32[4]
    ref_max: np.ndarray
    camera: dict = field(default_factory=dict)

    @property
    def n_quads(self) -> int:
        return sum(b.shape[0] // 4 for b in self.batches)

    def quad_boxes(self) -> np.ndarray:
        """Occludee set of SURVEY 8d: per-quad AABBs in batch order, w := 1 -> float32 [n, 8]."""
        out = []
        for b in self.batches:
            q = b.reshape(-1, 4, 4)
            mn, mx = q.min(axis=1), q.max(axis=1)
            mn[:, 3] = 1.0
            mx[:, 3] = 1.0
            out.append(np.concatenate([mn, mx], axis=1))
        return np.ascontiguousarray(np.concatenate(out).astype(f32))

    def save(self, path: str):
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "wb") as f:
            f.write(MAGIC)
            f.write(struct.pack("<II", len(self.batches), self.n_quads))
            f.write(np.asarray(self.ref_min, f32).tobytes())
            f.write(np.asarray(self.ref_max, f32).tobytes())
            f.write(np.array([b.shape[0] // 4 for b in self.batches], np.uint32).tobytes())
            for b in self.batches:
                f.write(np.ascontiguousarray(b, f32).tobytes())

    @classmethod
    def load(cls, path: str, name: str = "", camera: dict | None = None) -> "PreparedScene":
        with open(path, "rb") as f:
            data = f.read()
        if data[:8] != MAGIC:
            raise ValueError(f"{path}: not a prepared scene")
        nb, nq = struct.unpack_from("<II", data, 8)
        off = 16
        ref_min = np.frombuffer(data, f32, 4, off).copy()
        ref_max = np.frombuffer(data, f32, 4, off + 16).copy()
        off += 32
        counts = np.frombuffer(data, np.uint32, nb, off)
        off += 4 * nb
        batches = []
        for c in counts:
            n = int(c) * 16
            batches.append(np.frombuffer(data, f32, n, off).reshape(-1, 4).copy())
            off += 4 * n
        assert sum(int(c) for c in counts) == nq
        return cls(name or os.path.basename(path), batches, ref_min, ref_max, camera or {})


def prepared_path(name: str) -> str:
    return os.path.join(PREPARED_DIR, f"{name.lower()}.orzscn")


def have_scene(name: str) -> bool:
    return os.path.exists(prepared_path(name))


def load_scene(name: str) -> PreparedScene:
    """'castle' / 'sponza' (prepared files) or 'city' (synthetic, always available)."""
    key = name.lower()
    if key == "city":
        return synthetic_city()
    camera = {"castle": cam.CASTLE_CAMERA, "sponza": cam.SPONZA_CAMERA}[key]
    return PreparedScene.load(prepared_path(key), key, dict(camera))


# ------------------------------------------------------------------------------------------------
# Main.cpp:86-128 on a raw mesh
def pad_quads(quad_indices: np.ndarray) -> np.ndarray:
    """Main.cpp:91-94: repeat the first index until the quad count is a multiple of 8."""
    q = np.ascontiguousarray(quad_indices, np.uint32).reshape(-1)
    pad = (-q.size) % 32
    return np.concatenate([q, np.full(pad, q[0], np.uint32)]) if pad else q


def _fold_minmax(points: np.ndarray):
    """Aabb::include (VectorMath.h:44-48) over axis 1: minps / maxps keep the NEW operand on ties,
    which decides the sign of a zero bound."""
    mn = np.full((points.shape[0], points.shape[2]), np.inf, f32)
    mx = np.full_like(mn, -np.inf)
    for k in range(points.shape[1]):
        p = points[:, k]
        mn = np.where(mn < p, mn, p)
        mx = np.where(mx > p, mx, p)
    return mn, mx


def quad_aabbs(quad_indices: np.ndarray, vertices: np.ndarray) -> np.ndarray:
    """Main.cpp:96-105: one Aabb per quad -> float32 [n, 8] (min4, max4)."""
    v = np.ascontiguousarray(vertices, f32).reshape(-1, 4)
    mn, mx = _fold_minmax(v[quad_indices.reshape(-1, 4)])
    return np.ascontiguousarray(np.concatenate([mn, mx], axis=1))


def prepare_mesh(name: str, indices: np.ndarray, vertices: np.ndarray, camera: dict | None = None, target_size: int = 512,
                 split_granularity: int = 8, generate_batches=None) -> PreparedScene:
    """Triangle list + float4 vertices -> prepared scene, the steps of Main.cpp:86-128 before bake:
    quad decomposition, padding, per-quad AABBs, SAH batches, reference AABB over ALL vertices.
    `generate_batches`: the batching entry to use (default: host; `Context.generate_batches` for the GPU)."""
    from . import api

    v = np.ascontiguousarray(vertices, f32).reshape(-1, 4)
    quads = pad_quads(api.quad_decompose(indices, v))
    batching = generate_batches or api.generate_batches
    groups = batching(quad_aabbs(quads, v), target_size, split_granularity)
    q4 = quads.reshape(-1, 4)
    batches = [np.ascontiguousarray(v[q4[g].reshape(-1)]) for g in groups]
    ref_min, ref_max = _fold_minmax(v[None])
    return PreparedScene(name, batches, ref_min[0].copy(), ref_max[0].copy(), dict(camera or {}))


# ------------------------------------------------------------------------------------------------
# synthetic scenes
def _box_quads(lo, hi):
    """6 outward-facing quads of an axis-aligned box, same winding convention as the scenes'
    front faces (clockwise seen from outside in a left-handed frame)."""
    x0, y0, z0 = lo
    x1, y1, z1 = hi
    c = np.array([[x0, y0, z0], [x1, y0, z0], [x1, y1, z0], [x0, y1, z0], [x0, y0, z1], [x1, y0, z1], [x1, y1, z1], [x0, y1, z1]], f32)
    faces = [(0, 3, 2, 1), (4, 5, 6, 7), (0, 1, 5, 4), (3, 7, 6, 2), (0, 4, 7, 3), (1, 2, 6, 5)]
    return np.stack([c[list(f)] for f in faces])  # [6, 4, 3]


def synthetic_city(seed: int = 7, n_blocks: int = 10, quads_per_batch: int = 96) -> PreparedScene:
    """Small deterministic 'city of boxes + ground tiles + a few triangles' scene used when the
    reference's scene data is not available (and for the committed golden fixtures)."""
    rng = np.random.default_rng(seed)
    quads = []
    for i in range(n_blocks):
        for j in range(n_blocks):
            cx, cz = 12.0 * i + rng.uniform(-2, 2), 12.0 * j + rng.uniform(-2, 2)
            w, d, h = rng.uniform(2, 5), rng.uniform(2, 5), rng.uniform(3, 25)
            quads.append(_box_quads((cx - w, 0.0, cz - d), (cx + w, h, cz + d)))
            # ground tile, facing up
            g = np.array([[cx - 6, 0, cz - 6], [cx - 6, 0, cz + 6], [cx + 6, 0, cz + 6], [cx + 6, 0, cz - 6]], f32)
            quads.append(g[None])
            # a lone triangle stored as a degenerate quad (i0,i2,i1,i0), QuadDecomposition.cpp:405-408
            t = np.array([[cx, h, cz], [cx + 1.5, h + 2.5, cz], [cx, h + 2.5, cz + 1.5]], f32)
            quads.append(np.stack([t[0], t[2], t[1], t[0]])[None])
    q = np.concatenate(quads).astype(f32)            # [n, 4, 3]
    # spatially coherent batches: sort by a coarse grid key of the quad centre
    ctr = q.mean(axis=1)
    key = np.floor(ctr[:, 0] / 30.0) * 1000 + np.floor(ctr[:, 2] / 30.0)
    q = q[np.argsort(key, kind="stable")]
    verts = np.concatenate([q, np.ones(q.shape[:2] + (1,), f32)], axis=2).reshape(-1, 4)
    ref_min = np.append(verts[:, :3].min(axis=0), f32(1.0)).astype(f32)
    ref_max = np.append(verts[:, :3].max(axis=0), f32(1.0)).astype(f32)
    batches = []
    n_quads = q.shape[0]
    for s in range(0, n_quads, quads_per_batch):
        b = verts[4 * s: 4 * min(s + quads_per_batch, n_quads)]
        pad = (-(b.shape[0] // 4)) % 8                  # Main.cpp:91-94 pads with a collapsed quad
        if pad:
            b = np.concatenate([b, np.repeat(b[:1], 4 * pad, axis=0)])
        batches.append(np.ascontiguousarray(b))
    camera = dict(pos=(-20.0, 6.0, -14.0), dir=(0.70, -0.05, 0.71), up=(0.0, 1.0, 0.0), fov=0.9)
    return PreparedScene("city", batches, ref_min, ref_max, camera)


def synthetic_soup(n_quads: int, seed: int = 0x5EED, batch: int = 512, cube: float = 200.0) -> PreparedScene:
    """Config 4 shape: planar convex quads in batches of 512, batch centres uniform in a
    cube, quad centres Gaussian around the batch centre, random orientation, edge lengths
    U(0.5, 3); the camera sits at the cube centre so many quads straddle the near plane."""
    rng = np.random.default_rng(seed)
    n_batches = (n_quads + batch - 1) // batch
    batches = []
    for _ in range(n_batches):
        bc = rng.uniform(-cube / 2, cube / 2, 3)
        c = bc + rng.normal(0.0, 6.0, (batch, 3))
        u = rng.normal(size=(batch, 3))
        u /= np.linalg.norm(u, axis=1, keepdims=True)
        t = rng.normal(size=(batch, 3))
        v = np.cross(u, t)
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        a = rng.uniform(0.5, 3.0, (batch, 1)) * 0.5
        b = rng.uniform(0.5, 3.0, (batch, 1)) * 0.5
        q = np.stack([c - a * u - b * v, c + a * u - b * v, c + a * u + b * v, c - a * u + b * v], axis=1)
        verts = np.concatenate([q, np.ones((batch, 4, 1))], axis=2).reshape(-1, 4).astype(f32)
        batches.append(np.ascontiguousarray(verts))
    allv = np.concatenate(batches)
    ref_min = np.append(allv[:, :3].min(axis=0), f32(1.0)).astype(f32)
    ref_max = np.append(allv[:, :3].max(axis=0), f32(1.0)).astype(f32)
    camera = dict(pos=(0.0, 0.0, 0.0), dir=(0.0, 0.0, 1.0), up=(0.0, 1.0, 0.0), fov=1.0)
    return PreparedScene(f"soup{n_quads}", batches, ref_min, ref_max, camera)


# ------------------------------------------------------------------------------------------------
# camera sets
def camera_path(scene: PreparedScene, n: int, width: int, height: int):
    """Config 3: deterministic closed orbit through the scene -> (mvps [n,16] f32, positions [n,3] f32)."""
    lo, hi = scene.ref_min[:3].astype(np.float64), scene.ref_max[:3].astype(np.float64)
    up = np.asarray(scene.camera.get("up", (0, 1, 0)), np.float64)
    upi = int(np.argmax(np.abs(up)))
    a, b = [i for i in range(3) if i != upi]
    key = scene.name.lower()
    if key == "castle":     # orbit around the keep at walking height (scene extents: SURVEY 8)
        ctr = {a: 92.0, b: -4.0}
        rad = {a: 62.0, b: 52.0}
        base_h, amp_h = 4.0, 2.5
    else:
        ctr = {a: 0.5 * (lo[a] + hi[a]), b: 0.5 * (lo[b] + hi[b])}
        rad = {a: 0.33 * (hi[a] - lo[a]), b: 0.33 * (hi[b] - lo[b])}
        base_h, amp_h = lo[upi] + 0.25 * (hi[upi] - lo[upi]), 0.05 * (hi[upi] - lo[upi])
    fov = scene.camera.get("fov", 0.8)
    mvps = np.zeros((n, 16), f32)
    poss = np.zeros((n, 3), f32)
    for i in range(n):
        th = 2.0 * np.pi * i / n
        p = np.zeros(3)
        p[a] = ctr[a] + rad[a] * np.cos(th)
        p[b] = ctr[b] + rad[b] * np.sin(th)
        p[upi] = base_h + amp_h * np.sin(3.0 * th)
        tangent = np.zeros(3)
        tangent[a], tangent[b] = -rad[a] * np.sin(th), rad[b] * np.cos(th)
        inward = np.zeros(3)
        inward[a], inward[b] = ctr[a] - p[a], ctr[b] - p[b]
        d = tangent / np.linalg.norm(tangent) * 0.55 + inward / np.linalg.norm(inward) * (0.45 + 0.35 * np.sin(2.0 * th))
        d[upi] = 0.08 * np.sin(5.0 * th)
        d /= np.linalg.norm(d)
        poss[i] = p.astype(f32)
        mvps[i] = cam.view_projection(poss[i], d.astype(f32), up.astype(f32), fov, width, height)
    return mvps, poss


def probe_views(scene: PreparedScene, n: int, width: int, height: int, seed: int = 1234):
    """Config 5: visibility probes -- positions jittered on a 3-D grid inside the scene AABB, the six
    axis look directions cycled per position -> (mvps [n,16], positions [n,3])."""
    rng = np.random.default_rng(seed)
    lo, hi = scene.ref_min[:3].astype(np.float64), scene.ref_max[:3].astype(np.float64)
    up = np.asarray(scene.camera.get("up", (0, 1, 0)), np.float64)
    upi = int(np.argmax(np.abs(up)))
    n_pos = (n + 5) // 6
    g = int(np.ceil(n_pos ** (1.0 / 3.0)))
    dirs = [(1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)]
    fov = scene.camera.get("fov", 0.8)
    mvps = np.zeros((n, 16), f32)
    poss = np.zeros((n, 3), f32)
    k = 0
    for ip in range(n_pos):
        cell = np.array([ip % g, (ip // g) % g, ip // (g * g)], np.float64)
        p = lo + (cell + rng.uniform(0.2, 0.8, 3)) / g * (hi - lo)
        for d in dirs:
            if k >= n:
                break
            d = np.asarray(d, np.float64)
            # keep the up vector off the view axis
            u = up if abs(d[upi]) < 0.5 else np.roll(up, 1)
            poss[k] = p.astype(f32)
            mvps[k] = cam.view_projection(poss[k], d.astype(f32), u.astype(f32), fov, width, height)
            k += 1
    return mvps, poss


def orders_for(centers: np.ndarray, positions: np.ndarray) -> np.ndarray:
    """Front-to-back occluder order per view (Main.cpp:185-190) -> uint32 [nViews, nOccluders]."""
    return np.stack([cam.front_to_back_order(centers, p) for p in positions]).astype(np.uint32)
