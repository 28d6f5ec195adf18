"""Python host side over the C ABI (include/orz.h) of librasterizer_b200.so.

Mirrors the reference's interface for the hot path -- `Occluder.bake`, `Rasterizer`
(`setModelViewProjection`, `clear`, `rasterize`, `queryVisibility`, `query2D`, `readBackDepth`;
Rasterizer.h:13-26, Occluder.h:9) -- plus the view-batch entry point that runs the frame loop of
Main.cpp:181-206 for many independent views on the GPU.

There is no CPU path: importing works anywhere, but every compute call needs the CUDA library
and a device, and raises otherwise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ORZ_LIB") or os.path.join(HERE, "librasterizer_b200.so")  # ORZ_LIB: profiling builds (tools/)

BATCH_NO_GATE = 1
BATCH_FORCE_CLIPPED = 2
BATCH_TARGETS_ON_DEVICE = 4
BATCH_WIDE = 8

_lib = None


class OrzError(RuntimeError):
    pass


class PrimRecord(C.Structure):
    _fields_ = [("mode", C.c_uint32), ("minX", C.c_int32), ("minY", C.c_int32), ("rangeX", C.c_int32), ("rangeY", C.c_int32),
                ("maxZ", C.c_uint32), ("dzdx", C.c_float), ("dzdy", C.c_float), ("plane0", C.c_float),
                ("nx", C.c_float * 4), ("ny", C.c_float * 4), ("off", C.c_float * 4), ("slope", C.c_uint32 * 4)]


class MeshSceneInfo(C.Structure):
    _fields_ = [("nOccluders", C.c_uint32), ("nQuads", C.c_uint32), ("refMin", C.c_float * 4), ("refMax", C.c_float * 4)]


class ViewBatch(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("nViews", C.c_uint32), ("flags", C.c_uint32),
                ("mvps", C.c_void_p), ("orders", C.c_void_p), ("camPos", C.c_void_p),
                ("visBits", C.c_void_p), ("clipBits", C.c_void_p), ("gate", C.c_void_p),
                ("depth", C.c_void_p), ("hiz", C.c_void_p), ("quadsSubmitted", C.c_void_p)]


EXPORTS = {
    "orz_scene_occludee_count": (C.c_uint32, [C.c_void_p]),
    "orz_scene_from_mesh_files": (C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(C.c_void_p), C.c_void_p]),
    "orz_scene_save": (C.c_int, [C.c_void_p, C.c_char_p]),
    "orz_scene_load": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p)]),
    "orz_scene_from_mesh": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_int,
                                      C.POINTER(C.c_void_p), C.c_void_p]),
    "orz_scene_get_occluders": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint32), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "orz_generate_batches_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p,
                                              C.c_uint32, C.POINTER(C.c_uint32)]),
    "orz_quad_decompose": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_size_t)]),
    "orz_generate_batches": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32,
                                       C.POINTER(C.c_uint32)]),
    "orz_last_error": (C.c_char_p, []),
    "orz_version": (C.c_int, []),
    "orz_context_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "orz_context_destroy": (None, [C.c_void_p]),
    "orz_context_synchronize": (C.c_int, [C.c_void_p]),
    "orz_context_stream": (C.c_void_p, [C.c_void_p]),
    "orz_context_set_rcp_table": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "orz_context_get_rcp_table": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]),
    "orz_context_get_lut": (C.c_int, [C.c_void_p, C.c_void_p]),
    "orz_context_launch_count": (C.c_uint64, [C.c_void_p]),
    "orz_context_set_group_warps": (C.c_int, [C.c_void_p, C.c_int]),
    "orz_context_set_arena_bytes": (C.c_int, [C.c_void_p, C.c_size_t]),
    "orz_context_set_cluster_views": (C.c_int, [C.c_void_p, C.c_int]),
    "orz_context_set_cluster_size": (C.c_int, [C.c_void_p, C.c_int]),
    "orz_context_set_tile_height": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "orz_context_set_traversal": (C.c_int, [C.c_void_p, C.c_int]),
    "orz_edge_mask_table": (C.c_int, [C.c_void_p]),
    "orz_probe_host_rcp": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "orz_bake": (C.c_uint32, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "orz_set_rsqrt_table": (C.c_int, [C.c_void_p, C.c_int]),
    "orz_occluder_create": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "orz_occluder_destroy": (None, [C.c_void_p]),
    "orz_rasterizer_create": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p)]),
    "orz_rasterizer_destroy": (None, [C.c_void_p]),
    "orz_rasterizer_set_mvp": (C.c_int, [C.c_void_p, C.c_void_p]),
    "orz_rasterizer_clear": (C.c_int, [C.c_void_p]),
    "orz_rasterizer_rasterize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "orz_rasterizer_query_visibility": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "orz_rasterizer_query2d": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_int)]),
    "orz_rasterizer_query_boxes": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]),
    "orz_rasterizer_readback_depth": (C.c_int, [C.c_void_p, C.c_void_p]),
    "orz_rasterizer_download": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "orz_rasterizer_debug_setup": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "orz_scene_create": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "orz_scene_bake": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "orz_scene_set_occludees": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32]),
    "orz_scene_destroy": (None, [C.c_void_p]),
    "orz_render_views": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(ViewBatch)]),
    "orz_render_views_device": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(ViewBatch)]),
    "orz_context_device": (C.c_int, [C.c_void_p]),
    "orz_comm_get_unique_id": (C.c_int, [C.c_void_p]),
    "orz_comm_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "orz_comm_create_all": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_void_p)]),
    "orz_comm_destroy": (None, [C.c_void_p]),
    "orz_comm_rank": (C.c_int, [C.c_void_p]),
    "orz_comm_size": (C.c_int, [C.c_void_p]),
    "orz_comm_group_begin": (C.c_int, []),
    "orz_comm_group_end": (C.c_int, []),
    "orz_gather_bits": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "orz_gather_bits_overlapped": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "orz_comm_join": (C.c_int, [C.c_void_p]),
    "orz_comm_join_older": (C.c_int, [C.c_void_p]),
    "orz_comm_synchronize": (C.c_int, [C.c_void_p]),
}
COMM_ID_BYTES = 128


def lib():
    """The CUDA library; raises when it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OrzError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(make -C rasterizer_b200/csrc); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def _check(code: int):
    if code != 0:
        raise OrzError(f"orz error {code}: {lib().orz_last_error().decode()}")


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a, shape=None):
    a = np.ascontiguousarray(a, np.float32)
    return a if shape is None else a.reshape(shape)


def bake(vertices, ref_min, ref_max):
    """Occluder::bake (Occluder.cpp:7-181) on the host -> (packets uint32 in the reference layout,
    center, boundsMin, boundsMax)."""
    v = _f32(vertices).reshape(-1, 4)
    if v.shape[0] % 32 != 0:
        raise ValueError("bake needs a multiple of 8 quads (32 vertices)")
    mn, mx = _f32(ref_min), _f32(ref_max)
    packets = np.zeros(v.shape[0], np.uint32)
    c, bmin, bmax = np.zeros(4, np.float32), np.zeros(4, np.float32), np.zeros(4, np.float32)
    n = lib().orz_bake(_p(v), v.shape[0], _p(mn), _p(mx), _p(packets), _p(c), _p(bmin), _p(bmax))
    if n * 8 != v.shape[0]:
        raise OrzError("orz_bake failed")
    return packets, c, bmin, bmax


def quad_decompose(indices, vertices) -> np.ndarray:
    """QuadDecomposition::decompose (QuadDecomposition.cpp:346-445): triangle list -> quad list (4 indices per quad)."""
    idx = np.ascontiguousarray(indices, np.uint32).reshape(-1)
    v = _f32(vertices).reshape(-1, 4)
    out = np.zeros(4 * (idx.size // 3), np.uint32)
    n = C.c_size_t()
    _check(lib().orz_quad_decompose(_p(idx), idx.size, _p(v), v.shape[0], _p(out), C.byref(n)))
    return out[: n.value].copy()


def generate_batches(aabbs, target_size: int = 512, split_granularity: int = 8):
    """SurfaceAreaHeuristic::generateBatches (SurfaceAreaHeuristic.cpp:96-104) -> list of index arrays."""
    b = _f32(aabbs).reshape(-1, 8)
    order = np.zeros(b.shape[0], np.uint32)
    cap = b.shape[0] // max(split_granularity, 1) + 2
    sizes = np.zeros(cap, np.uint32)
    n = C.c_uint32()
    _check(lib().orz_generate_batches(_p(b), b.shape[0], target_size, split_granularity, _p(order), _p(sizes), cap, C.byref(n)))
    return np.split(order, np.cumsum(sizes[: n.value])[:-1])


def edge_mask_table() -> np.ndarray:
    """The 64x64 edge-mask table (Rasterizer.cpp:547-604), built on the host."""
    t = np.zeros(4096, np.int64)
    _check(lib().orz_edge_mask_table(_p(t)))
    return t


def probe_host_rcp():
    """-> (table uint32[2^bits], bits, exact) for the CPU this runs on."""
    bits, exact = C.c_int(), C.c_int()
    _check(lib().orz_probe_host_rcp(None, C.byref(bits), C.byref(exact)))
    t = np.zeros(1 << bits.value, np.uint32)
    _check(lib().orz_probe_host_rcp(_p(t), C.byref(bits), C.byref(exact)))
    return t, bits.value, bool(exact.value)


def set_rsqrt_table(table):
    if table is None:
        _check(lib().orz_set_rsqrt_table(None, 0))
    else:
        t = np.ascontiguousarray(table, np.uint32)
        _check(lib().orz_set_rsqrt_table(_p(t), int(np.log2(t.size)) - 1))


class Context:
    """One per (host thread, GPU): CUDA stream, edge-mask table, rcpps model."""

    def __init__(self, device: int = 0):
        h = C.c_void_p()
        _check(lib().orz_context_create(device, C.byref(h)))
        self.h = h
        self.device = device

    def close(self):
        if self.h:
            lib().orz_context_destroy(self.h)
            self.h = None

    def synchronize(self):
        _check(lib().orz_context_synchronize(self.h))

    @property
    def stream(self) -> int:
        return int(lib().orz_context_stream(self.h) or 0)

    @property
    def launch_count(self) -> int:
        return int(lib().orz_context_launch_count(self.h))

    def generate_batches(self, aabbs, target_size: int = 512, split_granularity: int = 8):
        """SurfaceAreaHeuristic::generateBatches on the GPU -> list of index arrays (same as api.generate_batches)."""
        b = _f32(aabbs).reshape(-1, 8)
        order = np.zeros(b.shape[0], np.uint32)
        cap = b.shape[0] // max(split_granularity, 1) + 2
        sizes = np.zeros(cap, np.uint32)
        n = C.c_uint32()
        _check(lib().orz_generate_batches_device(self.h, _p(b), b.shape[0], target_size, split_granularity, _p(order), _p(sizes), cap,
                                                 C.byref(n)))
        return np.split(order, np.cumsum(sizes[: n.value])[:-1])

    def set_group_warps(self, warps: int):
        _check(lib().orz_context_set_group_warps(self.h, warps))

    def set_traversal(self, mapping: int):
        _check(lib().orz_context_set_traversal(self.h, mapping))

    def set_cluster_views(self, max_views: int):
        _check(lib().orz_context_set_cluster_views(self.h, max_views))

    def set_cluster_size(self, ctas: int):
        _check(lib().orz_context_set_cluster_size(self.h, ctas))

    def set_tile_height(self, cluster: int = 0, per_call: int = 1):
        _check(lib().orz_context_set_tile_height(self.h, cluster, per_call))

    def set_arena_bytes(self, nbytes: int):
        _check(lib().orz_context_set_arena_bytes(self.h, nbytes))

    def set_rcp_table(self, table):
        t = np.ascontiguousarray(table, np.uint32)
        _check(lib().orz_context_set_rcp_table(self.h, _p(t), int(np.log2(t.size))))

    def rcp_table(self) -> np.ndarray:
        bits = C.c_int()
        _check(lib().orz_context_get_rcp_table(self.h, None, C.byref(bits)))
        t = np.zeros(1 << bits.value, np.uint32)
        _check(lib().orz_context_get_rcp_table(self.h, _p(t), C.byref(bits)))
        return t

    def lut(self) -> np.ndarray:
        t = np.zeros(4096, np.int64)
        _check(lib().orz_context_get_lut(self.h, _p(t)))
        return t


class Occluder:
    """Baked quad batch (Occluder.h:7-21): host copy in the reference layout + device copy."""

    def __init__(self, ctx: Context, packets, ref_min, ref_max, center=None, bounds_min=None, bounds_max=None):
        self.ctx = ctx
        self.packets = np.ascontiguousarray(packets, np.uint32)
        self.m_packetCount = self.packets.size // 8
        self.m_refMin, self.m_refMax = _f32(ref_min), _f32(ref_max)
        self.m_center, self.m_boundsMin, self.m_boundsMax = center, bounds_min, bounds_max
        h = C.c_void_p()
        _check(lib().orz_occluder_create(ctx.h, _p(self.packets), self.m_packetCount, _p(self.m_refMin), _p(self.m_refMax), C.byref(h)))
        self.h = h

    @classmethod
    def bake(cls, ctx: Context, vertices, ref_min, ref_max) -> "Occluder":
        packets, c, bmin, bmax = bake(vertices, ref_min, ref_max)
        return cls(ctx, packets, ref_min, ref_max, c, bmin, bmax)

    def close(self):
        if self.h:
            lib().orz_occluder_destroy(self.h)
            self.h = None


DEFAULT_CLUSTER_VIEWS = 16384  # orz_context's default for set_cluster_views


class Rasterizer:
    """Per-call API of the reference (Rasterizer.h:10-61) on one view's device buffers."""

    def __init__(self, ctx: Context, width: int, height: int):
        self.ctx, self.width, self.height = ctx, width, height
        self.blocks = (width // 8) * (height // 8)
        h = C.c_void_p()
        _check(lib().orz_rasterizer_create(ctx.h, width, height, C.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            lib().orz_rasterizer_destroy(self.h)
            self.h = None

    def setModelViewProjection(self, matrix):
        m = _f32(matrix).reshape(16)
        _check(lib().orz_rasterizer_set_mvp(self.h, _p(m)))

    def clear(self):
        _check(lib().orz_rasterizer_clear(self.h))

    def rasterize(self, occluder: Occluder, possiblyNearClipped: bool):
        _check(lib().orz_rasterizer_rasterize(self.h, occluder.h, int(possiblyNearClipped)))

    def queryVisibility(self, bounds_min, bounds_max):
        """-> (visible, needsClipping)"""
        a, b = _f32(bounds_min), _f32(bounds_max)
        vis, clip = C.c_int(0), C.c_int(0)
        _check(lib().orz_rasterizer_query_visibility(self.h, _p(a), _p(b), C.byref(vis), C.byref(clip)))
        return bool(vis.value), bool(clip.value)

    def query2D(self, min_x, max_x, min_y, max_y, max_z) -> bool:
        vis = C.c_int(0)
        _check(lib().orz_rasterizer_query2d(self.h, min_x, max_x, min_y, max_y, max_z, C.byref(vis)))
        return bool(vis.value)

    def query_boxes(self, boxes) -> np.ndarray:
        b = _f32(boxes).reshape(-1, 8)
        out = np.zeros(b.shape[0], np.uint8)
        _check(lib().orz_rasterizer_query_boxes(self.h, _p(b), b.shape[0], _p(out)))
        return out

    def readBackDepth(self) -> np.ndarray:
        out = np.zeros(self.width * self.height * 4, np.uint8)
        _check(lib().orz_rasterizer_readback_depth(self.h, _p(out)))
        return out

    def download(self):
        depth, hiz = np.zeros(self.blocks * 64, np.uint16), np.zeros(self.blocks, np.uint16)
        _check(lib().orz_rasterizer_download(self.h, _p(depth), _p(hiz)))
        return depth, hiz

    def debug_setup(self, occluder: Occluder, clipped: bool):
        n = occluder.m_packetCount * 2
        out = (PrimRecord * n)()
        _check(lib().orz_rasterizer_debug_setup(self.h, occluder.h, int(clipped), out))
        return out


class Scene:
    """All baked batches of a scene + its occludee boxes, resident in HBM."""

    def __init__(self, ctx: Context, packed_list, ref_min, ref_max, bounds_min, bounds_max, centers, boxes=None):
        self.ctx = ctx
        n = len(packed_list)
        self.n_occluders = n
        self.packed_list = [np.ascontiguousarray(p, np.uint32) for p in packed_list]
        packets = np.ascontiguousarray(np.concatenate(self.packed_list))
        counts = np.array([p.size // 8 for p in self.packed_list], np.uint32)
        self.quads_per_occluder = counts * 2
        rmn = np.ascontiguousarray(np.broadcast_to(_f32(ref_min).reshape(-1, 4), (n, 4)))
        rmx = np.ascontiguousarray(np.broadcast_to(_f32(ref_max).reshape(-1, 4), (n, 4)))
        self.ref_min, self.ref_max = rmn, rmx
        self.bounds_min, self.bounds_max, self.centers = _f32(bounds_min, (n, 4)), _f32(bounds_max, (n, 4)), _f32(centers, (n, 4))
        h = C.c_void_p()
        _check(lib().orz_scene_create(ctx.h, _p(packets), _p(counts), n, _p(rmn), _p(rmx), _p(self.bounds_min), _p(self.bounds_max),
                                      _p(self.centers), C.byref(h)))
        self.h = h
        self.n_boxes = 0
        if boxes is not None:
            self.set_occludees(boxes)

    @classmethod
    def from_prepared(cls, ctx: Context, prepared, boxes="quads") -> "Scene":
        """Bake every batch of a workloads.PreparedScene on the host and upload it."""
        baked = [bake(b, prepared.ref_min, prepared.ref_max) for b in prepared.batches]
        bx = prepared.quad_boxes() if isinstance(boxes, str) and boxes == "quads" else boxes
        return cls(ctx, [b[0] for b in baked], prepared.ref_min, prepared.ref_max, np.stack([b[2] for b in baked]),
                   np.stack([b[3] for b in baked]), np.stack([b[1] for b in baked]), bx)

    @classmethod
    def bake_on_device(cls, ctx: Context, batches, ref_min, ref_max, boxes=None) -> "Scene":
        """Occluder::bake for every batch on the GPU (orz_scene_bake); the host copies of the baked data
        (packets in the reference layout, bounds, centres) come back for the caller's own use."""
        self = cls.__new__(cls)
        self.ctx = ctx
        vs = [_f32(b).reshape(-1, 4) for b in batches]
        n = len(vs)
        counts = np.array([v.shape[0] for v in vs], np.uint32)
        verts = np.ascontiguousarray(np.concatenate(vs))
        rmn, rmx = _f32(ref_min).reshape(4), _f32(ref_max).reshape(4)
        packets = np.zeros(int(counts.sum()), np.uint32)
        cen, bmn, bmx = (np.zeros((n, 4), np.float32) for _ in range(3))
        h = C.c_void_p()
        _check(lib().orz_scene_bake(ctx.h, _p(verts), _p(counts), n, _p(rmn), _p(rmx), _p(packets), _p(cen), _p(bmn), _p(bmx), C.byref(h)))
        self.h = h
        self.n_occluders = n
        ofs = np.concatenate([[0], np.cumsum(counts.astype(np.int64))]).astype(np.int64)
        self.packed_list = [packets[int(ofs[i]):int(ofs[i + 1])] for i in range(n)]
        self.quads_per_occluder = counts // 4
        self.ref_min = np.ascontiguousarray(np.broadcast_to(rmn, (n, 4)))
        self.ref_max = np.ascontiguousarray(np.broadcast_to(rmx, (n, 4)))
        self.bounds_min, self.bounds_max, self.centers = bmn, bmx, cen
        self.n_boxes = 0
        if boxes is not None:
            self.set_occludees(boxes)
        return self

    @classmethod
    def from_mesh(cls, ctx: Context, indices, vertices, target_size: int = 512, split_granularity: int = 8, occludees_from_quads: bool = True) -> "Scene":
        """Main.cpp:86-128 in one call (orz_scene_from_mesh): triangle mesh in, baked scene resident in HBM out."""
        self = cls.__new__(cls)
        self.ctx = ctx
        idx = np.ascontiguousarray(indices, np.uint32).reshape(-1)
        v = _f32(vertices).reshape(-1, 4)
        h, info = C.c_void_p(), MeshSceneInfo()
        _check(lib().orz_scene_from_mesh(ctx.h, _p(idx), idx.size, _p(v), v.shape[0], target_size, split_granularity, int(occludees_from_quads),
                                         C.byref(h), C.byref(info)))
        self.h = h
        n = self.n_occluders = int(info.nOccluders)
        self.n_quads = int(info.nQuads)
        self.centers, self.bounds_min, self.bounds_max = (np.zeros((n, 4), np.float32) for _ in range(3))
        counts = np.zeros(n, np.uint32)
        _check(lib().orz_scene_get_occluders(h, None, _p(self.centers), _p(self.bounds_min), _p(self.bounds_max), _p(counts)))
        self.quads_per_occluder = counts
        self.ref_min = np.ascontiguousarray(np.broadcast_to(np.array(info.refMin, np.float32), (n, 4)))
        self.ref_max = np.ascontiguousarray(np.broadcast_to(np.array(info.refMax, np.float32), (n, 4)))
        self.packed_list = None  # the baked words exist in HBM only
        self.n_boxes = self.n_quads if occludees_from_quads else 0
        return self

    def _adopt(self, ctx, h):
        """Fill the host-side fields of a scene that was created inside the library."""
        self.ctx, self.h = ctx, h
        n = C.c_uint32()
        _check(lib().orz_scene_get_occluders(h, C.byref(n), None, None, None, None))
        n = self.n_occluders = int(n.value)
        self.centers, self.bounds_min, self.bounds_max = (np.zeros((n, 4), np.float32) for _ in range(3))
        counts = np.zeros(n, np.uint32)
        _check(lib().orz_scene_get_occluders(h, None, _p(self.centers), _p(self.bounds_min), _p(self.bounds_max), _p(counts)))
        self.quads_per_occluder = counts
        self.n_quads = int(counts.sum())
        self.packed_list = None  # the baked words exist in HBM only
        self.n_boxes = int(lib().orz_scene_occludee_count(h))
        return self

    @classmethod
    def from_mesh_files(cls, ctx: Context, index_path: str, vertex_path: str, target_size: int = 512, split_granularity: int = 8,
                        occludees_from_quads: bool = True) -> "Scene":
        """The reference's raw scene files (Main.cpp:56-84) -> baked scene in HBM (orz_scene_from_mesh_files)."""
        h = C.c_void_p()
        _check(lib().orz_scene_from_mesh_files(ctx.h, index_path.encode(), vertex_path.encode(), target_size, split_granularity,
                                               int(occludees_from_quads), C.byref(h), None))
        return cls.__new__(cls)._adopt(ctx, h)

    def save(self, path: str):
        """Cached baked scene file (orz_scene_save)."""
        _check(lib().orz_scene_save(self.h, path.encode()))

    @classmethod
    def load(cls, ctx: Context, path: str) -> "Scene":
        """orz_scene_load: a file written by `save`, straight to HBM."""
        h = C.c_void_p()
        _check(lib().orz_scene_load(ctx.h, path.encode(), C.byref(h)))
        return cls.__new__(cls)._adopt(ctx, h)

    def set_occludees(self, boxes):
        b = _f32(boxes).reshape(-1, 8)
        _check(lib().orz_scene_set_occludees(self.h, _p(b), b.shape[0]))
        self.n_boxes = b.shape[0]

    def close(self):
        if self.h:
            lib().orz_scene_destroy(self.h)
            self.h = None

    def render_views(self, width, height, mvps, orders=None, cam_pos=None, flags=0, want=("vis",)):
        """Host-buffer entry point (orz_render_views).  `want` picks outputs out of
        vis, clip, gate, depth, hiz, quads.  Returns a dict of numpy arrays."""
        mvps = _f32(mvps).reshape(-1, 16)
        nv = mvps.shape[0]
        blocks = (width // 8) * (height // 8)
        words = (self.n_boxes + 31) // 32
        out = {}
        b = ViewBatch()
        b.width, b.height, b.nViews, b.flags = width, height, nv, flags
        keep = [mvps]
        b.mvps = mvps.ctypes.data
        if orders is not None:
            o = np.ascontiguousarray(orders, np.uint32).reshape(nv, self.n_occluders)
            keep.append(o)
            b.orders = o.ctypes.data
        else:
            cp = _f32(cam_pos).reshape(nv, 3)
            keep.append(cp)
            b.camPos = cp.ctypes.data
        if "vis" in want:
            out["vis"] = np.zeros((nv, words), np.uint32); b.visBits = out["vis"].ctypes.data
        if "clip" in want:
            out["clip"] = np.zeros((nv, words), np.uint32); b.clipBits = out["clip"].ctypes.data
        if "gate" in want:
            out["gate"] = np.zeros((nv, self.n_occluders), np.uint8); b.gate = out["gate"].ctypes.data
        if "depth" in want or "hiz" in want:
            out["depth"] = np.zeros((nv, blocks * 64), np.uint16); b.depth = out["depth"].ctypes.data
            out["hiz"] = np.zeros((nv, blocks), np.uint16); b.hiz = out["hiz"].ctypes.data
        if "quads" in want:
            out["quads"] = np.zeros(nv, np.uint32); b.quadsSubmitted = out["quads"].ctypes.data
        _check(lib().orz_render_views(self.ctx.h, self.h, C.byref(b)))
        return out

    def render_views_raw(self, batch: ViewBatch, device: bool):
        fn = lib().orz_render_views_device if device else lib().orz_render_views
        _check(fn(self.ctx.h, self.h, C.byref(batch)))


class Comm:
    """NCCL communicator of the view-batch path behind the C ABI (orz_comm_*): one rank per (process, GPU).
    `unique_id()` on rank 0, the 128 bytes reach the other ranks by the caller's means, then every rank constructs."""

    def __init__(self, ctx: Context, n_ranks: int, rank: int, unique_id: bytes):
        assert len(unique_id) == COMM_ID_BYTES
        self.ctx = ctx
        buf = C.create_string_buffer(bytes(unique_id), COMM_ID_BYTES)
        h = C.c_void_p()
        _check(lib().orz_comm_create(ctx.h, n_ranks, rank, buf, C.byref(h)))
        self.h = h
        self.n_ranks, self.rank = n_ranks, rank

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(COMM_ID_BYTES)
        _check(lib().orz_comm_get_unique_id(buf))
        return buf.raw

    def gather_bits(self, local_ptr: int, words_per_rank: int, all_ptr: int, overlapped: bool = False):
        """Device pointers: local [words_per_rank] u32 -> all [n_ranks x words_per_rank] u32 (orz_gather_bits[_overlapped])."""
        fn = lib().orz_gather_bits_overlapped if overlapped else lib().orz_gather_bits
        _check(fn(self.h, local_ptr, words_per_rank, all_ptr))

    def join(self):
        _check(lib().orz_comm_join(self.h))

    def join_older(self):
        _check(lib().orz_comm_join_older(self.h))

    def synchronize(self):
        _check(lib().orz_comm_synchronize(self.h))

    def close(self):
        if self.h:
            lib().orz_comm_destroy(self.h)
            self.h = None


def unpack_bits(words: np.ndarray, n: int) -> np.ndarray:
    """[nViews, words] uint32 -> [nViews, n] bool"""
    w = np.ascontiguousarray(words, np.uint32)
    bits = np.unpackbits(w.view(np.uint8).reshape(w.shape[0], -1), axis=1, bitorder="little")
    return bits[:, :n].astype(bool)
