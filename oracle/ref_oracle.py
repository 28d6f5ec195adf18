"""TEST INFRASTRUCTURE ONLY: ctypes view of oracle/_ref/libref_oracle.so, i.e. the UNMODIFIED
reference sources (Rasterizer.cpp, Occluder.cpp, QuadDecomposition.cpp, SurfaceAreaHeuristic.cpp)
compiled by oracle/Makefile plus the headless harness oracle/ref_harness.cpp.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libref_oracle.so")
SCENE_DIR = os.path.join(HERE, "_ref", "scenes")

_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def scene_available(name: str) -> bool:
    return os.path.exists(os.path.join(SCENE_DIR, name, "IndexBuffer.bin"))


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        vp, u32, f_p = C.c_void_p, C.c_uint32, C.POINTER(C.c_float)
        sig = {
            "ref_scene_load": (vp, [C.c_char_p, C.c_char_p]),
            "ref_scene_from_batches": (vp, [vp, vp, u32, vp, vp]),
            "ref_scene_free": (None, [vp]),
            "ref_quad_decompose": (C.c_size_t, [vp, C.c_size_t, vp, C.c_size_t, vp]),
            "ref_generate_batches": (u32, [vp, u32, u32, u32, vp, vp]),
            "ref_scene_num_occluders": (u32, [vp]),
            "ref_scene_num_boxes": (u32, [vp]),
            "ref_scene_boxes": (vp, [vp]),
            "ref_scene_ref_aabb": (None, [vp, vp, vp]),
            "ref_scene_batch_quads": (u32, [vp, u32]),
            "ref_scene_batch_vertices": (None, [vp, u32, vp]),
            "ref_scene_occluder_meta": (None, [vp, u32, vp, vp, vp, vp]),
            "ref_scene_occluder_packets": (vp, [vp, u32]),
            "ref_rast_create": (vp, [u32, u32]),
            "ref_rast_free": (None, [vp]),
            "ref_rast_set_mvp": (None, [vp, vp]),
            "ref_rast_clear": (None, [vp, C.c_int]),
            "ref_rast_rasterize": (None, [vp, vp, u32, C.c_int]),
            "ref_rast_query": (C.c_int, [vp, vp, vp]),
            "ref_rast_query2d": (C.c_int, [vp, u32, u32, u32, u32, u32]),
            "ref_rast_query_boxes": (None, [vp, vp, u32, vp]),
            "ref_rast_readback": (None, [vp, vp]),
            "ref_rast_get_hiz": (None, [vp, vp]),
            "ref_rast_get_depth": (None, [vp, vp, C.c_int]),
            "ref_rast_get_lut": (None, [vp, vp]),
            "ref_rast_get_matrices": (None, [vp, vp, vp]),
            "ref_rast_frame": (C.c_uint64, [vp, vp, vp, vp, u32, C.c_int, vp]),
            "ref_rast_submit_all": (None, [vp, vp, vp, vp, u32, C.c_int, C.c_int]),
            "ref_rcp_ps": (None, [vp, vp, C.c_size_t]),
            "ref_rsqrt_ps": (None, [vp, vp, C.c_size_t]),
            "ref_bench_views": (C.c_double, [vp, u32, u32, vp, vp, u32, u32, vp, u32, u32, u32, vp]),
            "ref_check_views": (u32, [vp, u32, u32, vp, vp, u32, u32, vp, u32, u32, u32, vp, vp, vp, vp, vp, vp, vp]),
            "ref_selftest_alignment": (C.c_int, []),
            "ref_pool_create": (vp, [u32, u32, u32]),
            "ref_pool_free": (None, [vp]),
            "ref_pool_bench": (C.c_double, [vp, vp, vp, vp, u32, u32, vp, u32, u32, vp]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        if not L.ref_selftest_alignment():
            raise RuntimeError("oracle/_ref: the harness' aligned operator new is not in effect in this process (rebuild with oracle/Makefile)")
        _lib = L
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def host_rcp(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(x)
    lib().ref_rcp_ps(_p(x), _p(out), x.size)
    return out


def host_rsqrt(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(x)
    lib().ref_rsqrt_ps(_p(x), _p(out), x.size)
    return out


def quad_decompose(indices, vertices) -> np.ndarray:
    """QuadDecomposition::decompose of the unmodified reference."""
    idx = np.ascontiguousarray(indices, np.uint32).reshape(-1)
    v = np.ascontiguousarray(vertices, np.float32).reshape(-1, 4)
    out = np.zeros(4 * (idx.size // 3), np.uint32)
    n = lib().ref_quad_decompose(_p(idx), idx.size, _p(v), v.shape[0], _p(out))
    return out[:n].copy()


def generate_batches(aabbs, target_size=512, granularity=8):
    """SurfaceAreaHeuristic::generateBatches of the unmodified reference -> list of index arrays."""
    b = np.ascontiguousarray(aabbs, np.float32).reshape(-1, 8)
    order = np.zeros(b.shape[0], np.uint32)
    sizes = np.zeros(max(b.shape[0], 2), np.uint32)
    n = lib().ref_generate_batches(_p(b), b.shape[0], target_size, granularity, _p(order), _p(sizes))
    return np.split(order, np.cumsum(sizes[:n])[:-1])


def load_mesh(name: str):
    """Raw triangle list + float4 vertices of a reference scene (Main.cpp:56-84) from oracle/_ref/scenes."""
    d = os.path.join(SCENE_DIR, name)
    return np.fromfile(os.path.join(d, "IndexBuffer.bin"), np.uint32), np.fromfile(os.path.join(d, "VertexBuffer.bin"), np.float32).reshape(-1, 4)


class RefScene:
    """Scene prepared and baked by the reference's own code (Main.cpp:56-128)."""

    def __init__(self, handle):
        if not handle:
            raise RuntimeError("reference scene could not be created")
        self.h = handle
        L = lib()
        self.n_occluders = L.ref_scene_num_occluders(self.h)
        self.ref_min = np.zeros(4, np.float32)
        self.ref_max = np.zeros(4, np.float32)
        L.ref_scene_ref_aabb(self.h, _p(self.ref_min), _p(self.ref_max))
        self.centers = np.zeros((self.n_occluders, 4), np.float32)
        self.bounds_min = np.zeros((self.n_occluders, 4), np.float32)
        self.bounds_max = np.zeros((self.n_occluders, 4), np.float32)
        self.packet_counts = np.zeros(self.n_occluders, np.uint32)
        for i in range(self.n_occluders):
            pc = C.c_uint32()
            L.ref_scene_occluder_meta(self.h, i, _p(self.centers[i]), _p(self.bounds_min[i]), _p(self.bounds_max[i]), C.byref(pc))
            self.packet_counts[i] = pc.value
        nb = L.ref_scene_num_boxes(self.h)
        if nb:
            buf = (C.c_float * (8 * nb)).from_address(L.ref_scene_boxes(self.h))
            self.boxes = np.frombuffer(buf, dtype=np.float32).reshape(nb, 8).copy()
        else:
            self.boxes = np.zeros((0, 8), np.float32)

    @classmethod
    def load(cls, name: str) -> "RefScene":
        d = os.path.join(SCENE_DIR, name)
        return cls(lib().ref_scene_load(os.path.join(d, "IndexBuffer.bin").encode(), os.path.join(d, "VertexBuffer.bin").encode()))

    @classmethod
    def from_batches(cls, batches, ref_min, ref_max) -> "RefScene":
        """batches: list of float32 arrays [nQuads*4, 4] (nQuads a multiple of 8)."""
        verts = np.ascontiguousarray(np.concatenate([np.asarray(b, np.float32).reshape(-1, 4) for b in batches]))
        counts = np.array([np.asarray(b).reshape(-1, 4).shape[0] // 4 for b in batches], np.uint32)
        mn = np.ascontiguousarray(ref_min, np.float32)
        mx = np.ascontiguousarray(ref_max, np.float32)
        return cls(lib().ref_scene_from_batches(_p(verts), _p(counts), len(batches), _p(mn), _p(mx)))

    def batch_vertices(self, i: int) -> np.ndarray:
        n = lib().ref_scene_batch_quads(self.h, i)
        out = np.zeros((n * 4, 4), np.float32)
        lib().ref_scene_batch_vertices(self.h, i, _p(out))
        return out

    def packed(self, i: int) -> np.ndarray:
        """Baked packets of occluder i as uint32 [packetCount*8] (Occluder.cpp:146-156 layout)."""
        n = int(self.packet_counts[i]) * 8
        buf = (C.c_uint32 * n).from_address(lib().ref_scene_occluder_packets(self.h, i))
        return np.frombuffer(buf, dtype=np.uint32).copy()

    def close(self):
        if self.h:
            lib().ref_scene_free(self.h)
            self.h = None


class RefRasterizer:
    def __init__(self, width: int, height: int):
        self.w, self.hgt = width, height
        self.blocks = (width // 8) * (height // 8)
        self.r = lib().ref_rast_create(width, height)

    def close(self):
        if self.r:
            lib().ref_rast_free(self.r)
            self.r = None

    def set_mvp(self, m):
        m = np.ascontiguousarray(m, np.float32)
        lib().ref_rast_set_mvp(self.r, _p(m))

    def clear(self, zero_depth=True):
        lib().ref_rast_clear(self.r, int(zero_depth))

    def rasterize(self, scene: RefScene, occ: int, clipped: bool):
        lib().ref_rast_rasterize(self.r, scene.h, occ, int(clipped))

    def query(self, bmin, bmax) -> int:
        a = np.ascontiguousarray(bmin, np.float32)
        b = np.ascontiguousarray(bmax, np.float32)
        return lib().ref_rast_query(self.r, _p(a), _p(b))

    def query2d(self, min_x, max_x, min_y, max_y, max_z) -> bool:
        return bool(lib().ref_rast_query2d(self.r, min_x, max_x, min_y, max_y, max_z))

    def query_boxes(self, boxes: np.ndarray) -> np.ndarray:
        boxes = np.ascontiguousarray(boxes, np.float32)
        out = np.zeros(boxes.shape[0], np.uint8)
        lib().ref_rast_query_boxes(self.r, _p(boxes), boxes.shape[0], _p(out))
        return out

    def frame(self, scene: RefScene, mvp, order, zero_depth=True):
        mvp = np.ascontiguousarray(mvp, np.float32)
        order = np.ascontiguousarray(order, np.uint32)
        gate = np.zeros(order.size, np.uint8)
        quads = lib().ref_rast_frame(self.r, scene.h, _p(mvp), _p(order), order.size, int(zero_depth), _p(gate))
        return gate, int(quads)

    def submit_all(self, scene: RefScene, mvp, order, clipped: bool, zero_depth=True):
        mvp = np.ascontiguousarray(mvp, np.float32)
        order = np.ascontiguousarray(order, np.uint32)
        lib().ref_rast_submit_all(self.r, scene.h, _p(mvp), _p(order), order.size, int(clipped), int(zero_depth))

    def depth(self, canonical=True) -> np.ndarray:
        out = np.zeros(self.blocks * 64, np.uint16)
        lib().ref_rast_get_depth(self.r, _p(out), int(canonical))
        return out

    def hiz(self) -> np.ndarray:
        out = np.zeros(self.blocks, np.uint16)
        lib().ref_rast_get_hiz(self.r, _p(out))
        return out

    def lut(self) -> np.ndarray:
        out = np.zeros(4096, np.int64)
        lib().ref_rast_get_lut(self.r, _p(out))
        return out

    def matrices(self):
        baked = np.zeros(16, np.float32)
        raw = np.zeros(16, np.float32)
        lib().ref_rast_get_matrices(self.r, _p(baked), _p(raw))
        return baked, raw

    def readback(self) -> np.ndarray:
        out = np.zeros(self.w * self.hgt * 4, np.uint8)
        lib().ref_rast_readback(self.r, _p(out))
        return out


def bench_views(scene: RefScene, w, h, mvps, orders, boxes, n_threads=1, reps=1):
    """Wall seconds + (frame thread-seconds, query thread-seconds, quads, visible) for the stock path."""
    mvps = np.ascontiguousarray(mvps, np.float32).reshape(-1, 16)
    orders = np.ascontiguousarray(orders, np.uint32).reshape(mvps.shape[0], -1)
    boxes = np.ascontiguousarray(boxes, np.float32).reshape(-1, 8)
    out = np.zeros(4, np.float64)
    wall = lib().ref_bench_views(scene.h, w, h, _p(mvps), _p(orders), mvps.shape[0], orders.shape[1], _p(boxes), boxes.shape[0], n_threads, reps, _p(out))
    return wall, out


MISMATCH_NAMES = {1: "gate", 2: "hiz", 4: "depth", 8: "vis", 16: "clip", 32: "quads"}


def check_views(scene: RefScene, w, h, mvps, orders, boxes=None, n_threads=0, mode=0, gate=None, depth=None, hiz=None, vis=None, clip=None,
                quads=None):
    """Render every view with the unmodified reference (fresh state per view, `n_threads` host threads, 0 = all) and
    compare bit for bit with the arrays given (any may be None): gate [n, nOcc] u8, depth [n, w*h] u16 (canonical),
    hiz [n, blocks] u16, vis / clip [n, ceil(nBoxes/32)] u32, quads [n] u32.  mode: 1 = no gate, 3 = no gate +
    rasterize<true>.  Returns the per-view mismatch masks (see MISMATCH_NAMES); all zero = parity."""
    mvps = np.ascontiguousarray(mvps, np.float32).reshape(-1, 16)
    n = mvps.shape[0]
    orders = np.ascontiguousarray(orders, np.uint32).reshape(n, -1)
    boxes = np.zeros((0, 8), np.float32) if boxes is None else np.ascontiguousarray(boxes, np.float32).reshape(-1, 8)
    blocks = (w // 8) * (h // 8)
    words = (boxes.shape[0] + 31) // 32

    def arr(a, dt, cols):
        if a is None:
            return None
        a = np.ascontiguousarray(a).view(dt) if np.asarray(a).dtype.itemsize == np.dtype(dt).itemsize else np.ascontiguousarray(a, dt)
        assert a.size == n * cols, (a.shape, n, cols)
        return a

    keep = [arr(gate, np.uint8, orders.shape[1]), arr(depth, np.uint16, blocks * 64), arr(hiz, np.uint16, blocks), arr(vis, np.uint32, words),
            arr(clip, np.uint32, words), arr(quads, np.uint32, 1)]
    mism = np.zeros(n, np.uint32)
    import os as _os
    lib().ref_check_views(scene.h, w, h, _p(mvps), _p(orders), n, orders.shape[1], _p(boxes), boxes.shape[0], n_threads or (_os.cpu_count() or 1), mode,
                          *[None if a is None else _p(a) for a in keep], _p(mism))
    return mism


def describe_mismatch(mism) -> str:
    bad = np.nonzero(mism)[0]
    if bad.size == 0:
        return "parity"
    kinds = sorted({name for bit, name in MISMATCH_NAMES.items() if (np.bitwise_or.reduce(mism) & bit)})
    return f"{bad.size} of {len(mism)} views differ ({', '.join(kinds)}); first views {bad[:8].tolist()} masks {mism[bad[:8]].tolist()}"


class RefPool:
    """One reference Rasterizer per host thread, built once and in parallel (bench.py --impl reference)."""

    def __init__(self, w, h, n_threads):
        self.n_threads = n_threads
        self.p = lib().ref_pool_create(w, h, n_threads)

    def bench(self, scene: RefScene, mvps, orders, boxes, reps=1):
        mvps = np.ascontiguousarray(mvps, np.float32).reshape(-1, 16)
        orders = np.ascontiguousarray(orders, np.uint32).reshape(mvps.shape[0], -1)
        boxes = np.ascontiguousarray(boxes, np.float32).reshape(-1, 8)
        out = np.zeros(4, np.float64)
        wall = lib().ref_pool_bench(self.p, scene.h, _p(mvps), _p(orders), mvps.shape[0], orders.shape[1], _p(boxes), boxes.shape[0], reps, _p(out))
        return wall, out

    def close(self):
        if self.p:
            lib().ref_pool_free(self.p)
            self.p = None


def fnv1a64(a: np.ndarray) -> str:
    """FNV-1a-64 over the bytes of `a` (the SURVEY's fingerprint function)."""
    h = 0xCBF29CE484222325
    data = np.ascontiguousarray(a).view(np.uint8).tobytes()
    # chunked pure-python loop is slow for MBs; use a vectorised-by-blocks C-speed fallback
    import zlib  # noqa: F401  (kept for parity of imports; the loop below is the definition)
    for byte in data:
        h = ((h ^ byte) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return f"{h:016x}"
