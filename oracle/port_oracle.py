"""TEST INFRASTRUCTURE ONLY: ctypes view of oracle/liboracle_port.so (oracle/oracle_port.c), the
scalar C restatement of the reference hot path.  Built by `make -C oracle port`."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle_port.so")
_lib = None


class OrcPrim(C.Structure):
    _fields_ = [("mode", C.c_uint32), ("minX", C.c_int32), ("minY", C.c_int32), ("rangeX", C.c_int32), ("rangeY", C.c_int32),
                ("maxZ", C.c_uint32), ("dzdx", C.c_float), ("dzdy", C.c_float), ("plane0", C.c_float),
                ("nx", C.c_float * 4), ("ny", C.c_float * 4), ("off", C.c_float * 4), ("slope", C.c_uint32 * 4)]


def build():
    subprocess.check_call(["make", "-C", HERE, "port"], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        vp, u32 = C.c_void_p, C.c_uint32
        sig = {
            "orc_set_rcp_table": (None, [vp, C.c_int]),
            "orc_set_rsqrt_table": (None, [vp, C.c_int]),
            "orc_use_host_tables": (None, []),
            "orc_rcp": (C.c_float, [C.c_float]),
            "orc_rsqrt": (C.c_float, [C.c_float]),
            "orc_probe_host_rcp": (None, [vp, C.c_int]),
            "orc_probe_host_rsqrt": (None, [vp, C.c_int]),
            "orc_build_lut": (None, [vp]),
            "orc_create": (vp, [u32, u32, vp]),
            "orc_destroy": (None, [vp]),
            "orc_set_mvp": (None, [vp, vp]),
            "orc_clear": (None, [vp]),
            "orc_rasterize": (None, [vp, vp, u32, vp, vp, C.c_int]),
            "orc_query_visibility": (C.c_int, [vp, vp, vp]),
            "orc_query2d": (C.c_int, [vp, u32, u32, u32, u32, u32]),
            "orc_readback_depth": (None, [vp, vp]),
            "orc_depth": (vp, [vp]),
            "orc_hiz": (vp, [vp]),
            "orc_lut": (vp, [vp]),
            "orc_get_matrices": (None, [vp, vp, vp]),
            "orc_bake": (u32, [vp, u32, vp, vp, vp, vp, vp, vp]),
            "orc_setup_quad": (None, [vp, vp, vp, vp, C.c_int, vp]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def probe_host_rcp(bits=11) -> np.ndarray:
    t = np.zeros(1 << bits, np.uint32)
    lib().orc_probe_host_rcp(_p(t), bits)
    return t


def probe_host_rsqrt(bits=10) -> np.ndarray:
    t = np.zeros(2 << bits, np.uint32)
    lib().orc_probe_host_rsqrt(_p(t), bits)
    return t


def set_tables(rcp: np.ndarray | None = None, rsqrt: np.ndarray | None = None):
    """Explicit rcpps / rsqrtps models (e.g. the committed Intel tables) or the host's own (None)."""
    if rcp is None and rsqrt is None:
        lib().orc_use_host_tables()
        return
    if rcp is not None:
        rcp = np.ascontiguousarray(rcp, np.uint32)
        lib().orc_set_rcp_table(_p(rcp), int(np.log2(rcp.size)))
    if rsqrt is not None:
        rsqrt = np.ascontiguousarray(rsqrt, np.uint32)
        lib().orc_set_rsqrt_table(_p(rsqrt), int(np.log2(rsqrt.size)) - 1)


def rcp(x: np.ndarray) -> np.ndarray:
    f = lib().orc_rcp
    return np.array([f(float(v)) for v in np.asarray(x, np.float32).ravel()], np.float32)


def rsqrt(x: np.ndarray) -> np.ndarray:
    f = lib().orc_rsqrt
    return np.array([f(float(v)) for v in np.asarray(x, np.float32).ravel()], np.float32)


_LUT = None


def build_lut() -> np.ndarray:
    global _LUT
    if _LUT is None:
        t = np.zeros(4096, np.int64)
        lib().orc_build_lut(_p(t))
        _LUT = t
    return _LUT.copy()


def bake(vertices: np.ndarray, ref_min, ref_max):
    """-> (packets uint32 [nVerts] in the reference layout, center, boundsMin, boundsMax)."""
    v = np.ascontiguousarray(vertices, np.float32).reshape(-1, 4)
    mn = np.ascontiguousarray(ref_min, np.float32)
    mx = np.ascontiguousarray(ref_max, np.float32)
    packets = np.zeros(v.shape[0], np.uint32)
    c, bmin, bmax = np.zeros(4, np.float32), np.zeros(4, np.float32), np.zeros(4, np.float32)
    n = lib().orc_bake(_p(v), v.shape[0], _p(mn), _p(mx), _p(packets), _p(c), _p(bmin), _p(bmax))
    assert n * 8 == v.shape[0]
    return packets, c, bmin, bmax


class PortRasterizer:
    def __init__(self, width: int, height: int, lut: np.ndarray | None = None):
        self.w, self.hgt = width, height
        self.blocks = (width // 8) * (height // 8)
        lut = build_lut() if lut is None else np.ascontiguousarray(lut, np.int64)
        self.r = lib().orc_create(width, height, _p(lut))

    def close(self):
        if self.r:
            lib().orc_destroy(self.r)
            self.r = None

    def set_mvp(self, m):
        m = np.ascontiguousarray(m, np.float32)
        lib().orc_set_mvp(self.r, _p(m))

    def clear(self):
        lib().orc_clear(self.r)

    def rasterize(self, packets: np.ndarray, ref_min, ref_max, clipped: bool):
        packets = np.ascontiguousarray(packets, np.uint32)
        mn = np.ascontiguousarray(ref_min, np.float32)
        mx = np.ascontiguousarray(ref_max, np.float32)
        lib().orc_rasterize(self.r, _p(packets), packets.size // 8, _p(mn), _p(mx), int(clipped))

    def query(self, bmin, bmax) -> int:
        a = np.ascontiguousarray(bmin, np.float32)
        b = np.ascontiguousarray(bmax, np.float32)
        return lib().orc_query_visibility(self.r, _p(a), _p(b))

    def query2d(self, min_x, max_x, min_y, max_y, max_z) -> bool:
        return bool(lib().orc_query2d(self.r, min_x, max_x, min_y, max_y, max_z))

    def query_boxes(self, boxes: np.ndarray) -> np.ndarray:
        boxes = np.ascontiguousarray(boxes, np.float32).reshape(-1, 8)
        f, r = lib().orc_query_visibility, self.r
        base = boxes.ctypes.data
        return np.array([f(r, base + 32 * i, base + 32 * i + 16) for i in range(boxes.shape[0])], np.uint8)

    def frame(self, packed_list, bounds_min, bounds_max, ref_min, ref_max, mvp, order):
        """Main.cpp:181-206 with a given order -> (gate bits per order slot, quads submitted)."""
        self.clear()
        self.set_mvp(mvp)
        gate = np.zeros(len(order), np.uint8)
        quads = 0
        for i, o in enumerate(order):
            g = self.query(bounds_min[o], bounds_max[o])
            gate[i] = g
            if g & 1:
                self.rasterize(packed_list[o], ref_min, ref_max, bool(g & 2))
                quads += packed_list[o].size // 4
        return gate, quads

    def setup_quad(self, words, ref_min, ref_max, clipped: bool) -> OrcPrim:
        w = np.ascontiguousarray(words, np.uint32)
        mn = np.ascontiguousarray(ref_min, np.float32)
        mx = np.ascontiguousarray(ref_max, np.float32)
        out = OrcPrim()
        lib().orc_setup_quad(self.r, _p(w), _p(mn), _p(mx), int(clipped), C.byref(out))
        return out

    def depth(self) -> np.ndarray:
        buf = (C.c_uint16 * (self.blocks * 64)).from_address(lib().orc_depth(self.r))
        d = np.frombuffer(buf, dtype=np.uint16).copy()
        # canonical form: cleared blocks read as zero
        d.reshape(-1, 64)[self.hiz() == 1] = 0
        return d

    def hiz(self) -> np.ndarray:
        buf = (C.c_uint16 * self.blocks).from_address(lib().orc_hiz(self.r))
        return np.frombuffer(buf, dtype=np.uint16).copy()

    def matrices(self):
        baked, raw = np.zeros(16, np.float32), np.zeros(16, np.float32)
        lib().orc_get_matrices(self.r, _p(baked), _p(raw))
        return baked, raw

    def readback(self) -> np.ndarray:
        out = np.zeros(self.w * self.hgt * 4, np.uint8)
        lib().orc_readback_depth(self.r, _p(out))
        return out
