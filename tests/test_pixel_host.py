"""The eighth-of-a-block update the cluster kernel runs on eight lanes per block (rasterizer_b200/csrc/orz_pixel.h),
compiled for the host, against the lane-per-block form of Rasterizer.cpp:1241-1290 the round-1 kernels used: random and
adversarial depth lanes (huge, negative, denormal, inf, NaN), random coverage masks, cleared and written blocks.
CPU only -- the arithmetic of the new lane mapping is pinned before GPU time is spent."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def shim():
    out = os.path.join(HERE, "_build", "libpixel_host.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-mfma", "-fPIC", "-shared", "-o", out,
                           os.path.join(HERE, "pixel_host_shim.cpp")])
    L = C.CDLL(out)
    for fn in (L.pixel_block_lane, L.pixel_block_items):
        fn.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        fn.restype = None
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_items_equal_lane_per_block(shim):
    rng = np.random.default_rng(11)
    specials = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e-42, -1e-42, 3.4e38, -3.4e38, 1.0, 0.99999994, 2.0 ** -12, 65535.0 * 2.0 ** -115], np.float32)
    n = 0
    for trial in range(20000):
        kind = trial % 5
        if kind == 0:  # plausible depth values: bit patterns around the packing range
            dv = (rng.integers(0x0F000000, 0x10100000, 8, dtype=np.uint32)).view(np.float32).copy()
            dzdx, dzdy = (rng.normal(0, 1, 2) * float(np.float32(2.0) ** -100)).astype(np.float32)
        elif kind == 1:
            dv = rng.normal(0, 1, 8).astype(np.float32)
            dzdx, dzdy = rng.normal(0, 0.1, 2).astype(np.float32)
        elif kind == 2:  # arbitrary bit patterns
            dv = rng.integers(0, 2 ** 32, 8, dtype=np.uint64).astype(np.uint32).view(np.float32).copy()
            dzdx, dzdy = rng.integers(0, 2 ** 32, 2, dtype=np.uint64).astype(np.uint32).view(np.float32)
        elif kind == 3:
            dv = rng.choice(specials, 8).astype(np.float32)
            dzdx, dzdy = rng.choice(specials, 2).astype(np.float32)
        else:
            dv = (rng.integers(0x0F000000, 0x10100000, 8, dtype=np.uint32)).view(np.float32).copy()
            dv[rng.integers(0, 8)] = rng.choice(specials)
            dzdx, dzdy = (rng.normal(0, 1, 2) * float(np.float32(2.0) ** -103)).astype(np.float32)
        mk = rng.integers(0, 2 ** 32, 2, dtype=np.uint64).astype(np.uint32)
        if trial % 7 == 0:
            mk[:] = 0
        if trial % 11 == 0:
            mk[:] = 0xFFFFFFFF
        h_old = int(rng.choice([1, 0, 7, 65535, int(rng.integers(0, 65536))]))
        d0 = rng.integers(0, 2 ** 32, 32, dtype=np.uint64).astype(np.uint32)
        if trial % 3 == 0:
            d0 = (d0 & 0x00FF00FF).astype(np.uint32)
        da, db = d0.copy(), d0.copy()
        ha, hb = np.zeros(1, np.uint32), np.zeros(1, np.uint32)
        shim.pixel_block_lane(_p(dv), C.c_float(float(dzdx)), C.c_float(float(dzdy)), int(mk[0]), int(mk[1]), h_old, _p(da), _p(ha))
        shim.pixel_block_items(_p(dv), C.c_float(float(dzdx)), C.c_float(float(dzdy)), int(mk[0]), int(mk[1]), h_old, _p(db), _p(hb))
        assert np.array_equal(da, db), (trial, dv.view(np.uint32), mk, h_old)
        assert ha[0] == hb[0], (trial, ha, hb)
        n += 1
    assert n == 20000
