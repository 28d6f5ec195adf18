"""The scalar cores the CUDA kernels execute per lane (rasterizer_b200/csrc/orz_core.h), compiled
for the host by tests/core_host_shim.cpp, against the oracle port field by field: per-quad setup
records (both possiblyNearClipped variants), the queryVisibility front half, matrix baking.
CPU only -- catches transcription errors before GPU time is spent."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import port_oracle as po
from rasterizer_b200 import camera as cam
from rasterizer_b200 import workloads as wl

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def core():
    out = os.path.join(HERE, "_build", "libcore_host.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-mfma", "-fPIC", "-shared", "-o", out,
                           os.path.join(HERE, "core_host_shim.cpp")])
    return C.CDLL(out)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _scenes():
    out = [wl.synthetic_city(), wl.synthetic_soup(2048, cube=50.0)]
    if wl.have_scene("castle"):
        out.append(wl.load_scene("castle"))
    return out


def test_setup_records_match_port(core):
    po.set_tables()
    rcp = po.probe_host_rcp(11)
    checked = valid = 0
    for ps in _scenes():
        w, h = (1920, 1080) if ps.name == "castle" else (640, 360)
        cm = ps.camera
        views = [cam.view_projection(cm["pos"], cm["dir"], cm["up"], cm["fov"], w, h)]
        views += list(wl.camera_path(ps, 3, w, h)[0])
        port = po.PortRasterizer(w, h, np.zeros(4096, np.int64))
        for m in views:
            port.set_mvp(m)
            m = np.ascontiguousarray(m, np.float32)
            for b in ps.batches[:: max(1, len(ps.batches) // 12)]:
                packets, _, _, _ = po.bake(b, ps.ref_min, ps.ref_max)
                nq = packets.size // 4
                for q in range(0, nq, 3):
                    g, l = divmod(q, 8)
                    words = np.array([packets[(4 * g + j) * 8 + l] for j in range(4)], np.uint32)
                    for clipped in (0, 1):
                        ref = port.setup_quad(words, ps.ref_min, ps.ref_max, bool(clipped))
                        got = po.OrcPrim()
                        ok = core.core_setup_quad(_p(m), w, h, _p(ps.ref_min), _p(ps.ref_max), _p(words), clipped, _p(rcp), 11, C.byref(got))
                        assert bool(ok) == bool(ref.mode)
                        assert bytes(got) == bytes(ref), (ps.name, q, clipped)
                        checked += 1
                        valid += ok
        port.close()
    assert checked > 5000 and valid > 800


def test_box_front_half_composes_to_port_query(core):
    po.set_tables()
    rcp = po.probe_host_rcp(11)
    lut = po.build_lut()
    n_rect = n_clip = 0
    for ps in _scenes()[:3]:
        w, h = 640, 360
        cm = ps.camera
        packed = [po.bake(b, ps.ref_min, ps.ref_max) for b in ps.batches]
        port = po.PortRasterizer(w, h, lut)
        for m in [cam.view_projection(cm["pos"], cm["dir"], cm["up"], cm["fov"], w, h)] + list(wl.camera_path(ps, 2, w, h)[0]):
            m = np.ascontiguousarray(m, np.float32)
            port.clear(); port.set_mvp(m)
            for pk, _, bmin, bmax in packed[:40]:      # put something into the buffers
                g = port.query(bmin, bmax)
                if g & 1:
                    port.rasterize(pk, ps.ref_min, ps.ref_max, bool(g & 2))
            boxes = ps.quad_boxes()[::5]
            out = np.zeros(6, np.uint32)
            for bx in boxes:
                mn, mx = np.ascontiguousarray(bx[:4]), np.ascontiguousarray(bx[4:])
                core.core_box_front(_p(m), w, h, _p(mn), _p(mx), _p(rcp), 11, _p(out))
                want = port.query(mn, mx)
                if out[0] == 0:
                    assert want == 0
                elif out[0] == 1:
                    assert want == 3
                    n_clip += 1
                else:
                    assert want == int(port.query2d(int(out[1]), int(out[2]), int(out[3]), int(out[4]), int(out[5])))
                    n_rect += 1
        port.close()
    assert n_rect > 500


def test_view_matrices_match_port(core):
    ps = wl.synthetic_city()
    port = po.PortRasterizer(1920, 1080, np.zeros(4096, np.int64))
    for m in wl.camera_path(ps, 5, 1920, 1080)[0]:
        m = np.ascontiguousarray(m, np.float32)
        port.set_mvp(m)
        baked, raw = np.zeros(16, np.float32), np.zeros(16, np.float32)
        core.core_bake_view(_p(m), 1920, 1080, _p(baked), _p(raw))
        pb, pr = port.matrices()
        assert np.array_equal(baked.view(np.uint32), pb.view(np.uint32))
        assert np.array_equal(raw.view(np.uint32), pr.view(np.uint32))
    port.close()
