"""Pins the oracle: the scalar C restatement (oracle/oracle_port.c) must reproduce the UNMODIFIED
reference sources (oracle/_ref/libref_oracle.so, built from /root/reference by oracle/Makefile)
bit-for-bit -- depth, HiZ, every gate decision, occludee bits, baked words, LUT, matrices,
read-back image -- on the reference's own scenes and on synthetic ones.  CPU only."""
import numpy as np
import pytest

from oracle import port_oracle as po
from oracle import ref_oracle as ro
from rasterizer_b200 import camera as cam
from rasterizer_b200 import workloads as wl

needs_ref = pytest.mark.skipif(not ro.available(), reason="oracle/_ref/libref_oracle.so not built (needs /root/reference)")


@pytest.fixture(scope="module")
def lut():
    return po.build_lut()


@pytest.fixture(scope="module", autouse=True)
def host_tables():
    po.set_tables()  # the reference build uses this host's rcpps / rsqrtps
    yield


def _ref_scene(name):
    if not ro.scene_available(name):
        pytest.skip(f"no {name} data under oracle/_ref/scenes")
    return ro.RefScene.load(name)


def _compare_frames(s, packed, w, h, mvps, poss, lut, n_boxes=2000):
    r, p = ro.RefRasterizer(w, h), po.PortRasterizer(w, h, lut)
    boxes = s.boxes[:: max(1, s.boxes.shape[0] // n_boxes)] if s.boxes.shape[0] else s.boxes
    for m, pos in zip(mvps, poss):
        order = cam.front_to_back_order(s.centers, pos)
        g1, q1 = r.frame(s, m, order)
        g2, q2 = p.frame(packed, s.bounds_min, s.bounds_max, s.ref_min, s.ref_max, m, order)
        assert np.array_equal(g1, g2) and q1 == q2
        assert np.array_equal(r.hiz(), p.hiz())
        assert np.array_equal(r.depth(), p.depth())
        if boxes.shape[0]:
            assert np.array_equal(r.query_boxes(boxes), p.query_boxes(boxes))
    r.close(); p.close()


@needs_ref
def test_rcp_rsqrt_model_matches_host_instruction():
    rng = np.random.default_rng(1)
    x = rng.integers(0, 2**32, 300000, dtype=np.uint64).astype(np.uint32)
    edge = np.array([0, 0x80000000, 1, 0x007fffff, 0x00800000, 0x7f7fffff, 0x7f800000, 0xff800000, 0x7fc00000, 0x7f800001,
                     0x7e800000, 0x7e7fffff, 0x7f000000, 0x3f800000, 0x3fc00000, 0x40400000], np.uint32)
    x = np.concatenate([x, edge]).view(np.float32)
    assert np.array_equal(ro.host_rcp(x).view(np.uint32), po.rcp(x).view(np.uint32))
    assert np.array_equal(ro.host_rsqrt(x).view(np.uint32), po.rsqrt(x).view(np.uint32))


@needs_ref
def test_lut_matrices_readback(lut):
    r = ro.RefRasterizer(256, 128)
    assert np.array_equal(lut, r.lut())
    assert (lut == 0).sum() == 64 and (lut == -1).sum() == 128
    s = wl.synthetic_city()
    p = po.PortRasterizer(256, 128, lut)
    m = cam.view_projection(s.camera["pos"], s.camera["dir"], s.camera["up"], s.camera["fov"], 256, 128)
    r.set_mvp(m); p.set_mvp(m)
    for a, b in zip(r.matrices(), p.matrices()):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    r.close(); p.close()


@needs_ref
def test_castle_1080p_views(lut):
    s = _ref_scene("Castle")
    packed = [s.packed(i) for i in range(s.n_occluders)]
    ps = wl.load_scene("castle") if wl.have_scene("castle") else None
    c = cam.CASTLE_CAMERA
    mvps = [cam.view_projection(c["pos"], c["dir"], c["up"], c["fov"], 1920, 1080)]
    poss = [np.array(c["pos"], np.float32)]
    if ps is not None:
        pm, pp = wl.camera_path(ps, 12, 1920, 1080)
        mvps += list(pm[::2]); poss += list(pp[::2])
    _compare_frames(s, packed, 1920, 1080, mvps, poss, lut)
    # read-back image after the last view
    s.close()


@needs_ref
def test_castle_probe_views_512x256(lut):
    s = _ref_scene("Castle")
    if not wl.have_scene("castle"):
        pytest.skip("prepared castle scene missing")
    packed = [s.packed(i) for i in range(s.n_occluders)]
    mvps, poss = wl.probe_views(wl.load_scene("castle"), 24, 512, 256)
    _compare_frames(s, packed, 512, 256, mvps, poss, lut, n_boxes=800)
    s.close()


@needs_ref
def test_sponza_1080p(lut):
    s = _ref_scene("Sponza")
    packed = [s.packed(i) for i in range(s.n_occluders)]
    c = cam.SPONZA_CAMERA
    mvps = [cam.view_projection(c["pos"], c["dir"], c["up"], c["fov"], 1920, 1080),
            cam.view_projection((2.0, -3.0, 4.0), (-0.6, 0.75, -0.1), c["up"], c["fov"], 1920, 1080)]
    poss = [np.array(c["pos"], np.float32), np.array((2.0, -3.0, 4.0), np.float32)]
    _compare_frames(s, packed, 1920, 1080, mvps, poss, lut)
    s.close()


@needs_ref
@pytest.mark.parametrize("size", [(1280, 720), (3840, 2160)])
def test_forced_near_clip_path_and_4k_wrap(lut, size):
    """Every batch through rasterize<true> with no gate (config 4 shape); 3840x2160 additionally
    exercises the 16-bit block-row wrap of Rasterizer.cpp:1054."""
    s = _ref_scene("Castle")
    packed = [s.packed(i) for i in range(s.n_occluders)]
    w, h = size
    r, p = ro.RefRasterizer(w, h), po.PortRasterizer(w, h, lut)
    c = cam.CASTLE_CAMERA
    for pos, d in ((c["pos"], c["dir"]), ((92.0, 6.0, -4.0), (0.3, -0.1, 0.9))):
        m = cam.view_projection(pos, d, c["up"], c["fov"], w, h)
        order = cam.front_to_back_order(s.centers, pos)
        r.submit_all(s, m, order, True)
        p.clear(); p.set_mvp(m)
        for o in order:
            p.rasterize(packed[o], s.ref_min, s.ref_max, True)
        assert np.array_equal(r.hiz(), p.hiz())
        assert np.array_equal(r.depth(), p.depth())
    assert np.array_equal(r.readback(), p.readback())
    r.close(); p.close(); s.close()


@needs_ref
def test_bake_matches_reference():
    for name, stride in (("Castle", 1), ("Sponza", 9)):
        if not ro.scene_available(name):
            continue
        s = ro.RefScene.load(name)
        for i in range(0, s.n_occluders, stride):
            packets, c, bmin, bmax = po.bake(s.batch_vertices(i), s.ref_min, s.ref_max)
            assert np.array_equal(packets, s.packed(i)), (name, i)
            assert np.array_equal(c.view(np.uint32), s.centers[i].view(np.uint32))
            assert np.array_equal(bmin.view(np.uint32), s.bounds_min[i].view(np.uint32))
            assert np.array_equal(bmax.view(np.uint32), s.bounds_max[i].view(np.uint32))
        s.close()


@needs_ref
def test_synthetic_city_and_soup(lut):
    for ps, size, clip_all in ((wl.synthetic_city(), (640, 360), False), (wl.synthetic_soup(4096, cube=60.0), (640, 360), True)):
        s = ro.RefScene.from_batches(ps.batches, ps.ref_min, ps.ref_max)
        packed = []
        for i, b in enumerate(ps.batches):
            pk, c, bmin, bmax = po.bake(b, ps.ref_min, ps.ref_max)
            assert np.array_equal(pk, s.packed(i))
            packed.append(pk)
        w, h = size
        cm = ps.camera
        m = cam.view_projection(cm["pos"], cm["dir"], cm["up"], cm["fov"], w, h)
        order = cam.front_to_back_order(s.centers, cm["pos"])
        r, p = ro.RefRasterizer(w, h), po.PortRasterizer(w, h, lut)
        if clip_all:
            r.submit_all(s, m, order, True)
            p.clear(); p.set_mvp(m)
            for o in order:
                p.rasterize(packed[o], ps.ref_min, ps.ref_max, True)
        else:
            g1, _ = r.frame(s, m, order)
            g2, _ = p.frame(packed, s.bounds_min, s.bounds_max, ps.ref_min, ps.ref_max, m, order)
            assert np.array_equal(g1, g2)
            assert (g1 & 1).sum() > 3
        assert (r.hiz() != 1).sum() > 50          # the scene really covers pixels
        assert np.array_equal(r.hiz(), p.hiz())
        assert np.array_equal(r.depth(), p.depth())
        boxes = ps.quad_boxes()[::3]
        assert np.array_equal(r.query_boxes(boxes), p.query_boxes(boxes))
        r.close(); p.close(); s.close()


@needs_ref
def test_block_bound_skip_changes_nothing(lut):
    """Planning switch of the port: updates whose largest corner sample is not above the block's
    HiZ are skipped -- a quarter of Sponza's updates -- and depth, HiZ, gates and queries still equal the reference."""
    import ctypes as C

    L = po.lib()
    L.orc_set_block_bound_skip.argtypes, L.orc_stats.argtypes = [C.c_int], [C.c_void_p, C.c_int]
    s = _ref_scene("Sponza")
    packed = [s.packed(i) for i in range(s.n_occluders)]
    c = cam.SPONZA_CAMERA
    views = [(c["pos"], c["dir"]), ((2.0, -3.0, 4.0), (-0.6, 0.75, -0.1)), ((-6.0, 1.0, 2.5), (0.9, -0.2, 0.1))]
    mvps = [cam.view_projection(p, d, c["up"], c["fov"], 1280, 720) for p, d in views]
    poss = [np.array(p, np.float32) for p, _ in views]
    L.orc_set_block_bound_skip(1)
    L.orc_stats(None, 1)
    try:
        _compare_frames(s, packed, 1280, 720, mvps, poss, lut)
        st = np.zeros(5, np.uint64)
        L.orc_stats(st.ctypes.data_as(C.c_void_p), 1)
        assert st[4] > 0.1 * st[2], "the bound should have skipped a good part of the updates"
    finally:
        L.orc_set_block_bound_skip(0)
    s.close()
