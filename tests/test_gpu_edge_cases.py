"""Edge cases of the CUDA path through the C ABI: empty and ragged inputs, minimum sizes, chunked
batches, degenerate / random boxes, cleared buffers, argument errors."""
import ctypes as C

import numpy as np
import pytest

from oracle import port_oracle as po
from rasterizer_b200 import api
from rasterizer_b200 import camera as cam
from rasterizer_b200 import workloads as wl

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = api.Context(0)
    po.set_tables()
    yield c
    c.close()


@pytest.fixture(scope="module")
def city(ctx):
    ps = wl.synthetic_city()
    baked = [api.bake(b, ps.ref_min, ps.ref_max) for b in ps.batches]
    return ps, baked


def _scene(ctx, ps, baked, boxes=None, sel=None):
    sel = range(len(baked)) if sel is None else sel
    return api.Scene(ctx, [baked[i][0] for i in sel], ps.ref_min, ps.ref_max, np.stack([baked[i][2] for i in sel]),
                     np.stack([baked[i][3] for i in sel]), np.stack([baked[i][1] for i in sel]), boxes)


def test_empty_batch_and_no_occludees(ctx, city):
    ps, baked = city
    sc = _scene(ctx, ps, baked)           # no occludee boxes at all
    out = sc.render_views(640, 360, np.zeros((0, 16), np.float32), cam_pos=np.zeros((0, 3), np.float32), want=("vis", "gate"))
    assert out["vis"].shape == (0, 0) and out["gate"].shape[0] == 0
    mvps, poss = wl.camera_path(ps, 2, 640, 360)
    out = sc.render_views(640, 360, mvps, cam_pos=poss, want=("vis", "gate", "depth", "hiz"))
    port = po.PortRasterizer(640, 360)
    for v in range(2):
        order = cam.front_to_back_order(sc.centers, poss[v])
        gate, _ = port.frame(sc.packed_list, sc.bounds_min, sc.bounds_max, ps.ref_min, ps.ref_max, mvps[v], order)
        assert np.array_equal(out["gate"][v], gate) and np.array_equal(out["hiz"][v], port.hiz()) and np.array_equal(out["depth"][v], port.depth())
    sc.close(); port.close()


@pytest.mark.parametrize("size", [(8, 8), (16, 8), (64, 64), (264, 136)])
def test_minimum_and_odd_sizes(ctx, city, size):
    """1-block and odd block-count targets; a single occluder of one 8-quad group."""
    ps, baked = city
    w, h = size
    one = (baked[0][0][:32].copy(), baked[0][1], baked[0][2], baked[0][3])      # first packet group only
    sc = api.Scene(ctx, [one[0]], ps.ref_min, ps.ref_max, one[2][None], one[3][None], one[1][None], ps.quad_boxes()[:40])
    c = ps.camera
    mvps = np.stack([cam.view_projection(c["pos"], c["dir"], c["up"], c["fov"], w, h)] * 3)
    order = np.zeros((3, 1), np.uint32)
    blocks = (w // 8) * (h // 8)
    port = po.PortRasterizer(w, h)
    gate, _ = port.frame([one[0]], one[2][None], one[3][None], ps.ref_min, ps.ref_max, mvps[0], order[0])
    want_q = port.query_boxes(ps.quad_boxes()[:40])
    if blocks % 8 == 0:
        out = sc.render_views(w, h, mvps, orders=order, want=("vis", "gate", "depth", "hiz"))
        assert np.array_equal(out["hiz"][2], port.hiz()) and np.array_equal(out["depth"][2], port.depth())
    else:
        with pytest.raises(api.OrzError, match="multiple of 8"):
            sc.render_views(w, h, mvps, orders=order, want=("vis", "depth", "hiz"))
        out = sc.render_views(w, h, mvps, orders=order, want=("vis", "gate"))    # internal targets: any size
    assert np.array_equal(out["gate"][1], gate)
    assert np.array_equal(api.unpack_bits(out["vis"], 40)[1], (want_q & 1).astype(bool))
    # per-call API on the same target
    r = api.Rasterizer(ctx, w, h)
    occ = api.Occluder(ctx, one[0], ps.ref_min, ps.ref_max)
    r.clear(); r.setModelViewProjection(mvps[0])
    if gate[0] & 1:
        r.rasterize(occ, bool(gate[0] & 2))
    d, hz = r.download()
    assert np.array_equal(hz, port.hiz()) and np.array_equal(d, port.depth())
    assert np.array_equal(r.query_boxes(ps.quad_boxes()[:40]), want_q)
    occ.close(); r.close(); sc.close(); port.close()


def test_chunked_batch_matches_unchunked(ctx, city):
    """A tiny arena forces the batch through several chunks of views."""
    ps, baked = city
    sc = _scene(ctx, ps, baked, ps.quad_boxes())
    mvps, poss = wl.camera_path(ps, 23, 320, 184)
    ref = sc.render_views(320, 184, mvps, cam_pos=poss, want=("vis", "clip", "gate", "quads"))
    ctx.set_arena_bytes(5 * (40 * 23 * 130))        # room for ~5 views
    try:
        got = sc.render_views(320, 184, mvps, cam_pos=poss, want=("vis", "clip", "gate", "quads"))
    finally:
        ctx.set_arena_bytes(8 << 30)
    for k in ref:
        assert np.array_equal(ref[k], got[k]), k
    sc.close()


def test_random_and_degenerate_boxes(ctx, city):
    ps, baked = city
    rng = np.random.default_rng(11)
    lo, hi = ps.ref_min[:3] - 30, ps.ref_max[:3] + 30
    a = rng.uniform(lo, hi, (3000, 3)); b = a + rng.uniform(0, 1, (3000, 3)) * rng.choice([0.0, 0.5, 5.0, 60.0, 400.0], (3000, 1))
    boxes = np.concatenate([a, np.ones((3000, 1)), b, np.ones((3000, 1))], axis=1).astype(np.float32)
    boxes[:5, 4:7] = boxes[:5, 0:3]                      # zero-size boxes
    boxes[5, 0:3], boxes[5, 4:7] = lo - 1000, hi + 1000  # box containing everything (near clip)
    boxes[6, 0:3], boxes[6, 4:7] = (1e30, 1e30, 1e30), (3e38, 3e38, 3e38)   # overflowing box (inf/NaN arise inside)
    boxes[7, 4:7] = boxes[7, 0:3] - 1.0                  # inverted box
    sc = _scene(ctx, ps, baked, boxes)
    mvps, poss = wl.camera_path(ps, 3, 640, 360)
    out = sc.render_views(640, 360, mvps, cam_pos=poss, want=("vis", "clip"))
    vis, clip = api.unpack_bits(out["vis"], 3000), api.unpack_bits(out["clip"], 3000)
    port = po.PortRasterizer(640, 360)
    # NaN *inputs* are outside the contract: the reference's answer then depends on which operand's
    # NaN sign x86 propagates (movemask of NaN distances, Rasterizer.cpp:161-164); GPUs canonicalise NaNs.
    for v in range(3):
        port.frame(sc.packed_list, sc.bounds_min, sc.bounds_max, ps.ref_min, ps.ref_max, mvps[v], cam.front_to_back_order(sc.centers, poss[v]))
        q = port.query_boxes(boxes)
        bad = np.nonzero((vis[v] != (q & 1).astype(bool)) | (clip[v] != (q & 2).astype(bool)))[0]
        assert bad.size == 0, (v, bad[:10], boxes[bad[:3]])
    # cleared buffers: everything inside the frustum is visible
    r = api.Rasterizer(ctx, 640, 360)
    r.clear(); r.setModelViewProjection(mvps[0])
    port.clear(); port.set_mvp(mvps[0])
    assert np.array_equal(r.query_boxes(boxes), port.query_boxes(boxes))
    r.close(); sc.close(); port.close()


def test_all_quads_culled_leaves_buffers_cleared(ctx, city):
    ps, baked = city
    w, h = 320, 184
    c = ps.camera
    away = cam.view_projection(c["pos"], tuple(-np.asarray(c["dir"])), c["up"], c["fov"], w, h)   # looking away from the city
    r = api.Rasterizer(ctx, w, h)
    r.clear(); r.setModelViewProjection(away)
    for i in range(0, len(baked), 3):
        occ = api.Occluder(ctx, baked[i][0], ps.ref_min, ps.ref_max)
        r.rasterize(occ, False)
        occ.close()
    d, hz = r.download()
    port = po.PortRasterizer(w, h); port.clear(); port.set_mvp(away)
    for i in range(0, len(baked), 3):
        port.rasterize(baked[i][0], ps.ref_min, ps.ref_max, False)
    assert np.array_equal(hz, port.hiz()) and np.array_equal(d, port.depth())
    r.close(); port.close()


def test_argument_errors(ctx, city):
    ps, baked = city
    with pytest.raises(api.OrzError):
        api.Rasterizer(ctx, 100, 64)                 # width % 8 != 0 (Rasterizer.cpp:68 asserts)
    with pytest.raises(api.OrzError):
        api.Rasterizer(ctx, 0, 64)
    with pytest.raises(ValueError):
        api.bake(np.zeros((12, 4), np.float32), ps.ref_min, ps.ref_max)   # not a multiple of 8 quads
    sc = _scene(ctx, ps, baked, ps.quad_boxes()[:10])
    with pytest.raises(api.OrzError):
        sc.render_views(100, 64, np.zeros((1, 16), np.float32), cam_pos=np.zeros((1, 3), np.float32))
    r = api.Rasterizer(ctx, 64, 64)
    with pytest.raises(api.OrzError):
        r.query2D(0, 64, 0, 10, 5)                   # rectangle outside the buffer
    r.close(); sc.close()


@pytest.mark.parametrize("name", ["city", "castle", "soup"])
def test_device_bake_matches_host_bake(ctx, name):
    """Occluder::bake on the GPU (orz_scene_bake, one CTA per batch) against the host bake: packets in the
    reference layout, bounds and centres bit for bit; a scene baked on the device renders the same."""
    if name == "soup":
        ps = wl.synthetic_soup(4096, cube=60.0)
    elif name == "city":
        ps = wl.synthetic_city()
    else:
        if not wl.have_scene("castle"):
            pytest.skip("castle scene data not shipped")
        ps = wl.load_scene("castle")
    host = [api.bake(b, ps.ref_min, ps.ref_max) for b in ps.batches]
    boxes = ps.quad_boxes()[::11]
    sd = api.Scene.bake_on_device(ctx, ps.batches, ps.ref_min, ps.ref_max, boxes)
    for i, hb in enumerate(host):
        assert np.array_equal(sd.packed_list[i], hb[0]), (name, i)
        assert np.array_equal(sd.centers[i].view(np.uint32), hb[1].view(np.uint32)), (name, i)
        assert np.array_equal(sd.bounds_min[i].view(np.uint32), hb[2].view(np.uint32)), (name, i)
        assert np.array_equal(sd.bounds_max[i].view(np.uint32), hb[3].view(np.uint32)), (name, i)
    sh = api.Scene(ctx, [b[0] for b in host], ps.ref_min, ps.ref_max, np.stack([b[2] for b in host]), np.stack([b[3] for b in host]),
                   np.stack([b[1] for b in host]), boxes)
    mvps, poss = wl.camera_path(ps, 3, 640, 360)
    a = sd.render_views(640, 360, mvps, cam_pos=poss, want=("vis", "gate", "depth", "hiz"))
    b = sh.render_views(640, 360, mvps, cam_pos=poss, want=("vis", "gate", "depth", "hiz"))
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    sd.close(); sh.close()


def test_device_bake_signed_zero_bounds_and_errors(ctx):
    """minps / maxps keep the earlier operand on ties: +0 / -0 in the bounds must follow the serial order."""
    rng = np.random.default_rng(5)
    v = rng.uniform(0.5, 4.0, (64, 4)).astype(np.float32)
    v[:, 3] = 1.0
    v[40, 0] = 0.0; v[41, 0] = -0.0; v[50, 1] = -0.0; v[60, 1] = 0.0     # the minima of x and y are zeros of both signs
    v[:, 2] = -np.abs(v[:, 2]); v[7, 2] = -0.0; v[30, 2] = 0.0            # the maximum of z too
    rmn, rmx = np.array([-1, -1, -5, 0], np.float32), np.array([5, 5, 1, 0], np.float32)
    hb = api.bake(v, rmn, rmx)
    sd = api.Scene.bake_on_device(ctx, [v], rmn, rmx)
    assert np.array_equal(sd.packed_list[0], hb[0])
    assert np.array_equal(sd.bounds_min[0].view(np.uint32), hb[2].view(np.uint32))
    assert np.array_equal(sd.bounds_max[0].view(np.uint32), hb[3].view(np.uint32))
    assert np.array_equal(sd.centers[0].view(np.uint32), hb[1].view(np.uint32))
    sd.close()
    with pytest.raises(api.OrzError, match="multiple of 8 quads"):
        api.Scene.bake_on_device(ctx, [v[:36]], rmn, rmx)
