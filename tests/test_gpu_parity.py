"""GPU parity tests proper: the CUDA path, called through the C ABI (rasterizer_b200.api ->
librasterizer_b200.so), against the oracle on the same seeded inputs.  Bit-exact: depth, HiZ,
gate decisions, occludee visibility / needsClipping bits, setup records, read-back image.
Run on the B200 box: python -m pytest tests -m gpu -x -q"""
import numpy as np
import pytest

from oracle import port_oracle as po
from oracle import ref_oracle as ro
from rasterizer_b200 import api
from rasterizer_b200 import camera as cam
from rasterizer_b200 import workloads as wl

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = api.Context(0)
    # oracle and GPU must model the same rcpps: the one of the CPU this test runs on
    po.set_tables()
    assert np.array_equal(c.rcp_table(), po.probe_host_rcp(int(np.log2(c.rcp_table().size))))
    yield c
    c.close()


@pytest.fixture(scope="module")
def lut(ctx):
    t = ctx.lut()
    assert np.array_equal(t, po.build_lut())
    return t


class Bundle:
    """Prepared scene + its baked batches (host bake through the C ABI)."""

    def __init__(self, name):
        self.ps = wl.load_scene(name) if name == "city" or wl.have_scene(name) else None
        if self.ps is None:
            pytest.skip(f"prepared scene {name} not available")
        baked = [api.bake(b, self.ps.ref_min, self.ps.ref_max) for b in self.ps.batches]
        self.packed = [b[0] for b in baked]
        self.centers = np.stack([b[1] for b in baked])
        self.bmin = np.stack([b[2] for b in baked])
        self.bmax = np.stack([b[3] for b in baked])
        self.boxes = self.ps.quad_boxes()

    def scene(self, ctx, boxes=None):
        return api.Scene(ctx, self.packed, self.ps.ref_min, self.ps.ref_max, self.bmin, self.bmax, self.centers,
                         self.boxes if boxes is None else boxes)

    def default_view(self, w, h):
        c = self.ps.camera
        return cam.view_projection(c["pos"], c["dir"], c["up"], c["fov"], w, h), np.array(c["pos"], np.float32)


_bundles = {}


def bundle(name):
    if name not in _bundles:
        _bundles[name] = Bundle(name)
    return _bundles[name]


def port_frame(B, port, mvp, order, boxes):
    gate, quads = port.frame(B.packed, B.bmin, B.bmax, B.ps.ref_min, B.ps.ref_max, mvp, order)
    q = port.query_boxes(boxes) if boxes is not None and len(boxes) else np.zeros(0, np.uint8)
    return gate, quads, port.depth(), port.hiz(), q


def _same_record(a, b) -> bool:
    """Bitwise equality, except that any NaN equals any NaN: a zero-length edge (v3 == v0 in the
    lone-triangle quads) gives 0 * inf, for which x86 produces 0xFFC00000 and the GPU 0x7FFFFFFF;
    no consumer can tell them apart (comparisons false, cvtt -> 0x80000000 for both)."""
    wa = np.frombuffer(bytes(a), np.uint32).copy()
    wb = np.frombuffer(bytes(b), np.uint32).copy()
    fa, fb = wa[6:21].view(np.float32), wb[6:21].view(np.float32)
    both_nan = np.isnan(fa) & np.isnan(fb)
    wa[6:21][both_nan] = 0
    wb[6:21][both_nan] = 0
    return bool(np.array_equal(wa, wb))


@pytest.mark.parametrize("name", ["city", "castle"])
def test_setup_records(ctx, name):
    B = bundle(name)
    w, h = 1920, 1080
    mvp, _ = B.default_view(w, h)
    r = api.Rasterizer(ctx, w, h)
    r.setModelViewProjection(mvp)
    port = po.PortRasterizer(w, h, np.zeros(4096, np.int64))
    port.set_mvp(mvp)
    n_valid = 0
    for bi in range(0, len(B.packed), max(1, len(B.packed) // 10)):
        occ = api.Occluder(ctx, B.packed[bi], B.ps.ref_min, B.ps.ref_max)
        for clipped in (False, True):
            recs = r.debug_setup(occ, clipped)
            pk = B.packed[bi]
            for q in range(len(recs)):
                g, l = divmod(q, 8)
                words = np.array([pk[(4 * g + j) * 8 + l] for j in range(4)], np.uint32)
                want = port.setup_quad(words, B.ps.ref_min, B.ps.ref_max, clipped)
                assert _same_record(recs[q], want), (name, bi, q, clipped)
                n_valid += int(want.mode != 0)
        occ.close()
    assert n_valid > 100
    r.close(); port.close()


@pytest.mark.parametrize("name,size", [("city", (640, 360)), ("castle", (1920, 1080))])
def test_per_call_api_frame(ctx, lut, name, size):
    """The reference's own call sequence (Main.cpp:181-206) through the drop-in per-call API."""
    B = bundle(name)
    w, h = size
    mvp, pos = B.default_view(w, h)
    order = cam.front_to_back_order(B.centers, pos)
    port = po.PortRasterizer(w, h, lut)
    boxes = B.boxes[:: max(1, len(B.boxes) // 4000)]
    gate, quads, depth, hiz, qv = port_frame(B, port, mvp, order, boxes)

    occs = [api.Occluder(ctx, p, B.ps.ref_min, B.ps.ref_max) for p in B.packed]
    r = api.Rasterizer(ctx, w, h)
    r.clear()
    r.setModelViewProjection(mvp)
    for slot, o in enumerate(order):
        vis, clip = r.queryVisibility(B.bmin[o], B.bmax[o])
        assert (int(vis) | (int(clip) << 1)) == gate[slot], (slot, o)
        if vis:
            r.rasterize(occs[o], clip)
    d, hz = r.download()
    assert np.array_equal(hz, hiz)
    assert np.array_equal(d, depth)
    assert np.array_equal(r.query_boxes(boxes), qv)
    assert np.array_equal(r.readBackDepth(), port.readback())
    # query2D directly (Rasterizer.cpp:283-349)
    rng = np.random.default_rng(3)
    for _ in range(40):
        x0, x1 = sorted(rng.integers(0, w, 2)); y0, y1 = sorted(rng.integers(0, h, 2)); z = int(rng.integers(0, 65536))
        assert r.query2D(int(x0), int(x1), int(y0), int(y1), z) == port.query2d(int(x0), int(x1), int(y0), int(y1), z)
    for o in occs:
        o.close()
    r.close(); port.close()


@pytest.mark.parametrize("name,size,tile_h", [("castle", (1920, 1080), 1), ("castle", (1920, 1080), 4), ("castle", (512, 256), 1), ("city", (640, 360), 1)])
def test_per_call_api_frames_with_predicted_chains(ctx, lut, name, size, tile_h):
    """Frame after frame through the per-call API: from the second frame on every launch also answers the queries the
    previous frame's sequence predicts (orz_percall_kernels.cuh).  Right predictions (same camera again), partly wrong
    ones (the camera moves: other order, other gate decisions), wrong ones (reversed order, queries without rasterize,
    repeated queries) -- every answer and every buffer must still be the reference's."""
    B = bundle(name)
    w, h = size
    mvps, poss = wl.camera_path(B.ps, 5, w, h)
    m0, p0 = B.default_view(w, h)
    views = [(m0, p0), (m0, p0), (mvps[1], poss[1]), (mvps[4], poss[4]), (m0, p0)]
    port = po.PortRasterizer(w, h, lut)
    occs = [api.Occluder(ctx, p, B.ps.ref_min, B.ps.ref_max) for p in B.packed]
    r = api.Rasterizer(ctx, w, h)
    ctx.set_tile_height(0, tile_h)
    for f, (mvp, pos) in enumerate(views):
        order = cam.front_to_back_order(B.centers, pos)
        if f == 3:
            order = order[::-1].copy()  # back to front: nothing the last frame predicts comes in that order
        r.clear(); port.clear()
        r.setModelViewProjection(mvp); port.set_mvp(mvp)
        n_visible = 0
        for slot, o in enumerate(order):
            g = port.query(B.bmin[o], B.bmax[o])
            vis, clip = r.queryVisibility(B.bmin[o], B.bmax[o])
            assert (int(vis) | (int(clip) << 1)) == g, (f, slot, o)
            if f == 2 and slot % 5 == 0:  # something unrelated in between, then the same box again
                o2 = order[(slot * 7 + 3) % len(order)]
                assert r.queryVisibility(B.bmin[o2], B.bmax[o2])[0] == bool(port.query(B.bmin[o2], B.bmax[o2]) & 1), (f, slot, o2)
                assert r.queryVisibility(B.bmin[o], B.bmax[o]) == (vis, clip)
            if vis:
                n_visible += 1
                r.rasterize(occs[o], clip)
                port.rasterize(B.packed[o], B.ps.ref_min, B.ps.ref_max, clip)
        assert n_visible > 2
        d, hz = r.download()
        assert np.array_equal(hz, port.hiz()), f
        assert np.array_equal(d, port.depth()), f
    # queries only, no rasterize in between (the chain must not stop answering after a visible one)
    r.clear()
    port.clear(); port.set_mvp(m0)
    r.setModelViewProjection(m0)
    for o in range(len(occs)):
        assert r.queryVisibility(B.bmin[o], B.bmax[o])[0] == bool(port.query(B.bmin[o], B.bmax[o]) & 1)
    ctx.set_tile_height(0, 1)
    for o in occs:
        o.close()
    r.close(); port.close()


def test_per_call_two_rasterizers_interleaved(ctx, lut):
    """Two Rasterizer objects of one context (one stream, one mailbox) driven in lock step, three frames each: their
    predicted chains, tags and mailbox words must not get into each other's way."""
    B = bundle("castle")
    sizes = [(1920, 1080), (512, 256)]
    m0, p0 = B.default_view(*sizes[0])
    m1, p1 = B.default_view(*sizes[1])
    order = cam.front_to_back_order(B.centers, p0)
    ports = [po.PortRasterizer(w, h, lut) for w, h in sizes]
    rs = [api.Rasterizer(ctx, w, h) for w, h in sizes]
    occs = [api.Occluder(ctx, p, B.ps.ref_min, B.ps.ref_max) for p in B.packed]
    for f in range(3):
        for r, port, m in zip(rs, ports, (m0, m1)):
            r.clear(); port.clear()
            r.setModelViewProjection(m); port.set_mvp(m)
        for slot, o in enumerate(order):
            for k, (r, port) in enumerate(zip(rs, ports)):
                g = port.query(B.bmin[o], B.bmax[o])
                vis, clip = r.queryVisibility(B.bmin[o], B.bmax[o])
                assert (int(vis) | (int(clip) << 1)) == g, (f, slot, o, k)
                if vis:
                    r.rasterize(occs[o], clip)
                    port.rasterize(B.packed[o], B.ps.ref_min, B.ps.ref_max, clip)
        for r, port in zip(rs, ports):
            d, hz = r.download()
            assert np.array_equal(hz, port.hiz()) and np.array_equal(d, port.depth()), f
    for o in occs:
        o.close()
    for r, port in zip(rs, ports):
        r.close(); port.close()


@pytest.mark.parametrize("trav", [1, 2])
@pytest.mark.parametrize("gw", [1, 2, 4, 8])
@pytest.mark.parametrize("name,size,nviews", [("city", (640, 360), 6), ("castle", (1920, 1080), 5), ("castle", (512, 256), 12)])
def test_view_batch(ctx, lut, name, size, nviews, gw, trav):
    """Both traversal mappings (1 = warp per block, 2 = lane per block) and every group size."""
    B = bundle(name)
    w, h = size
    ctx.set_group_warps(gw)
    ctx.set_traversal(trav)
    ctx.set_cluster_views(0)  # the one-CTA-per-view batch kernel (few views would take the cluster path)
    mvps, poss = wl.camera_path(B.ps, nviews - 1, w, h) if size[0] >= 640 else wl.probe_views(B.ps, nviews - 1, w, h)
    m0, p0 = B.default_view(w, h)
    mvps = np.concatenate([m0[None], mvps]); poss = np.concatenate([p0[None], poss])
    orders = wl.orders_for(B.centers, poss)
    boxes = B.boxes[:: max(1, len(B.boxes) // 3000)]
    sc = B.scene(ctx, boxes)
    out = sc.render_views(w, h, mvps, orders=orders, want=("vis", "clip", "gate", "depth", "hiz", "quads"))
    vis = api.unpack_bits(out["vis"], len(boxes)); clip = api.unpack_bits(out["clip"], len(boxes))
    port = po.PortRasterizer(w, h, lut)
    for v in range(nviews):
        gate, quads, depth, hiz, qv = port_frame(B, port, mvps[v], orders[v], boxes)
        assert np.array_equal(out["gate"][v], gate), v
        assert out["quads"][v] == quads
        assert np.array_equal(out["hiz"][v], hiz), v
        assert np.array_equal(out["depth"][v], depth), v
        assert np.array_equal(vis[v], (qv & 1).astype(bool)), v
        assert np.array_equal(clip[v], (qv & 2).astype(bool)), v
    # bits-only call (scratch depth per CTA) and GPU-side ordering must give the same bits
    out2 = sc.render_views(w, h, mvps, cam_pos=poss, want=("vis", "gate"))
    assert np.array_equal(out2["vis"], out["vis"])
    assert np.array_equal(out2["gate"], out["gate"])
    ctx.set_group_warps(0)
    ctx.set_traversal(2)
    ctx.set_cluster_views(api.DEFAULT_CLUSTER_VIEWS)
    sc.close(); port.close()


@pytest.mark.parametrize("name,size,nviews", [("city", (640, 360), 6), ("castle", (1920, 1080), 5), ("castle", (512, 256), 12),
                                              ("castle", (1280, 720), 3), ("castle", (2560, 1440), 2), ("sponza", (1920, 1080), 2),
                                              ("castle", (3840, 2160), 2)])  # 3840x2160: 16-bit index wrap -> two records per wrapped primitive
def test_view_cluster_path(ctx, lut, name, size, nviews):
    """Latency path for few views: speculative setup kernel + one thread-block cluster (2-16 CTAs) per
    view run as a dataflow machine (tile-local gates, DSMEM decision flags, k_raster_views_cluster)."""
    B = bundle(name)
    w, h = size
    ctx.set_cluster_views(api.DEFAULT_CLUSTER_VIEWS)
    mvps, poss = wl.camera_path(B.ps, nviews - 1, w, h) if size[0] >= 640 else wl.probe_views(B.ps, nviews - 1, w, h)
    m0, p0 = B.default_view(w, h)
    mvps = np.concatenate([m0[None], mvps]); poss = np.concatenate([p0[None], poss])
    orders = wl.orders_for(B.centers, poss)
    boxes = B.boxes[:: max(1, len(B.boxes) // 3000)]
    sc = B.scene(ctx, boxes)
    n0 = ctx.launch_count
    out = sc.render_views(w, h, mvps, orders=orders, want=("vis", "clip", "gate", "depth", "hiz", "quads"))
    assert ctx.launch_count - n0 == 4  # prepare, speculative setup, cluster raster, queries
    vis = api.unpack_bits(out["vis"], len(boxes)); clip = api.unpack_bits(out["clip"], len(boxes))
    port = po.PortRasterizer(w, h, lut)
    for v in range(nviews):
        gate, quads, depth, hiz, qv = port_frame(B, port, mvps[v], orders[v], boxes)
        assert np.array_equal(out["gate"][v], gate), v
        assert out["quads"][v] == quads
        assert np.array_equal(out["hiz"][v], hiz), v
        assert np.array_equal(out["depth"][v], depth), v
        assert np.array_equal(vis[v], (qv & 1).astype(bool)), v
        assert np.array_equal(clip[v], (qv & 2).astype(bool)), v
    out2 = sc.render_views(w, h, mvps, cam_pos=poss, want=("vis", "gate"))
    assert np.array_equal(out2["vis"], out["vis"])
    assert np.array_equal(out2["gate"], out["gate"])
    # and the batch kernel gives the same answer
    ctx.set_cluster_views(0)
    out3 = sc.render_views(w, h, mvps, orders=orders, want=("vis", "gate", "depth", "hiz"))
    ctx.set_cluster_views(api.DEFAULT_CLUSTER_VIEWS)
    for k in ("vis", "gate", "depth", "hiz"):
        assert np.array_equal(out3[k], out[k]), k
    sc.close(); port.close()


@pytest.mark.parametrize("cluster", [0, 16])
@pytest.mark.parametrize("size", [(1280, 720), (3840, 2160)])
def test_no_gate_forced_clip_and_4k_wrap(ctx, lut, size, cluster):
    """Config-4 shape: every batch through rasterize<true>, no gate; 3840x2160 also exercises the
    16-bit wrap of the first-block index (Rasterizer.cpp:1054)."""
    B = bundle("castle" if wl.have_scene("castle") else "city")
    w, h = size
    ctx.set_cluster_views(cluster)
    mvps, poss = wl.camera_path(B.ps, 2, w, h)
    orders = wl.orders_for(B.centers, poss)
    sc = B.scene(ctx, B.boxes[::7])
    out = sc.render_views(w, h, mvps, orders=orders, flags=api.BATCH_NO_GATE | api.BATCH_FORCE_CLIPPED, want=("vis", "depth", "hiz"))
    ctx.set_cluster_views(api.DEFAULT_CLUSTER_VIEWS)
    vis = api.unpack_bits(out["vis"], len(B.boxes[::7]))
    port = po.PortRasterizer(w, h, lut)
    for v in range(2):
        port.clear(); port.set_mvp(mvps[v])
        for o in orders[v]:
            port.rasterize(B.packed[o], B.ps.ref_min, B.ps.ref_max, True)
        assert np.array_equal(out["hiz"][v], port.hiz())
        assert np.array_equal(out["depth"][v], port.depth())
        assert np.array_equal(vis[v], (port.query_boxes(B.boxes[::7]) & 1).astype(bool))
    sc.close(); port.close()


@pytest.mark.parametrize("cluster", [0, 16])
def test_soup_near_clipped(ctx, lut, cluster):
    ctx.set_cluster_views(cluster)
    ps = wl.synthetic_soup(8192, cube=60.0)
    baked = [api.bake(b, ps.ref_min, ps.ref_max) for b in ps.batches]
    sc = api.Scene(ctx, [b[0] for b in baked], ps.ref_min, ps.ref_max, np.stack([b[2] for b in baked]), np.stack([b[3] for b in baked]),
                   np.stack([b[1] for b in baked]), ps.quad_boxes()[::5])
    w, h = 1280, 720
    c = ps.camera
    mvps = np.stack([cam.view_projection(c["pos"], d, c["up"], c["fov"], w, h) for d in ((0, 0, 1), (1, 0.2, 0.1), (-0.3, 0.1, -1))])
    poss = np.zeros((3, 3), np.float32)
    orders = wl.orders_for(np.stack([b[1] for b in baked]), poss)
    port = po.PortRasterizer(w, h, lut)
    for flags in (0, api.BATCH_NO_GATE | api.BATCH_FORCE_CLIPPED):
        out = sc.render_views(w, h, mvps, orders=orders, flags=flags, want=("vis", "depth", "hiz", "gate"))
        for v in range(3):
            if flags:
                port.clear(); port.set_mvp(mvps[v])
                for o in orders[v]:
                    port.rasterize(baked[o][0], ps.ref_min, ps.ref_max, True)
            else:
                gate, _ = port.frame([b[0] for b in baked], np.stack([b[2] for b in baked]), np.stack([b[3] for b in baked]),
                                     ps.ref_min, ps.ref_max, mvps[v], orders[v])
                assert np.array_equal(out["gate"][v], gate)
            assert np.array_equal(out["hiz"][v], port.hiz())
            assert np.array_equal(out["depth"][v], port.depth())
    ctx.set_cluster_views(api.DEFAULT_CLUSTER_VIEWS)
    sc.close(); port.close()


@pytest.mark.skipif(not (ro.available() and ro.scene_available("Castle")), reason="reference build / scene data not shipped")
def test_against_reference_binary_on_this_host(ctx):
    """The real reference (unmodified sources, oracle/_ref) run on this box's CPU vs the GPU."""
    B = bundle("castle")
    s = ro.RefScene.load("Castle")
    assert all(np.array_equal(s.packed(i), B.packed[i]) for i in range(s.n_occluders))
    w, h = 1920, 1080
    mvps, poss = wl.camera_path(B.ps, 4, w, h)
    orders = wl.orders_for(B.centers, poss)
    sc = B.scene(ctx)
    out = sc.render_views(w, h, mvps, orders=orders, want=("vis", "clip", "gate", "depth", "hiz"))
    vis = api.unpack_bits(out["vis"], len(B.boxes)); clip = api.unpack_bits(out["clip"], len(B.boxes))
    r = ro.RefRasterizer(w, h)
    for v in range(4):
        gate, _ = r.frame(s, mvps[v], orders[v])
        assert np.array_equal(out["gate"][v], gate)
        assert np.array_equal(out["hiz"][v], r.hiz())
        assert np.array_equal(out["depth"][v], r.depth())
        q = r.query_boxes(B.boxes)
        assert np.array_equal(vis[v], (q & 1).astype(bool))
        assert np.array_equal(clip[v], (q & 2).astype(bool))
    r.close(); s.close(); sc.close()


@pytest.mark.parametrize("name,size", [("castle", (1920, 1080)), ("castle", (3840, 2160)), ("soup", (3840, 2160))])
def test_wide_single_view_path(ctx, lut, name, size):
    """One view split over the whole GPU (config-4 path: ungated, rasterize<true>, one warp per
    screen block-row walking globally set-up records) against the oracle."""
    w, h = size
    if name == "soup":
        ps = wl.synthetic_soup(32768, cube=80.0)
        baked = [api.bake(b, ps.ref_min, ps.ref_max) for b in ps.batches]
        packed, centers = [b[0] for b in baked], np.stack([b[1] for b in baked])
        sc = api.Scene(ctx, packed, ps.ref_min, ps.ref_max, np.stack([b[2] for b in baked]), np.stack([b[3] for b in baked]), centers, ps.quad_boxes()[::9])
        boxes, ref_min, ref_max = ps.quad_boxes()[::9], ps.ref_min, ps.ref_max
        c = ps.camera
        mvps = np.stack([cam.view_projection(c["pos"], d, c["up"], c["fov"], w, h) for d in ((0, 0, 1), (0.6, -0.2, 0.7))])
        poss = np.zeros((2, 3), np.float32)
    else:
        B = bundle(name)
        packed, centers, boxes, ref_min, ref_max = B.packed, B.centers, B.boxes[::7], B.ps.ref_min, B.ps.ref_max
        sc = B.scene(ctx, boxes)
        mvps, poss = wl.camera_path(B.ps, 2, w, h)
    orders = wl.orders_for(centers, poss)
    flags = api.BATCH_NO_GATE | api.BATCH_FORCE_CLIPPED | api.BATCH_WIDE
    out = sc.render_views(w, h, mvps, orders=orders, flags=flags, want=("vis", "depth", "hiz", "quads", "gate"))
    out2 = sc.render_views(w, h, mvps, cam_pos=poss, flags=flags, want=("vis",))
    vis = api.unpack_bits(out["vis"], len(boxes))
    port = po.PortRasterizer(w, h, lut)
    for v in range(2):
        port.clear(); port.set_mvp(mvps[v])
        for o in orders[v]:
            port.rasterize(packed[o], ref_min, ref_max, True)
        assert np.array_equal(out["hiz"][v], port.hiz())
        assert np.array_equal(out["depth"][v], port.depth())
        assert np.array_equal(vis[v], (port.query_boxes(boxes) & 1).astype(bool))
        assert out["quads"][v] == sum(p.size // 4 for p in packed) and out["gate"][v].all()
    assert np.array_equal(out2["vis"], out["vis"])
    sc.close(); port.close()
