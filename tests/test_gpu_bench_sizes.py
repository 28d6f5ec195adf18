"""Parity at the BENCHMARKED sizes and code variants ("what's weak" 1): the exact batches bench.py
times -- Castle 1 024 views at 1920x1080 (k_raster_views_cluster<2>, k_sort_views, four cost-sorted sub-batches on
auxiliary streams), 1 024 / 8 192 probes at 512x256 (single-CTA clusters / the large-batch kernel with four groups),
Sponza 256 views, config 4 at its stated size (5 M quads, 3840x2160, rasterize<true>) -- rendered through the C ABI
and compared, EVERY view, bit for bit (gate bytes, HiZ, depth, visibility and needsClipping bits, quads submitted)
with the unmodified reference (oracle/_ref) running on all host threads of this box."""
import numpy as np
import pytest

from oracle import ref_oracle as ro
from rasterizer_b200 import api
from rasterizer_b200 import workloads as wl

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ro.available(), reason="reference build (oracle/_ref) not shipped")]


@pytest.fixture(scope="module")
def ctx():
    c = api.Context(0)
    yield c
    c.close()


class Case:
    """Prepared scene baked by the product (host bake) + the same batches baked by the reference."""

    def __init__(self, name):
        if name != "city" and not wl.have_scene(name):
            pytest.skip(f"prepared scene {name} not available")
        self.ps = wl.load_scene(name)
        self.ref = ro.RefScene.from_batches(self.ps.batches, self.ps.ref_min, self.ps.ref_max)
        self.boxes = self.ps.quad_boxes()

    def close(self):
        self.ref.close()


_cases = {}


def case(name):
    if name not in _cases:
        _cases[name] = Case(name)
    return _cases[name]


def check(cs, w, h, mvps, poss, out, boxes, mode=0):
    """Every view of `out` against the reference; host-side order (Main.cpp:185-190) for the reference."""
    orders = wl.orders_for(cs.ref.centers, poss)
    mism = ro.check_views(cs.ref, w, h, mvps, orders, boxes, mode=mode, gate=out.get("gate"), depth=out.get("depth"), hiz=out.get("hiz"),
                          vis=out.get("vis"), clip=out.get("clip"), quads=out.get("quads"))
    assert not mism.any(), ro.describe_mismatch(mism)


ALL = ("vis", "clip", "gate", "depth", "hiz", "quads")


def test_castle_1080p_1024_views_default_path(ctx):
    """BASELINE config 3 exactly as bench.py runs it: camera positions in, order computed on the GPU, the automatic
    cluster size (2 CTAs per view), longest-first view order, four sub-batches on auxiliary streams."""
    cs = case("castle")
    w, h, n = 1920, 1080, 1024
    mvps, poss = wl.camera_path(cs.ps, n, w, h)
    sc = api.Scene.from_prepared(ctx, cs.ps)
    n0 = ctx.launch_count
    out = sc.render_views(w, h, mvps, cam_pos=poss, want=ALL)
    assert ctx.launch_count - n0 == 2 + 3 * 4  # prepare, sort, 4 x (setup + cluster raster + queries): the benchmarked launch chain
    check(cs, w, h, mvps, poss, out, cs.boxes)
    # the bits-only call of the bench (internal depth arena) must give the same bits
    out2 = sc.render_views(w, h, mvps, cam_pos=poss, want=("vis", "gate"))
    assert np.array_equal(out2["vis"], out["vis"]) and np.array_equal(out2["gate"], out["gate"])
    sc.close()


@pytest.mark.parametrize("n,cluster_views", [(1024, None), (8192, None), (8192, 1024)])
def test_castle_512x256_probes(ctx, n, cluster_views):
    """BASELINE config 5: the per-GPU slice of an 8-GPU run (1 024 probes) and the whole batch on one GPU (8 192: the cluster
    path in three chunks of views, and -- cluster_views 1024 -- the large-batch kernel with its four sub-batches)."""
    cs = case("castle")
    w, h = 512, 256
    mvps, poss = wl.probe_views(cs.ps, n, w, h)
    sc = api.Scene.from_prepared(ctx, cs.ps)
    if cluster_views is not None:
        ctx.set_cluster_views(cluster_views)
    try:
        out = sc.render_views(w, h, mvps, cam_pos=poss, want=ALL)
    finally:
        ctx.set_cluster_views(api.DEFAULT_CLUSTER_VIEWS)
    check(cs, w, h, mvps, poss, out, cs.boxes)
    out2 = sc.render_views(w, h, mvps, cam_pos=poss, want=("vis",))  # what the bench asks for: bits only
    assert np.array_equal(out2["vis"], out["vis"])
    # every second probe = what rank 0 of a 2-GPU run renders (views dealt round-robin)
    out3 = sc.render_views(w, h, mvps[0::2], cam_pos=poss[0::2], want=("vis",))
    assert np.array_equal(out3["vis"], out["vis"][0::2])
    sc.close()


def test_sponza_1080p_256_views(ctx):
    cs = case("sponza")
    w, h, n = 1920, 1080, 256
    mvps, poss = wl.camera_path(cs.ps, n, w, h)
    sc = api.Scene.from_prepared(ctx, cs.ps)
    out = sc.render_views(w, h, mvps, cam_pos=poss, want=ALL)
    check(cs, w, h, mvps, poss, out, cs.boxes)
    sc.close()


@pytest.mark.parametrize("size,csize,n", [((512, 256), 1, 40), ((512, 256), 2, 7), ((1920, 1080), 2, 24), ((1920, 1080), 4, 9),
                                          ((1920, 1080), 8, 5), ((1920, 1080), 16, 3), ((1280, 720), 1, 33)])
def test_explicit_cluster_sizes(ctx, size, csize, n):
    """Every cluster size the automatic choice can make, forced on few views (incl. view counts that are not a
    multiple of anything) -- the automatic choice alone only reaches 4, 8 and 16 on small batches."""
    cs = case("castle")
    w, h = size
    mvps, poss = wl.camera_path(cs.ps, n, w, h)
    boxes = cs.boxes[::3]
    sc = api.Scene.from_prepared(ctx, cs.ps, boxes=boxes)
    ctx.set_cluster_size(csize)
    try:
        out = sc.render_views(w, h, mvps, cam_pos=poss, want=ALL)
    finally:
        ctx.set_cluster_size(0)
    check(cs, w, h, mvps, poss, out, boxes)
    sc.close()


@pytest.mark.parametrize("scene,size,csize,tile_h,n", [("castle", (1920, 1080), 16, 1, 3), ("castle", (1920, 1080), 8, 1, 5), ("castle", (1920, 1080), 16, 4, 2),
                                                        ("castle", (512, 256), 16, 1, 2), ("castle", (512, 256), 4, 1, 9), ("castle", (512, 256), 2, 1, 12),
                                                        ("castle", (1280, 720), 8, 1, 4), ("sponza", (1920, 1080), 16, 1, 2), ("castle", (1920, 1080), 0, 0, 1),
                                                        ("castle", (1920, 1080), 0, 0, 4), ("castle", (640, 360), 0, 1, 3)])
def test_tile_heights(ctx, scene, size, csize, tile_h, n):
    """8 x 1 strips (the few-view / latency shape of the cluster path) and 8 x 4 tiles, forced and automatic, incl. a target
    whose block rows are not a multiple of anything (45 rows) -- identical results for every shape."""
    if not wl.have_scene(scene):
        pytest.skip("prepared scene missing")
    cs = case(scene)
    w, h = size
    mvps, poss = wl.camera_path(cs.ps, n, w, h)
    boxes = cs.boxes[::3]
    sc = api.Scene.from_prepared(ctx, cs.ps, boxes=boxes)
    ctx.set_cluster_size(csize)
    ctx.set_tile_height(tile_h, 1)
    try:
        out = sc.render_views(w, h, mvps, cam_pos=poss, want=ALL)
    finally:
        ctx.set_cluster_size(0)
        ctx.set_tile_height(0, 1)
    check(cs, w, h, mvps, poss, out, boxes)
    sc.close()


def test_batch_kernel_four_groups_1080p(ctx):
    """The large-batch kernel's >= 64-view path (four sub-batches, atomic view counters) at 1080p."""
    cs = case("castle")
    w, h, n = 1920, 1080, 160
    mvps, poss = wl.camera_path(cs.ps, n, w, h)
    sc = api.Scene.from_prepared(ctx, cs.ps)
    ctx.set_cluster_views(0)
    try:
        out = sc.render_views(w, h, mvps, cam_pos=poss, want=ALL)
    finally:
        ctx.set_cluster_views(api.DEFAULT_CLUSTER_VIEWS)
    check(cs, w, h, mvps, poss, out, cs.boxes)
    sc.close()


@pytest.mark.parametrize("n_quads,size", [(131_072, (1280, 720)), (131_072, (3840, 2160)), (5_000_000, (3840, 2160))])
def test_config4_soup_near_clip(ctx, n_quads, size):
    """BASELINE config 4 (the last case at its stated size: 5 M quads = 10 M triangles, 3840x2160): camera inside the
    geometry, every batch through rasterize<true>, no gate -- one view at a time over the whole GPU, tile major
    (k_raster_tiles); 3840x2160 includes the 16-bit first-block index wrap (Rasterizer.cpp:1054)."""
    from rasterizer_b200 import camera as cam

    ps = wl.synthetic_soup(n_quads, cube=200.0 if n_quads > 1_000_000 else 60.0)
    w, h = size
    ref = ro.RefScene.from_batches(ps.batches, ps.ref_min, ps.ref_max)
    boxes = ps.quad_boxes()[::97]
    sc = api.Scene.bake_on_device(ctx, ps.batches, ps.ref_min, ps.ref_max, boxes)
    c = ps.camera
    dirs = ((0.0, 0.0, 1.0), (0.6, -0.2, 0.7))
    mvps = np.stack([cam.view_projection(c["pos"], d, c["up"], c["fov"], w, h) for d in dirs])
    poss = np.zeros((len(dirs), 3), np.float32)
    orders = wl.orders_for(ref.centers, poss)
    flags = api.BATCH_NO_GATE | api.BATCH_FORCE_CLIPPED
    n0 = ctx.launch_count
    out = sc.render_views(w, h, mvps, orders=orders, flags=flags, want=("vis", "clip", "depth", "hiz", "quads", "gate"))
    assert ctx.launch_count - n0 == 1 + 2 * len(dirs) + 1  # prepare, per view (setup + k_raster_tiles), queries
    mism = ro.check_views(ref, w, h, mvps, orders, boxes, mode=3, depth=out["depth"], hiz=out["hiz"], vis=out["vis"], clip=out["clip"],
                          quads=out["quads"])
    assert not mism.any(), ro.describe_mismatch(mism)
    assert out["gate"].all()
    # the order computed on the GPU gives the same frame
    out2 = sc.render_views(w, h, mvps, cam_pos=poss, flags=flags, want=("vis", "hiz"))
    assert np.array_equal(out2["vis"], out["vis"]) and np.array_equal(out2["hiz"], out["hiz"])
    sc.close(); ref.close()


def test_more_than_65535_views(ctx):
    """the occludee-query grid put the view index on grid.y (limit 65 535)."""
    cs = case("city")
    w, h, n = 64, 64, 66_000
    base_m, base_p = wl.camera_path(cs.ps, 500, w, h)
    idx = np.arange(n) % 500
    mvps, poss = base_m[idx], base_p[idx]
    boxes = cs.boxes[::5]
    sc = api.Scene.from_prepared(ctx, cs.ps, boxes=boxes)
    out = sc.render_views(w, h, mvps, cam_pos=poss, want=("vis", "clip", "gate"))
    first = {k: out[k][:500] for k in out}
    check(cs, w, h, base_m, base_p, first, boxes)
    for k in out:  # the batch repeats the 500 distinct views
        assert np.array_equal(out[k], first[k][idx]), k
    sc.close()
