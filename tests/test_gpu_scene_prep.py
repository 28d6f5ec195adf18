"""Device batching (orz_generate_batches_device, orz_sah_kernels.cuh) against the host batching of the
product and -- where oracle/_ref travelled -- the unmodified reference: same batches, same order."""
import numpy as np
import pytest

from oracle import ref_oracle as ro
from rasterizer_b200 import api
from rasterizer_b200 import workloads as wl
from prep_cases import boxes_case, same_batches, terrain

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = api.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("seed,n,target,gran,snap", [
    (1, 24, 512, 8, None),          # one level, one tile
    (3, 4000, 512, 8, None),
    (4, 4000, 512, 8, 10.0),        # many equal centres: tie order of the stable sorts
    (5, 20000, 512, 8, 5.0),
    (6, 5000, 64, 16, 4.0),
    (7, 3001, 100, 7, 2.0),
    (8, 6000, 16, 1, 25.0),         # hundreds of segments per level: two segment digits in the sort
    (9, 70000, 8, 2, 3.0),          # thousands of segments: three segment digits
    (10, 300000, 512, 8, None),     # 147 sort tiles, 1 172 scan chunks in the root segment
])
def test_device_batches_equal_host(ctx, seed, n, target, gran, snap):
    boxes = boxes_case(np.random.default_rng(seed), n, snap)
    got = ctx.generate_batches(boxes, target, gran)
    assert ctx.launch_count > 0
    assert same_batches(got, api.generate_batches(boxes, target, gran))
    if ro.available() and n <= 20000:
        assert same_batches(got, ro.generate_batches(boxes, target, gran))


@pytest.mark.parametrize("name", ["castle", "sponza"])
def test_device_batches_of_the_reference_scenes(ctx, name):
    if not wl.have_scene(name):
        pytest.skip("prepared scene missing")
    boxes = wl.load_scene(name).quad_boxes()  # the quads' AABBs, in batch order instead of mesh order
    got = ctx.generate_batches(boxes, 512, 8)
    assert same_batches(got, api.generate_batches(boxes, 512, 8))
    assert len(got) >= 70 and all(len(b) % 8 == 0 and len(b) < 512 for b in got)


def test_device_batches_signed_zero_and_errors(ctx):
    boxes = boxes_case(np.random.default_rng(21), 3000, 8.0)
    boxes[::3, 1] = boxes[::3, 5] = 0.0
    boxes[1::6, 1] = boxes[1::6, 5] = -0.0
    boxes[::5, 0] = -boxes[::5, 4]
    assert same_batches(ctx.generate_batches(boxes, 256, 8), api.generate_batches(boxes, 256, 8))
    with pytest.raises(api.OrzError, match="no split position"):
        ctx.generate_batches(boxes[:16], 512, 8)


def test_prepare_mesh_on_device_feeds_the_renderer(ctx):
    """Mesh in, visibility out, nothing from the reference in between: terrain -> quads -> device batches
    -> device bake -> one view; equal to the same scene prepared on the host."""
    from rasterizer_b200 import camera as cam

    idx, verts = terrain(np.random.default_rng(5), 4000, 3.0)
    api.set_rsqrt_table(None)
    host = wl.prepare_mesh("terrain", idx, verts, {}, 128, 8)
    dev = wl.prepare_mesh("terrain", idx, verts, {}, 128, 8, generate_batches=ctx.generate_batches)
    assert len(host.batches) == len(dev.batches) > 8
    assert all(np.array_equal(a.view(np.uint32), b.view(np.uint32)) for a, b in zip(host.batches, dev.batches))
    scene = api.Scene.bake_on_device(ctx, dev.batches, dev.ref_min, dev.ref_max, dev.quad_boxes())
    scene_host = api.Scene.from_prepared(ctx, host)
    w, h = 640, 360
    pos = np.array([20.0, 15.0, -10.0], np.float32)
    mvp = cam.view_projection(pos, np.array([0.0, -0.4, 0.9], np.float32), np.array([0.0, 1.0, 0.0], np.float32), 0.9, w, h)
    out = scene.render_views(w, h, mvp[None], cam_pos=pos[None], want=("vis",))
    vis = api.unpack_bits(out["vis"], dev.n_quads)[0]
    assert 0 < vis.sum() < dev.n_quads
    assert np.array_equal(out["vis"], scene_host.render_views(w, h, mvp[None], cam_pos=pos[None], want=("vis",))["vis"])
    scene.close(); scene_host.close()


def test_scene_from_mesh_in_one_call(ctx):
    """orz_scene_from_mesh (decompose on the host, batches + bake on the GPU) against the step-by-step host
    preparation: same occluders, and the same frame."""
    from rasterizer_b200 import camera as cam

    api.set_rsqrt_table(None)
    cases = [("terrain", *terrain(np.random.default_rng(9), 6000, 2.0), dict(pos=(20.0, 15.0, -10.0), dir=(0.0, -0.4, 0.9), up=(0.0, 1.0, 0.0), fov=0.9), (640, 360))]
    if ro.available() and ro.scene_available("Castle"):
        cases.append(("castle", *ro.load_mesh("Castle"), cam.CASTLE_CAMERA, (1920, 1080)))
    for name, idx, verts, c, (w, h) in cases:
        host = wl.prepare_mesh(name, idx, verts, {})
        want = api.Scene.from_prepared(ctx, host)
        got = api.Scene.from_mesh(ctx, idx, verts)
        assert got.n_occluders == want.n_occluders and got.n_boxes == want.n_boxes == host.n_quads
        for a, b in ((got.centers, want.centers), (got.bounds_min, want.bounds_min), (got.bounds_max, want.bounds_max)):
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
        assert np.array_equal(got.quads_per_occluder, [b.shape[0] // 4 for b in host.batches])
        pos = np.asarray(c["pos"], np.float32)
        mvp = cam.view_projection(pos, np.asarray(c["dir"], np.float32), np.asarray(c["up"], np.float32), c["fov"], w, h)
        outs = [s.render_views(w, h, mvp[None], cam_pos=pos[None], want=("vis", "gate", "hiz", "depth", "quads")) for s in (got, want)]
        for key in ("vis", "gate", "hiz", "depth", "quads"):
            assert np.array_equal(outs[0][key], outs[1][key]), (name, key)
        assert outs[0]["quads"][0] > 0
        got.close(); want.close()
    with pytest.raises(api.OrzError, match="multiple of 8"):
        api.Scene.from_mesh(ctx, cases[0][1], cases[0][2], 512, 4)


def test_baked_scene_file_and_mesh_files(ctx, tmp_path):
    """orz_scene_save / orz_scene_load (the HBM layout as a file) and orz_scene_from_mesh_files (the reference's raw
    index / vertex files, Main.cpp:56-84): every way into HBM renders the same frame."""
    from rasterizer_b200 import camera as cam

    api.set_rsqrt_table(None)
    idx, verts = terrain(np.random.default_rng(3), 5000, 2.5)
    idx.astype(np.uint32).tofile(tmp_path / "IndexBuffer.bin")
    verts.astype(np.float32).tofile(tmp_path / "VertexBuffer.bin")
    first = api.Scene.from_mesh(ctx, idx, verts, 256, 8)
    first.save(str(tmp_path / "terrain.orzbake"))
    scenes = [first, api.Scene.load(ctx, str(tmp_path / "terrain.orzbake")),
              api.Scene.from_mesh_files(ctx, str(tmp_path / "IndexBuffer.bin"), str(tmp_path / "VertexBuffer.bin"), 256, 8)]
    w, h = 512, 256
    pos = np.array([18.0, 12.0, -8.0], np.float32)
    mvp = cam.view_projection(pos, np.array([0.1, -0.35, 0.9], np.float32), np.array([0.0, 1.0, 0.0], np.float32), 0.9, w, h)
    outs = [s.render_views(w, h, mvp[None], cam_pos=pos[None], want=("vis", "gate", "hiz", "quads")) for s in scenes]
    for s, o in zip(scenes[1:], outs[1:]):
        assert s.n_occluders == first.n_occluders and s.n_boxes == first.n_boxes == first.n_quads
        assert np.array_equal(s.centers.view(np.uint32), first.centers.view(np.uint32))
        for key in ("vis", "gate", "hiz", "depth", "quads"):
            assert np.array_equal(o[key], outs[0][key]), key
    assert outs[0]["quads"][0] > 0
    bad = tmp_path / "bad.orzbake"
    bad.write_bytes((tmp_path / "terrain.orzbake").read_bytes()[:-16])
    with pytest.raises(api.OrzError, match="truncated"):
        api.Scene.load(ctx, str(bad))
    with pytest.raises(api.OrzError, match="cannot read"):
        api.Scene.from_mesh_files(ctx, str(tmp_path / "nope.bin"), str(tmp_path / "VertexBuffer.bin"))
    for s in scenes:
        s.close()
