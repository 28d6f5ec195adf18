"""Scene preparation: orz_quad_decompose / orz_generate_batches against the
UNMODIFIED reference (QuadDecomposition.cpp, SurfaceAreaHeuristic.cpp compiled into oracle/_ref):
same quads in the same order, same batches in the same order.  CPU only; the device batching is
checked in tests/test_gpu_scene_prep.py."""
import numpy as np
import pytest

from oracle import ref_oracle as ro
from rasterizer_b200 import api
from rasterizer_b200 import workloads as wl
from prep_cases import boxes_case, same_batches, terrain

pytestmark = pytest.mark.skipif(not ro.available(), reason="oracle/_ref/libref_oracle.so not built (needs /root/reference)")


@pytest.fixture(scope="module", autouse=True)
def host_rsqrt():
    api.set_rsqrt_table(None)  # canMergeTrianglesToQuad normalises with this host's rsqrtps, as the reference build does
    yield


@pytest.mark.parametrize("name", ["Castle", "Sponza"])
def test_reference_scenes(name):
    """Main.cpp:86-107 on the reference's own meshes: decompose, pad, per-quad AABBs, batches."""
    if not ro.scene_available(name):
        pytest.skip(f"no {name} data under oracle/_ref/scenes")
    idx, verts = ro.load_mesh(name)
    quads = api.quad_decompose(idx, verts)
    assert np.array_equal(quads, ro.quad_decompose(idx, verts))
    boxes = wl.quad_aabbs(wl.pad_quads(quads), verts)
    assert same_batches(api.generate_batches(boxes, 512, 8), ro.generate_batches(boxes, 512, 8))


@pytest.mark.parametrize("seed,n,bump", [(1, 60, 0.0), (2, 400, 0.0), (3, 400, 6.0), (4, 3000, 2.0), (5, 3000, 30.0), (6, 9000, 0.5)])
def test_quad_decomposeterrain(seed, n, bump):
    idx, verts = terrain(np.random.default_rng(seed), n, bump)
    got, want = api.quad_decompose(idx, verts), ro.quad_decompose(idx, verts)
    assert np.array_equal(got, want)
    pairs = int(np.sum(got.reshape(-1, 4)[:, 0] != got.reshape(-1, 4)[:, 3]))
    if bump == 0.0:
        assert pairs > 0.45 * (len(idx) // 3) / 1  # flat sheet: nearly every triangle finds a partner


def test_quad_decompose_awkward_meshes():
    rng = np.random.default_rng(11)
    idx, verts = terrain(rng, 300, 1.0)
    tris = idx.reshape(-1, 3)
    cases = {
        "empty": np.zeros(0, np.uint32),
        "one": tris[:1].reshape(-1),
        "duplicates": np.concatenate([tris, tris[::3]]).reshape(-1),                 # several owners per directed edge
        "two-sided": np.concatenate([tris, tris[::2][:, [0, 2, 1]]]).reshape(-1),     # back faces share every edge reversed
        "needles": np.concatenate([tris[:50], np.stack([tris[:20, 0], tris[:20, 1], tris[:20, 0]], 1), tris[50:]]).reshape(-1),  # (a, b, a)
        "points": np.concatenate([tris[:50], np.stack([tris[:5, 0]] * 3, 1), tris[50:]]).reshape(-1),                            # (a, a, a)
        "trailing": np.concatenate([idx, idx[:2]]),                                    # nIndices % 3 != 0: tail ignored
    }
    for name, case in cases.items():
        got, want = api.quad_decompose(case, verts), ro.quad_decompose(case, verts)
        assert np.array_equal(got, want), name
    # needles AND points together: the reference never returns (a triangle paired with itself ends up in
    # the forest and the ancestor walk cycles); the product reports it
    both = np.concatenate([tris[:50], np.stack([tris[:20, 0], tris[:20, 1], tris[:20, 0]], 1), np.stack([tris[:5, 0]] * 3, 1), tris[50:]]).reshape(-1)
    with pytest.raises(api.OrzError, match="does not terminate"):
        api.quad_decompose(both, verts)
    with pytest.raises(api.OrzError, match="out of range"):
        api.quad_decompose(np.array([0, 1, 5000], np.uint32), verts)


@pytest.mark.parametrize("seed,n,target,gran,snap", [
    (1, 24, 512, 8, None),        # smaller than the target: the root is split anyway
    (2, 17, 16, 8, None),         # smallest size with a candidate position
    (3, 4000, 512, 8, None),
    (4, 4000, 512, 8, 10.0),
    (5, 20000, 512, 8, 5.0),
    (6, 5000, 64, 16, 4.0),
    (7, 3001, 100, 7, 2.0),
    (8, 6000, 256, 1, 25.0),
])
def test_generate_batches(seed, n, target, gran, snap):
    boxes = boxes_case(np.random.default_rng(seed), n, snap)
    got, want = api.generate_batches(boxes, target, gran), ro.generate_batches(boxes, target, gran)
    assert same_batches(got, want)
    assert sorted(np.concatenate(got).tolist()) == list(range(n))


def test_generate_batches_signed_zero_and_flatboxes_case():
    rng = np.random.default_rng(21)
    boxes = boxes_case(rng, 3000, 8.0)
    boxes[::3, 1] = boxes[::3, 5] = 0.0          # flat in y at y = 0 ...
    boxes[1::6, 1] = boxes[1::6, 5] = -0.0       # ... some with negative zero
    boxes[::5, 0] = -boxes[::5, 4]               # centres x = +-0
    assert same_batches(api.generate_batches(boxes, 256, 8), ro.generate_batches(boxes, 256, 8))


def test_generate_batches_rejects_what_the_reference_cannot_split():
    boxes = boxes_case(np.random.default_rng(3), 16)
    with pytest.raises(api.OrzError, match="no split position"):
        api.generate_batches(boxes, 512, 8)      # 16 <= 2 * 8: the reference reads areasFromLeft[-1]
    with pytest.raises(api.OrzError):
        api.generate_batches(boxes, 512, 0)


def test_prepare_mesh_matches_reference_scene():
    """The whole of Main.cpp:86-128 through the product's own preparation equals the scene the
    reference's code produced (the prepared scenes the benchmarks run on)."""
    if not (ro.scene_available("Castle") and wl.have_scene("castle")):
        pytest.skip("Castle data missing")
    idx, verts = ro.load_mesh("Castle")
    ps = wl.prepare_mesh("castle", idx, verts, {})
    ref = wl.load_scene("castle")
    assert len(ps.batches) == len(ref.batches)
    assert all(np.array_equal(a.view(np.uint32), b.view(np.uint32)) for a, b in zip(ps.batches, ref.batches))
    assert np.array_equal(ps.ref_min, ref.ref_min) and np.array_equal(ps.ref_max, ref.ref_max)


def test_cpp_dropin_preparation(tmp_path):
    """tests/dropin_prepare.cpp -- Main.cpp:86-128 written against the reference's class names -- built on
    the drop-in headers (host batching: ORZ_PREP_ON_HOST=1, no GPU here) gives the reference's baked scene."""
    import os
    import subprocess

    if not ro.scene_available("Castle"):
        pytest.skip("no Castle data under oracle/_ref/scenes")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "_build", "dropin_prepare")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-mavx2", "-msse4.1", "-Wno-ignored-attributes", "-I", os.path.join(root, "rasterizer_b200", "csrc", "dropin"),
                           "-o", exe, os.path.join(root, "tests", "dropin_prepare.cpp"), "-L", os.path.join(root, "rasterizer_b200"),
                           "-lrasterizer_b200", "-Wl,-rpath," + os.path.join(root, "rasterizer_b200")])
    d = os.path.join(ro.SCENE_DIR, "Castle")
    out = tmp_path / "castle_baked.bin"
    subprocess.check_call([exe, os.path.join(d, "IndexBuffer.bin"), os.path.join(d, "VertexBuffer.bin"), str(out)], env=dict(os.environ, ORZ_PREP_ON_HOST="1"))
    raw = np.fromfile(out, np.uint32)
    s = ro.RefScene.load("Castle")
    assert raw[0] == s.n_occluders
    at = 1
    for i in range(s.n_occluders):
        nq = int(raw[at]); at += 1
        meta = raw[at:at + 12].view(np.float32).reshape(3, 4); at += 12
        packed = raw[at:at + 4 * nq]; at += 4 * nq
        assert nq * 4 == int(s.packet_counts[i]) * 8
        assert np.array_equal(meta.view(np.uint32), np.stack([s.centers[i], s.bounds_min[i], s.bounds_max[i]]).view(np.uint32))
        assert np.array_equal(packed, s.packed(i))
    s.close()


@pytest.mark.parametrize("seed,n,target,gran,snap", [(1, 24, 512, 8, None), (4, 2300, 1024, 8, 10.0), (8, 260, 16, 1, 25.0)])
def test_device_batching_source_emulated_on_the_cpu(seed, n, target, gran, snap):
    """tests/sah_emulation.cpp compiles the CUDA kernel source and the host level loop of the device
    batching against a thread-per-lane emulation of launches, barriers, shuffles and atomics: their logic
    is checked here, where no GPU exists (tests/test_gpu_scene_prep.py runs the real thing)."""
    import ctypes as C
    import os
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = os.path.join(root, "tests", "_build", "libsah_emulation.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", so,
                           os.path.join(root, "tests", "sah_emulation.cpp"), "-lpthread"])
    L = C.CDLL(so)
    L.emu_last_error.restype = C.c_char_p
    boxes = boxes_case(np.random.default_rng(seed), n, snap)
    order, sizes, count = np.zeros(n, np.uint32), np.zeros(n // gran + 2, np.uint32), C.c_uint32()
    rc = L.emu_generate_batches(boxes.ctypes.data_as(C.c_void_p), n, target, gran, order.ctypes.data_as(C.c_void_p),
                                sizes.ctypes.data_as(C.c_void_p), sizes.size, C.byref(count), None)
    assert rc == 0, L.emu_last_error().decode()
    got = np.split(order, np.cumsum(sizes[: count.value])[:-1])
    assert same_batches(got, api.generate_batches(boxes, target, gran))
    if seed == 1:  # the radix histogram scan on its own, with carries across its 1 024-word rounds
        for tiles in (1, 1025, 2500):
            d = np.random.default_rng(tiles).integers(0, 3000, (16, tiles)).astype(np.uint32)
            want = np.concatenate([np.zeros((16, 1), np.uint32), np.cumsum(d, axis=1, dtype=np.uint32)[:, :-1]], axis=1)
            totals, sums = np.zeros(16, np.uint32), d.sum(axis=1, dtype=np.uint32)
            L.emu_scan(d.ctypes.data_as(C.c_void_p), tiles, totals.ctypes.data_as(C.c_void_p))
            assert np.array_equal(d, want) and np.array_equal(totals, sums)


def test_quad_decompose_random_non_manifold_soups():
    """Hundreds of small random triangle soups over a handful of vertices: every directed edge has many owners,
    the candidate graph is dense and non-planar, augmenting paths run through nested blossoms.  Faces are
    non-degenerate (a triangle that is its own neighbour makes the reference loop forever)."""
    rng = np.random.default_rng(1234)
    worst = 0
    for case in range(400):
        nv = int(rng.integers(4, 14))
        nt = int(rng.integers(1, 40))
        verts = np.zeros((nv, 4), np.float32)
        verts[:, :2] = rng.integers(0, 6, (nv, 2))
        verts[:, 2] = rng.choice([0.0, 0.0, 0.0, 0.3, 2.0], nv)      # mostly coplanar: most pairs may merge
        verts[:, 3] = 1.0
        tris = np.stack([rng.permutation(nv)[:3] for _ in range(nt)]).astype(np.uint32)
        got, want = api.quad_decompose(tris.reshape(-1), verts), ro.quad_decompose(tris.reshape(-1), verts)
        assert np.array_equal(got, want), (case, tris.tolist())
        q = got.reshape(-1, 4)
        worst = max(worst, int(np.sum(q[:, 0] != q[:, 3])))
    assert worst >= 10  # the soups do produce many pairs


def test_generate_batches_random_small_cases():
    """Random sizes / targets / granularities over coarse grids of centres (ties everywhere); only shapes the
    reference itself can split (every node it recurses into has more than two granules)."""
    rng = np.random.default_rng(4321)
    for case in range(150):
        gran = int(rng.choice([1, 2, 3, 8, 16]))
        target = int(rng.integers(2 * gran + 1, 2 * gran + 200))
        n = int(rng.integers(2 * gran + 1, 1500))
        boxes = boxes_case(rng, n, float(rng.choice([1.0, 4.0, 25.0, 100.0])))
        got, want = api.generate_batches(boxes, target, gran), ro.generate_batches(boxes, target, gran)
        assert same_batches(got, want), (case, n, target, gran)
