"""Scene preparation: orz_quad_decompose / orz_generate_batches against the
UNMODIFIED reference (QuadDecomposition.cpp, SurfaceAreaHeuristic.cpp compiled into oracle/_ref):
same quads in the same order, same batches in the same order.  CPU only; the device batching is
checked in tests/test_gpu_scene_prep.py."""
import numpy as np
import pytest

from oracle import ref_oracle as ro
from rasterizer_b200 import api
from rasterizer_b200 import workloads as wl

pytestmark = pytest.mark.skipif(not ro.available(), reason="oracle/_ref/libref_oracle.so not built (needs /root/reference)")


@pytest.fixture(scope="module", autouse=True)
def host_rsqrt():
    api.set_rsqrt_table(None)  # canMergeTrianglesToQuad normalises with this host's rsqrtps, as the reference build does
    yield


def _same_batches(a, b):
    return len(a) == len(b) and all(np.array_equal(x, y) for x, y in zip(a, b))


@pytest.mark.parametrize("name", ["Castle", "Sponza"])
def test_reference_scenes(name):
    """Main.cpp:86-107 on the reference's own meshes: decompose, pad, per-quad AABBs, batches."""
    if not ro.scene_available(name):
        pytest.skip(f"no {name} data under oracle/_ref/scenes")
    idx, verts = ro.load_mesh(name)
    quads = api.quad_decompose(idx, verts)
    assert np.array_equal(quads, ro.quad_decompose(idx, verts))
    boxes = wl.quad_aabbs(wl.pad_quads(quads), verts)
    assert _same_batches(api.generate_batches(boxes, 512, 8), ro.generate_batches(boxes, 512, 8))


def _terrain(rng, n, bump, shuffle=True):
    """Consistently wound Delaunay triangulation of random points: the dual graph is full of odd
    cycles (blossoms), `bump` makes part of the candidate pairs fail the planarity test."""
    from scipy.spatial import Delaunay

    pts = rng.uniform(0, 40, (n, 2)).astype(np.float32)
    tri = Delaunay(pts.astype(np.float64)).simplices.astype(np.uint32)
    a, b, c = pts[tri[:, 0]], pts[tri[:, 1]], pts[tri[:, 2]]
    flip = ((b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (c[:, 0] - a[:, 0])) < 0
    tri[flip] = tri[flip][:, [0, 2, 1]]
    if shuffle:
        tri = tri[rng.permutation(len(tri))]
    z = (rng.uniform(0, 1, n) ** 4 * bump).astype(np.float32)
    verts = np.stack([pts[:, 0], z, pts[:, 1], np.ones(n, np.float32)], axis=1).astype(np.float32)
    return tri.reshape(-1), verts


@pytest.mark.parametrize("seed,n,bump", [(1, 60, 0.0), (2, 400, 0.0), (3, 400, 6.0), (4, 3000, 2.0), (5, 3000, 30.0), (6, 9000, 0.5)])
def test_quad_decompose_terrain(seed, n, bump):
    idx, verts = _terrain(np.random.default_rng(seed), n, bump)
    got, want = api.quad_decompose(idx, verts), ro.quad_decompose(idx, verts)
    assert np.array_equal(got, want)
    pairs = int(np.sum(got.reshape(-1, 4)[:, 0] != got.reshape(-1, 4)[:, 3]))
    if bump == 0.0:
        assert pairs > 0.45 * (len(idx) // 3) / 1  # flat sheet: nearly every triangle finds a partner


def test_quad_decompose_awkward_meshes():
    rng = np.random.default_rng(11)
    idx, verts = _terrain(rng, 300, 1.0)
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


def _boxes(rng, n, snap=None):
    c = rng.uniform(-50, 50, (n, 3))
    if snap:
        c = np.round(c / snap) * snap  # many equal centres: the stable sorts' tie order becomes visible
    e = rng.uniform(0.0, 3.0, (n, 3)) if not snap else np.round(rng.uniform(0.0, 3.0, (n, 3)))
    mn, mx = (c - e).astype(np.float32), (c + e).astype(np.float32)
    one = np.ones((n, 1), np.float32)
    return np.concatenate([mn, one, mx, one], axis=1)


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
    boxes = _boxes(np.random.default_rng(seed), n, snap)
    got, want = api.generate_batches(boxes, target, gran), ro.generate_batches(boxes, target, gran)
    assert _same_batches(got, want)
    assert sorted(np.concatenate(got).tolist()) == list(range(n))


def test_generate_batches_signed_zero_and_flat_boxes():
    rng = np.random.default_rng(21)
    boxes = _boxes(rng, 3000, 8.0)
    boxes[::3, 1] = boxes[::3, 5] = 0.0          # flat in y at y = 0 ...
    boxes[1::6, 1] = boxes[1::6, 5] = -0.0       # ... some with negative zero
    boxes[::5, 0] = -boxes[::5, 4]               # centres x = +-0
    assert _same_batches(api.generate_batches(boxes, 256, 8), ro.generate_batches(boxes, 256, 8))


def test_generate_batches_rejects_what_the_reference_cannot_split():
    boxes = _boxes(np.random.default_rng(3), 16)
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
