"""Shared inputs of the scene-preparation tests (CPU and GPU): terrains with odd dual cycles, tie-heavy boxes."""
import numpy as np

def same_batches(a, b):
    return len(a) == len(b) and all(np.array_equal(x, y) for x, y in zip(a, b))


def terrain(rng, n, bump, shuffle=True):
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


def boxes_case(rng, n, snap=None):
    c = rng.uniform(-50, 50, (n, 3))
    if snap:
        c = np.round(c / snap) * snap  # many equal centres: the stable sorts' tie order becomes visible
    e = rng.uniform(0.0, 3.0, (n, 3)) if not snap else np.round(rng.uniform(0.0, 3.0, (n, 3)))
    mn, mx = (c - e).astype(np.float32), (c + e).astype(np.float32)
    one = np.ones((n, 1), np.float32)
    return np.concatenate([mn, one, mx, one], axis=1)
