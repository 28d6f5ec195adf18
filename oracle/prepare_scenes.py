"""BUILD-TIME DATA TOOL (test infrastructure side): run the reference's own offline preprocessing
(QuadDecomposition.cpp + SurfaceAreaHeuristic.cpp + the padding / batching of Main.cpp:86-128,
compiled unmodified into oracle/_ref/libref_oracle.so) over the reference's Castle / Sponza data
and store the resulting quad batches as prepared scenes under scenes/_prepared/ (git-ignored).

Quad decomposition and SAH batching are out of scope for the B200 path;
the product only ever sees their output: batches of quads + one reference AABB.
Usage: python -m oracle.prepare_scenes
"""
from __future__ import annotations

import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import ref_oracle as ro  # noqa: E402
from rasterizer_b200 import camera as cam  # noqa: E402
from rasterizer_b200 import workloads as wl  # noqa: E402


def main() -> int:
    if not ro.available():
        print("oracle/_ref/libref_oracle.so missing: run `make -C oracle ref` where /root/reference exists")
        return 1
    for name, camera in (("Castle", cam.CASTLE_CAMERA), ("Sponza", cam.SPONZA_CAMERA)):
        if not ro.scene_available(name):
            print(f"skip {name}: no scene data under oracle/_ref/scenes")
            continue
        s = ro.RefScene.load(name)
        batches = [s.batch_vertices(i) for i in range(s.n_occluders)]
        ps = wl.PreparedScene(name.lower(), batches, s.ref_min.copy(), s.ref_max.copy(), dict(camera))
        ps.save(wl.prepared_path(name))
        back = wl.PreparedScene.load(wl.prepared_path(name))
        assert all(np.array_equal(a, b) for a, b in zip(batches, back.batches))
        assert np.array_equal(back.quad_boxes().view(np.uint32), s.boxes.view(np.uint32))
        print(f"{name}: {len(batches)} batches, {ps.n_quads} quads -> {wl.prepared_path(name)}")
        s.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
