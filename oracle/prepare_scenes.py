"""BUILD-TIME DATA TOOL: turn the reference's Castle / Sponza meshes (copied to oracle/_ref/scenes, git-ignored)
into prepared scenes under scenes/_prepared/ (git-ignored) with the PRODUCT's own preparation
(workloads.prepare_mesh: orz_quad_decompose + orz_generate_batches, Main.cpp:86-113), and check the result
against what the reference's unmodified code (QuadDecomposition.cpp + SurfaceAreaHeuristic.cpp compiled into
oracle/_ref/libref_oracle.so) makes of the same files: same batches, same vertices, same reference AABB.
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
        ps = wl.prepare_mesh(name.lower(), *ro.load_mesh(name), dict(camera))
        batches = [s.batch_vertices(i) for i in range(s.n_occluders)]
        assert len(batches) == len(ps.batches) and all(np.array_equal(a.view(np.uint32), b.view(np.uint32)) for a, b in zip(batches, ps.batches))
        assert np.array_equal(ps.ref_min, s.ref_min) and np.array_equal(ps.ref_max, s.ref_max)
        ps.save(wl.prepared_path(name))
        back = wl.PreparedScene.load(wl.prepared_path(name))
        assert all(np.array_equal(a, b) for a, b in zip(batches, back.batches))
        assert np.array_equal(back.quad_boxes().view(np.uint32), s.boxes.view(np.uint32))
        print(f"{name}: {len(batches)} batches, {ps.n_quads} quads -> {wl.prepared_path(name)}")
        s.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
