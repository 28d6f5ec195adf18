"""One device batching (for ncu launch lists / captures).  usage: python tools/prep_one.py castle|sponza|soup [n_boxes]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rasterizer_b200 import api, workloads as wl
name = sys.argv[1] if len(sys.argv) > 1 else "castle"
if name == "soup":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
    rng = np.random.default_rng(5)
    c, e, one = rng.uniform(-500, 500, (n, 3)).astype(np.float32), rng.uniform(0.2, 3.0, (n, 3)).astype(np.float32), np.ones((n, 1), np.float32)
    boxes = np.concatenate([c - e, one, c + e, one], axis=1)
else:
    boxes = wl.load_scene(name).quad_boxes()
ctx = api.Context(0)
batches = ctx.generate_batches(boxes, 512, 8)
print(name, boxes.shape[0], "boxes,", len(batches), "batches,", ctx.launch_count, "launches")
