"""One device batching of a prepared scene's quad boxes (for ncu launch lists).  usage: python tools/prep_one.py [castle|sponza]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rasterizer_b200 import api, workloads as wl
name = sys.argv[1] if len(sys.argv) > 1 else "castle"
ctx = api.Context(0)
batches = ctx.generate_batches(wl.load_scene(name).quad_boxes(), 512, 8)
print(name, len(batches), "batches,", ctx.launch_count, "launches")
