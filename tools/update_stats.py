"""Work statistics of the reference's block loop on a scene's camera path, from the instrumented oracle port (CPU,
planning only): block visits, visits the HiZ test rejects, depth + HiZ updates, updates that change no pixel, and
updates an exact per-block bound (largest corner sample <= the block's HiZ) could skip.  DESIGN section 9.
usage: python tools/update_stats.py   -> profiles/r1_update_stats.json"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from oracle import port_oracle as po
from rasterizer_b200 import camera as cam, workloads as wl

L = po.lib()
L.orc_stats.argtypes, L.orc_stats.restype = [C.c_void_p, C.c_int], None
res = {}
for name, nv in (("castle", 16), ("sponza", 6)):
    if not wl.have_scene(name):
        continue
    ps = wl.load_scene(name)
    po.set_tables()
    baked = [po.bake(b, ps.ref_min, ps.ref_max) for b in ps.batches]
    packed = [b[0] for b in baked]
    centers, bmin, bmax = (np.stack([b[i] for b in baked]) for i in (1, 2, 3))
    w, h = 1920, 1080
    mvps, poss = wl.camera_path(ps, 1024, w, h)
    port = po.PortRasterizer(w, h)
    L.orc_stats(None, 1)
    for v in np.linspace(0, 1023, nv).astype(int):
        port.frame(packed, bmin, bmax, ps.ref_min, ps.ref_max, mvps[v], cam.front_to_back_order(centers, poss[v]))
    st = np.zeros(5, np.uint64)
    L.orc_stats(st.ctypes.data_as(C.c_void_p), 1)
    visits, rejected, updates, unchanged, bound = (int(x) for x in st)
    res[name] = dict(views=nv, width=w, height=h, visits_per_view=visits // nv, hiz_rejected_per_view=rejected // nv,
                     updates_per_view=updates // nv, updates_without_change_per_view=unchanged // nv,
                     skippable_by_block_bound_per_view=bound // nv, unchanged_fraction=unchanged / updates, bound_fraction=bound / updates)
    print(name, res[name])
    port.close()
json.dump(res, open("profiles/r1_update_stats.json", "w"), indent=1)
