"""Where a per-call frame spends its time: wall clock of every queryVisibility / rasterize call of one Castle
frame through the ctypes mirror of the per-call API (a few frames first, so that the query chains are predicted)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rasterizer_b200 import api, camera as cam, workloads as wl
from oracle import port_oracle as po

ps = wl.load_scene("castle"); w, h = 1920, 1080
po.set_tables()
baked = [po.bake(b, ps.ref_min, ps.ref_max) for b in ps.batches]
c = ps.camera
mvp = cam.view_projection(c["pos"], c["dir"], c["up"], c["fov"], w, h)
order = cam.front_to_back_order(np.stack([b[1] for b in baked]), c["pos"])
ctx = api.Context(0)
occs = [api.Occluder(ctx, b[0], ps.ref_min, ps.ref_max) for b in baked]
r = api.Rasterizer(ctx, w, h)
rows = []
for frame in range(6):
    t_frame = time.perf_counter()
    r.clear(); r.setModelViewProjection(mvp)
    log = []
    for o in order:
        t0 = time.perf_counter(); vis, clip = r.queryVisibility(baked[o][2], baked[o][3]); t1 = time.perf_counter()
        if vis:
            r.rasterize(occs[o], clip)
        t2 = time.perf_counter()
        log.append((int(vis), (t1 - t0) * 1e6, (t2 - t1) * 1e6))
    vis2, _ = r.queryVisibility(baked[order[0]][2], baked[order[0]][3])
    rows.append(((time.perf_counter() - t_frame) * 1e3, log))
ms, log = rows[-1]
q_vis = [q for v, q, _ in log if v]; q_inv = [q for v, q, _ in log if not v]; ras = [x for v, _, x in log if v]
print(json.dumps({"frame_ms": ms, "launches_total": ctx.launch_count, "visible": len(q_vis), "query_us_before_visible": {"median": float(np.median(q_vis)), "sum": float(np.sum(q_vis))},
                  "query_us_invisible": {"median": float(np.median(q_inv)), "max": float(np.max(q_inv)), "sum": float(np.sum(q_inv))},
                  "rasterize_call_us": {"median": float(np.median(ras)), "sum": float(np.sum(ras))}}))
