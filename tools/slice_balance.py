import os, sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from rasterizer_b200 import api, workloads as wl
ctx = api.Context(0); dev = torch.device("cuda", 0)
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
ps = wl.load_scene("castle"); sc = api.Scene.from_prepared(ctx, ps)
w, h, nv, world = 1920, 1080, 1024, 8
blocks = (w // 8) * (h // 8)
mv, po = wl.camera_path(ps, nv * world, w, h)
d_vis = torch.zeros((nv, (sc.n_boxes + 31) // 32), dtype=torch.int32, device=dev)
d_depth = torch.empty((nv, blocks * 64), dtype=torch.int16, device=dev); d_hiz = torch.empty((nv, blocks), dtype=torch.int16, device=dev)
def run(m, p):
    d_m, d_p = torch.from_numpy(np.ascontiguousarray(m)).to(dev), torch.from_numpy(np.ascontiguousarray(p)).to(dev)
    b = api.ViewBatch(); b.width, b.height, b.nViews = w, h, nv
    b.mvps, b.camPos, b.visBits, b.depth, b.hiz = d_m.data_ptr(), d_p.data_ptr(), d_vis.data_ptr(), d_depth.data_ptr(), d_hiz.data_ptr()
    for _ in range(2): sc.render_views_raw(b, device=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(5): sc.render_views_raw(b, device=True)
    e1.record(stream); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 5
print("contiguous slices:", [round(run(mv[r * nv:(r + 1) * nv], po[r * nv:(r + 1) * nv]), 2) for r in range(world)])
print("interleaved slices:", [round(run(mv[r::world], po[r::world]), 2) for r in range(world)])
