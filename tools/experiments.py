"""Timing experiments on the GPU box (not part of the bench contract): splits the step time of
the view-batch kernel by feature and sweeps the group size."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rasterizer_b200 import api, workloads as wl

def timed(ctx, scene, batch, stream, reps=10):
    for _ in range(2):
        scene.render_views_raw(batch, device=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(reps):
        scene.render_views_raw(batch, device=True)
    b.record(stream)
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "castle"
    w, h, nv = (int(x) for x in (sys.argv[2:5] if len(sys.argv) > 4 else (1920, 1080, 1024)))
    ps = wl.load_scene(name if wl.have_scene(name) else "city")
    dev = torch.device("cuda", 0)
    ctx = api.Context(0)
    scene = api.Scene.from_prepared(ctx, ps)
    mvps, poss = wl.camera_path(ps, nv, w, h)
    blocks = (w // 8) * (h // 8)
    words = (scene.n_boxes + 31) // 32
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    d_mvps, d_pos = torch.from_numpy(mvps).to(dev), torch.from_numpy(poss).to(dev)
    d_vis = torch.zeros((nv, words), dtype=torch.int32, device=dev)
    d_depth = torch.empty((nv, blocks * 64), dtype=torch.int16, device=dev)
    d_hiz = torch.empty((nv, blocks), dtype=torch.int16, device=dev)
    def mk(vis=True, targets=True, flags=0):
        b = api.ViewBatch(); b.width, b.height, b.nViews, b.flags = w, h, nv, flags
        b.mvps, b.camPos = d_mvps.data_ptr(), d_pos.data_ptr()
        if vis: b.visBits = d_vis.data_ptr()
        if targets: b.depth, b.hiz = d_depth.data_ptr(), d_hiz.data_ptr()
        return b
    res = {}
    for trav in (1, 2):
        ctx.set_traversal(trav)
        for gw in (1, 2, 4, 8):
            ctx.set_group_warps(gw)
            res[f"t{trav}_gw{gw}_full"] = timed(ctx, scene, mk(), stream)
            res[f"t{trav}_gw{gw}_noqueries"] = timed(ctx, scene, mk(vis=False), stream)
    for k, v in res.items():
        print(f"{k:28s} {v:8.3f} ms  {nv / v * 1e3:10.0f} views/s")
    json.dump(res, open("gpurun_out/experiments.json", "w"))

main()
