"""Device-resident step times of the view-batch entry for the library in ORZ_LIB (default: the product's) -- the
A/B tool for kernel variants (one gpurun call times several builds on the same box).
usage: python tools/step_time.py [case ...]   cases: castle1024 probes1024 probes8192 castle1 sponza256 sponza1 castle64
Prints one JSON object; with ORZ_STEP_ONCE=1 runs a single step per case (for ncu captures)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rasterizer_b200 import api, camera as cam, workloads as wl

CASES = {"castle1024": ("castle", 1920, 1080, 1024, "path", True), "probes1024": ("castle", 512, 256, 1024, "probes", False),
         "probes8192": ("castle", 512, 256, 8192, "probes", False), "castle1": ("castle", 1920, 1080, 1, "default", True),
         "sponza256": ("sponza", 1920, 1080, 256, "path", True), "sponza1": ("sponza", 1920, 1080, 1, "default", True),
         "castle64": ("castle", 1920, 1080, 64, "path", True), "castle256": ("castle", 1920, 1080, 256, "path", True),
         "probes2048": ("castle", 512, 256, 2048, "probes", False), "probes4096": ("castle", 512, 256, 4096, "probes", False)}


def main():
    names = sys.argv[1:] or ["castle1024", "probes1024", "castle1"]
    once = os.environ.get("ORZ_STEP_ONCE") == "1"
    dev = torch.device("cuda", 0)
    ctx = api.Context(0)
    if os.environ.get("ORZ_CLUSTER_SIZE"):
        ctx.set_cluster_size(int(os.environ["ORZ_CLUSTER_SIZE"]))
    if os.environ.get("ORZ_GROUP_WARPS"):
        ctx.set_group_warps(int(os.environ["ORZ_GROUP_WARPS"]))
    if os.environ.get("ORZ_CLUSTER_VIEWS"):
        ctx.set_cluster_views(int(os.environ["ORZ_CLUSTER_VIEWS"]))
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    out = {"lib": os.path.basename(api.LIB_PATH)}
    scenes = {}
    for name in names:
        sname, w, h, n, kind, targets = CASES[name]
        if sname not in scenes:
            ps = wl.load_scene(sname)
            scenes[sname] = (ps, api.Scene.from_prepared(ctx, ps))
        ps, scene = scenes[sname]
        if kind == "default":
            c = ps.camera
            mvps = cam.view_projection(c["pos"], c["dir"], c["up"], c["fov"], w, h)[None].astype(np.float32)
            poss = np.array(c["pos"], np.float32)[None]
        else:
            stride = int(os.environ.get("ORZ_VIEW_STRIDE", "1"))  # probes: every stride-th view of the 8192 (what one of `stride` GPUs gets)
            if kind == "probes" and stride > 1:
                mvps, poss = wl.probe_views(ps, n * stride, w, h)
                mvps, poss = np.ascontiguousarray(mvps[::stride]), np.ascontiguousarray(poss[::stride])
            else:
                mvps, poss = (wl.camera_path if kind == "path" else wl.probe_views)(ps, n, w, h)
        blocks, words = (w // 8) * (h // 8), (scene.n_boxes + 31) // 32
        d_mvps, d_pos = torch.from_numpy(mvps).to(dev), torch.from_numpy(poss).to(dev)
        d_vis = torch.zeros((n, words), dtype=torch.int32, device=dev)
        b = api.ViewBatch()
        b.width, b.height, b.nViews, b.flags = w, h, n, 0
        b.mvps, b.camPos, b.visBits = d_mvps.data_ptr(), d_pos.data_ptr(), d_vis.data_ptr()
        if targets:
            d_depth = torch.empty((n, blocks * 64), dtype=torch.int16, device=dev)
            d_hiz = torch.empty((n, blocks), dtype=torch.int16, device=dev)
            b.depth, b.hiz = d_depth.data_ptr(), d_hiz.data_ptr()
        reps = 1 if once else (10 if n >= 256 else 50)
        for _ in range(1 if once else 3):
            scene.render_views_raw(b, device=True)
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        n0 = ctx.launch_count
        for a, z in ev:
            a.record(stream)
            scene.render_views_raw(b, device=True)
            z.record(stream)
        torch.cuda.synchronize()
        ms = sorted(a.elapsed_time(z) for a, z in ev)
        out[name] = {"ms_median": ms[len(ms) // 2], "ms_min": ms[0], "views_per_s": n / ms[len(ms) // 2] * 1e3,
                     "vis_checksum": int(d_vis.to(torch.int64).sum().item()), "launches_per_step": (ctx.launch_count - n0) / reps}
    print(json.dumps(out))


main()
