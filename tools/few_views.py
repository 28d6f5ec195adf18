"""Latency of small view batches (1..128 views, Castle 1080p, device-pointer entry) through the
cluster-per-view kernel and through the one-CTA-per-view batch kernel; picks the crossover."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rasterizer_b200 import api, workloads as wl

def main():
    scene = sys.argv[1] if len(sys.argv) > 1 else "castle"
    w, h = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1920, 1080)
    ctx = api.Context(0)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    ps = wl.load_scene(scene)
    sc = api.Scene.from_prepared(ctx, ps)
    blocks = (w // 8) * (h // 8)
    res = {}
    sweep = [int(x) for x in os.environ.get('FEW_VIEWS_NV', '1,2,4,8,16,18,32,36,64,128').split(',')]
    paths = [("cluster", 1 << 20), ("cta", 0)] if not os.environ.get('FEW_VIEWS_ONLY_CLUSTER') else [("cluster", 1 << 20)]
    if os.environ.get('FEW_VIEWS_CSWEEP'):
        paths = [("cluster", 1 << 20)] + [(f"c{c}", -c) for c in (1, 2, 4, 8, 16)]
    for nv in sweep:
        mvps, poss = wl.camera_path(ps, nv, w, h)
        d_mvp, d_pos = torch.from_numpy(mvps).to(dev), torch.from_numpy(poss).to(dev)
        d_vis = torch.zeros((nv, (sc.n_boxes + 31) // 32), dtype=torch.int32, device=dev)
        d_depth = torch.empty((nv, blocks * 64), dtype=torch.int16, device=dev); d_hiz = torch.empty((nv, blocks), dtype=torch.int16, device=dev)
        b = api.ViewBatch(); b.width, b.height, b.nViews = w, h, nv
        b.mvps, b.camPos, b.visBits, b.depth, b.hiz = d_mvp.data_ptr(), d_pos.data_ptr(), d_vis.data_ptr(), d_depth.data_ptr(), d_hiz.data_ptr()
        row = {}
        keep = None
        for label, cv in paths:
            ctx.set_cluster_size(-cv if cv < 0 else 0)
            ctx.set_cluster_views(1 << 20 if cv < 0 else cv)
            for _ in range(3): sc.render_views_raw(b, device=True)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 20 if nv <= 16 else 8
            e0.record(stream)
            for _ in range(reps): sc.render_views_raw(b, device=True)
            e1.record(stream); torch.cuda.synchronize()
            row[label + "_ms"] = e0.elapsed_time(e1) / reps
            got = (d_vis.cpu().numpy().copy(), d_hiz.cpu().numpy().copy())
            if keep is None: keep = got
            else: row["same"] = bool(np.array_equal(keep[0], got[0]) and np.array_equal(keep[1], got[1]))
        row = {k: (round(v, 4) if isinstance(v, float) else v) for k, v in row.items()}
        row["cluster_views_per_s"] = nv / row["cluster_ms"] * 1e3
        if "cta_ms" in row: row["cta_views_per_s"] = nv / row["cta_ms"] * 1e3
        res[nv] = row
        print(nv, row, flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open(f"gpurun_out/few_views_{scene}_{w}x{h}.json", "w"), indent=1)
main()
