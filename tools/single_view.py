"""BASELINE configs 1 and 2: ONE view at 1920x1080 (Castle / Sponza default cameras, all occluders
offered in sorted order + queryVisibility on every per-quad AABB).  Single views cannot fill a
B200; this records the latency of the two GPU entry points next to the reference."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rasterizer_b200 import api, camera as cam, workloads as wl
from oracle import ref_oracle as ro

def main():
    res = {}
    ctx = api.Context(0)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    for name in ("castle", "sponza"):
        if not wl.have_scene(name): continue
        ps = wl.load_scene(name)
        w, h = 1920, 1080
        sc = api.Scene.from_prepared(ctx, ps)
        c = ps.camera
        mvp = cam.view_projection(c["pos"], c["dir"], c["up"], c["fov"], w, h)[None]
        pos = np.array(c["pos"], np.float32)[None]
        order = wl.orders_for(sc.centers, pos)
        # (a) batch entry with one view
        out = sc.render_views(w, h, mvp, orders=order, want=("vis", "gate", "depth", "hiz", "quads"))
        blocks = (w // 8) * (h // 8)
        d_mvp, d_ord = torch.from_numpy(mvp).to(dev), torch.from_numpy(order.astype(np.int32)).to(dev)
        d_vis = torch.zeros((1, (sc.n_boxes + 31) // 32), dtype=torch.int32, device=dev)
        d_depth = torch.empty((1, blocks * 64), dtype=torch.int16, device=dev); d_hiz = torch.empty((1, blocks), dtype=torch.int16, device=dev)
        b = api.ViewBatch(); b.width, b.height, b.nViews = w, h, 1
        b.mvps, b.orders, b.visBits, b.depth, b.hiz = d_mvp.data_ptr(), d_ord.data_ptr(), d_vis.data_ptr(), d_depth.data_ptr(), d_hiz.data_ptr()
        for _ in range(3): sc.render_views_raw(b, device=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(10): sc.render_views_raw(b, device=True)
        e1.record(stream); torch.cuda.synchronize()
        batch_ms = e0.elapsed_time(e1) / 10
        # (b) per-call API, the reference's own call sequence
        occs = [api.Occluder(ctx, p, ps.ref_min, ps.ref_max) for p in sc.packed_list]
        r = api.Rasterizer(ctx, w, h)
        def frame():
            r.clear(); r.setModelViewProjection(mvp[0])
            for o in order[0]:
                vis, clip = r.queryVisibility(sc.bounds_min[o], sc.bounds_max[o])
                if vis: r.rasterize(occs[o], clip)
            ctx.synchronize()
        frame()
        t0 = time.perf_counter()
        for _ in range(5): frame()
        percall_ms = (time.perf_counter() - t0) / 5 * 1e3
        boxes = ps.quad_boxes()
        t0 = time.perf_counter(); q = r.query_boxes(boxes); q_ms = (time.perf_counter() - t0) * 1e3
        # reference on one host core
        ref = {}
        if ro.available():
            s = ro.RefScene.from_batches(ps.batches, ps.ref_min, ps.ref_max)
            wall, o4 = ro.bench_views(s, w, h, mvp, order, boxes, 1, 50)
            ref = {"ref_frame_ms": 1e3 * o4[0] / 50, "ref_queries_ms": 1e3 * o4[1] / 50}
            rr = ro.RefRasterizer(w, h); g, _ = rr.frame(s, mvp[0], order[0])
            ref["bit_exact"] = bool(np.array_equal(rr.hiz(), out["hiz"][0]) and np.array_equal(rr.depth(), out["depth"][0]) and np.array_equal(g, out["gate"][0])
                                    and np.array_equal(api.unpack_bits(out["vis"], len(boxes))[0], (rr.query_boxes(boxes) & 1).astype(bool)))
            s.close()
        res[name] = dict(occluders=sc.n_occluders, quads_submitted=int(out["quads"][0]), occludees=len(boxes), visible=int((q & 1).sum()),
                         gpu_batch_api_1view_ms=batch_ms, gpu_per_call_frame_ms=percall_ms, gpu_per_call_queries_ms=q_ms, **ref)
        print(name, res[name], flush=True)
        for o in occs: o.close()
        r.close(); sc.close()
    json.dump(res, open("gpurun_out/single_view.json", "w"), indent=1)
main()
