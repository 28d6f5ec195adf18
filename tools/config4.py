"""BASELINE config 4 at full size: synthetic 10 M-triangle (5 M quad) occluder soup at 3840x2160,
camera inside the geometry, every batch through rasterize<true>, no gate.  Times the CUDA wide
path and checks it bit-for-bit against the unmodified reference (oracle/_ref) on the same box."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rasterizer_b200 import api, camera as cam, workloads as wl

def main():
    nq = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
    w, h = 3840, 2160
    t0 = time.time(); ps = wl.synthetic_soup(nq); print(f"generated {ps.n_quads} quads in {len(ps.batches)} batches, {time.time()-t0:.1f}s", flush=True)
    ctx = api.Context(0)
    t0 = time.time(); baked = [api.bake(b, ps.ref_min, ps.ref_max) for b in ps.batches]; print(f"host bake {time.time()-t0:.1f}s", flush=True)
    packed = [b[0] for b in baked]
    centers, bmin, bmax = (np.stack([b[i] for b in baked]) for i in (1, 2, 3))
    boxes = ps.quad_boxes()[::50]
    sc = api.Scene(ctx, packed, ps.ref_min, ps.ref_max, bmin, bmax, centers, boxes)
    c = ps.camera
    mvps = np.stack([cam.view_projection(c["pos"], d, c["up"], c["fov"], w, h) for d in ((0, 0, 1), (0.7, 0.1, -0.7))])
    poss = np.zeros((2, 3), np.float32)
    orders = wl.orders_for(centers, poss)
    flags = api.BATCH_NO_GATE | api.BATCH_FORCE_CLIPPED
    dev = torch.device("cuda", 0)
    blocks = (w // 8) * (h // 8)
    d_mvps, d_ord = torch.from_numpy(mvps).to(dev), torch.from_numpy(orders.astype(np.int32)).to(dev)
    d_depth = torch.empty((2, blocks * 64), dtype=torch.int16, device=dev); d_hiz = torch.empty((2, blocks), dtype=torch.int16, device=dev)
    d_vis = torch.zeros((2, (len(boxes) + 31) // 32), dtype=torch.int32, device=dev)
    b = api.ViewBatch(); b.width, b.height, b.nViews, b.flags = w, h, 2, flags
    b.mvps, b.orders, b.depth, b.hiz, b.visBits = d_mvps.data_ptr(), d_ord.data_ptr(), d_depth.data_ptr(), d_hiz.data_ptr(), d_vis.data_ptr()
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    for _ in range(2): sc.render_views_raw(b, device=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(3): sc.render_views_raw(b, device=True)
    e1.record(stream); torch.cuda.synchronize()
    ms_view = e0.elapsed_time(e1) / 6
    res = {"quads": ps.n_quads, "gpu_ms_per_view": ms_view, "gpu_mquads_per_s": ps.n_quads / ms_view / 1e3}
    print(res, flush=True)
    hiz = d_hiz.cpu().numpy().view(np.uint16); depth = d_depth.cpu().numpy().view(np.uint16)
    res["blocks_touched"] = [int((hiz[v] != 1).sum()) for v in range(2)]
    from oracle import ref_oracle as ro
    if ro.available():
        t0 = time.time(); s = ro.RefScene.from_batches(ps.batches, ps.ref_min, ps.ref_max); print(f"reference bake {time.time()-t0:.1f}s", flush=True)
        assert all(np.array_equal(s.packed(i), packed[i]) for i in range(0, len(packed), 97))
        r = ro.RefRasterizer(w, h)
        for v in range(2):
            t0 = time.perf_counter(); r.submit_all(s, mvps[v], orders[v], True, zero_depth=True); dt = time.perf_counter() - t0
            same_h, same_d = np.array_equal(r.hiz(), hiz[v]), np.array_equal(r.depth(), depth[v])
            q = r.query_boxes(boxes)
            vis = api.unpack_bits(d_vis[v:v+1].cpu().numpy().view(np.uint32), len(boxes))[0]
            same_v = np.array_equal(vis, (q & 1).astype(bool))
            print(f"view {v}: reference {dt*1e3:.1f} ms, hiz equal {same_h}, depth equal {same_d}, vis equal {same_v}", flush=True)
            res[f"ref_ms_view{v}"] = dt * 1e3; res[f"bit_exact_view{v}"] = bool(same_h and same_d and same_v)
    json.dump(res, open("gpurun_out/config4.json", "w"))
main()
