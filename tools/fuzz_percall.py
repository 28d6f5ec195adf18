"""Randomised sequences through the per-call API (queryVisibility / rasterize / clear / setModelViewProjection / query2D in
random order, boxes of occluders and occludees, repeated frames so that the predicted query chains are right, partly right
and wrong) against the scalar port of the reference, answer by answer; depth and HiZ compared every few hundred calls.
usage: python tools/fuzz_percall.py [seed] [ops]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import port_oracle as po
from rasterizer_b200 import api, camera as cam, workloads as wl


def main():
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    n_ops = int(sys.argv[2]) if len(sys.argv) > 2 else 6000
    rng = np.random.default_rng(seed)
    ps = wl.load_scene("castle") if wl.have_scene("castle") else wl.synthetic_city(n_blocks=5)
    po.set_tables()
    baked = [po.bake(b, ps.ref_min, ps.ref_max) for b in ps.batches]
    boxes = ps.quad_boxes()
    ctx = api.Context(0)
    occs = [api.Occluder(ctx, b[0], ps.ref_min, ps.ref_max) for b in baked]
    out = {"seed": seed, "ops": 0, "queries": 0, "rasterized": 0, "clears": 0, "compares": 0, "sizes": []}
    t0 = time.time()
    for (w, h) in ((1920, 1080), (512, 256), (640, 360)):
        r, port = api.Rasterizer(ctx, w, h), po.PortRasterizer(w, h)
        mvps, poss = wl.camera_path(ps, 16, w, h)
        ctx.set_tile_height(0, 1 if rng.random() < 0.7 else 4)
        cur = 0
        order = cam.front_to_back_order(np.stack([b[1] for b in baked]), poss[cur])
        r.clear(); port.clear(); r.setModelViewProjection(mvps[cur]); port.set_mvp(mvps[cur])
        pos_in_order = 0
        for op in range(n_ops // 3):
            x = rng.random()
            if x < 0.02:
                r.clear(); port.clear(); out["clears"] += 1
                pos_in_order = 0
                if rng.random() < 0.5:   # a new frame: same camera (right predictions), a neighbour (partly right) or far away (wrong)
                    cur = int(rng.choice([cur, (cur + 1) % 16, int(rng.integers(16))]))
                    r.setModelViewProjection(mvps[cur]); port.set_mvp(mvps[cur])
                    order = cam.front_to_back_order(np.stack([b[1] for b in baked]), poss[cur])
                    if rng.random() < 0.2:
                        order = rng.permutation(order)
            elif x < 0.05:
                cur = int(rng.integers(16)); r.setModelViewProjection(mvps[cur]); port.set_mvp(mvps[cur])   # matrix change in mid-frame
            elif x < 0.60:   # the application's loop: next occluder in order, query, rasterize when visible
                o = int(order[pos_in_order % len(order)]); pos_in_order += 1
                g = port.query(baked[o][2], baked[o][3])
                vis, clip = r.queryVisibility(baked[o][2], baked[o][3])
                assert (int(vis) | (int(clip) << 1)) == g, ("loop query", w, h, op, o)
                out["queries"] += 1
                if vis and rng.random() < 0.9:
                    r.rasterize(occs[o], clip); port.rasterize(baked[o][0], ps.ref_min, ps.ref_max, clip); out["rasterized"] += 1
            elif x < 0.80:   # queries out of the blue: occludee boxes, random occluders
                if rng.random() < 0.5:
                    bx = boxes[int(rng.integers(len(boxes)))]; mn, mx = bx[:4], bx[4:]
                else:
                    o = int(rng.integers(len(baked))); mn, mx = baked[o][2], baked[o][3]
                g = port.query(mn, mx)
                vis, clip = r.queryVisibility(mn, mx)
                assert (int(vis) | (int(clip) << 1)) == g, ("random query", w, h, op)
                out["queries"] += 1
            elif x < 0.88:   # rasterize without asking (either template)
                o = int(rng.integers(len(baked))); clip = bool(rng.random() < 0.5)
                r.rasterize(occs[o], clip); port.rasterize(baked[o][0], ps.ref_min, ps.ref_max, clip); out["rasterized"] += 1
            else:            # query2D directly
                x0, x1 = sorted(rng.integers(0, w, 2)); y0, y1 = sorted(rng.integers(0, h, 2)); z = int(rng.integers(0, 65536))
                assert r.query2D(int(x0), int(x1), int(y0), int(y1), z) == port.query2d(int(x0), int(x1), int(y0), int(y1), z), ("query2D", w, h, op)
                out["queries"] += 1
            out["ops"] += 1
            if op % 400 == 399:
                d, hz = r.download()
                assert np.array_equal(hz, port.hiz()) and np.array_equal(d, port.depth()), ("buffers", w, h, op)
                out["compares"] += 1
        d, hz = r.download()
        assert np.array_equal(hz, port.hiz()) and np.array_equal(d, port.depth()), ("buffers at the end", w, h)
        out["compares"] += 1
        out["sizes"].append([w, h])
        r.close(); port.close()
    out["seconds"] = time.time() - t0
    out["result"] = "every answer and every buffer identical"
    print(json.dumps(out))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(f"gpurun_out/fuzz_percall_seed{seed}.json", "w"), indent=1)


main()
