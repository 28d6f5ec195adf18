"""Randomised parity soak (not part of the test suite: run once per build on the GPU box, result under profiles/): random
camera positions inside the scene box with random look directions and fields of view, several resolutions and batch sizes
(odd sizes included), every view of every batch through the default path of the C ABI and compared bit for bit -- gate
bytes, HiZ, depth, visibility and needsClipping bits, quads submitted -- with the unmodified reference (oracle/_ref) on all
host threads.  usage: python tools/soak_parity.py [seed]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import ref_oracle as ro
from rasterizer_b200 import api, camera as cam, workloads as wl

ALL = ("vis", "clip", "gate", "depth", "hiz", "quads")


def random_views(ps, n, w, h, rng):
    lo, hi = ps.ref_min[:3].astype(np.float64), ps.ref_max[:3].astype(np.float64)
    up = ps.camera.get("up", (0, 1, 0))
    mvps, poss = np.zeros((n, 16), np.float32), np.zeros((n, 3), np.float32)
    for i in range(n):
        pos = lo + (hi - lo) * rng.uniform(-0.1, 1.1, 3)      # inside, on and a little outside the scene box
        d = rng.normal(size=3); d /= np.linalg.norm(d)
        if abs(np.dot(d, up)) > 0.97:
            d = np.array([1.0, 0.0, 0.0])
        fov = rng.uniform(0.4, 1.6)
        mvps[i] = cam.view_projection(pos, d, up, fov, w, h).reshape(16)
        poss[i] = pos
    return mvps, poss


def main():
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 7
    rng = np.random.default_rng(seed)
    ctx = api.Context(0)
    out = {"seed": seed, "cases": []}
    t0 = time.time()
    for name in ("castle", "sponza"):
        if not wl.have_scene(name):
            continue
        ps = wl.load_scene(name)
        ref = ro.RefScene.from_batches(ps.batches, ps.ref_min, ps.ref_max)
        boxes = ps.quad_boxes()[::2]
        sc = api.Scene.from_prepared(ctx, ps, boxes=boxes)
        for (w, h), n in (((1920, 1080), 333), ((1280, 720), 517), ((512, 256), 1500), ((640, 360), 3), ((1920, 1080), 1), ((3840, 2160), 5),
                          ((256, 128), 2049), ((1024, 1024), 64)):
            if name == "sponza":
                n = max(1, n // 4)
            mvps, poss = random_views(ps, n, w, h, rng)
            got = sc.render_views(w, h, mvps, cam_pos=poss, want=ALL)
            orders = wl.orders_for(ref.centers, poss)
            mism = ro.check_views(ref, w, h, mvps, orders, boxes, gate=got["gate"], depth=got["depth"], hiz=got["hiz"], vis=got["vis"], clip=got["clip"],
                                  quads=got["quads"])
            ok = not mism.any()
            out["cases"].append({"scene": name, "size": [w, h], "views": n, "bit_exact": bool(ok), "visible_occluders_mean": float((got["gate"] & 1).sum(axis=1).mean())})
            print(out["cases"][-1], flush=True)
            assert ok, ro.describe_mismatch(mism)
        sc.close(); ref.close()
    out["views_total"] = int(sum(c["views"] for c in out["cases"]))
    out["seconds"] = time.time() - t0
    print(json.dumps(out))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/soak_parity.json", "w"), indent=1)


main()
