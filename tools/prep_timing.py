"""Scene preparation timings: quad decomposition (host, against the reference) and SAH
batching (host / B200 / reference) on the prepared scenes' quad boxes and on synthetic soups; every result is
checked for equality.  Writes gpurun_out/prep_timing.json.   usage: python tools/prep_timing.py [max_soup_quads]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from oracle import ref_oracle as ro
from rasterizer_b200 import api, workloads as wl


def timed(fn, *a):
    t0 = time.perf_counter()
    r = fn(*a)
    return r, (time.perf_counter() - t0) * 1e3


def same(a, b):
    return len(a) == len(b) and all(np.array_equal(x, y) for x, y in zip(a, b))


def main():
    max_soup = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
    ctx = api.Context(0)
    res = {"host_threads": os.cpu_count()}
    ctx.generate_batches(wl.synthetic_city().quad_boxes(), 64, 8)  # warm-up: module load
    cases = []
    for name in ("castle", "sponza"):
        if wl.have_scene(name):
            cases.append((name, wl.load_scene(name).quad_boxes()))
    rng = np.random.default_rng(5)
    n = 500_000
    while n <= max_soup:
        c = rng.uniform(-500, 500, (n, 3)).astype(np.float32)
        e = rng.uniform(0.2, 3.0, (n, 3)).astype(np.float32)
        one = np.ones((n, 1), np.float32)
        cases.append((f"soup_{n}", np.concatenate([c - e, one, c + e, one], axis=1)))
        n *= 4
    for name, boxes in cases:
        runs = [timed(ctx.generate_batches, boxes, 512, 8) for _ in range(6)]  # wall clock: upload, level loop, download
        dev, times = runs[0][0], sorted(t for _, t in runs[1:])
        launches = ctx.launch_count
        host, t_host = timed(api.generate_batches, boxes, 512, 8)
        r = dict(boxes=int(boxes.shape[0]), batches=len(dev), device_ms=times[0], device_ms_median=times[len(times) // 2],
                 device_launches=int(launches), host_ms=t_host, device_equals_host=bool(same(dev, host)))
        if ro.available() and boxes.shape[0] <= 600_000:
            ref, t_ref = timed(ro.generate_batches, boxes, 512, 8)
            r.update(reference_ms=t_ref, equals_reference=bool(same(dev, ref)))
        res[name] = r
        print(name, r, flush=True)
    for name in ("Castle", "Sponza"):
        if ro.available() and ro.scene_available(name):
            idx, verts = ro.load_mesh(name)
            ours, t_ours = timed(api.quad_decompose, idx, verts)
            ref, t_ref = timed(ro.quad_decompose, idx, verts)
            res[f"decompose_{name.lower()}"] = dict(triangles=int(idx.size // 3), quads=int(ours.size // 4), host_ms=t_ours, reference_ms=t_ref,
                                                    equal=bool(np.array_equal(ours, ref)))
            print(name, res[f"decompose_{name.lower()}"], flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/prep_timing.json", "w"), indent=1)


main()
