"""Occluder::bake for a whole scene: host bake (orz_bake, one thread) against the device bake
(orz_scene_bake: H2D of the vertices + one CTA per batch + D2H of packets/bounds), same output."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rasterizer_b200 import api, workloads as wl

def main():
    ctx = api.Context(0)
    res = {}
    for name in ("castle", "sponza"):
        if not wl.have_scene(name): continue
        ps = wl.load_scene(name)
        t0 = time.perf_counter()
        host = [api.bake(b, ps.ref_min, ps.ref_max) for b in ps.batches]
        t_host = time.perf_counter() - t0
        sd = api.Scene.bake_on_device(ctx, ps.batches, ps.ref_min, ps.ref_max); sd.close()   # warm-up (rsqrt probe, module load)
        t0 = time.perf_counter()
        sd = api.Scene.bake_on_device(ctx, ps.batches, ps.ref_min, ps.ref_max)
        t_dev = time.perf_counter() - t0
        same = all(np.array_equal(sd.packed_list[i], host[i][0]) and np.array_equal(sd.bounds_min[i], host[i][2]) for i in range(len(host)))
        res[name] = dict(batches=len(host), quads=int(sum(len(b) for b in ps.batches) // 4), host_bake_ms=t_host * 1e3, device_bake_ms=t_dev * 1e3, bit_exact=bool(same))
        print(name, res[name], flush=True)
        sd.close()
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/bake_timing.json", "w"), indent=1)
main()
