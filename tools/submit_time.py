import os, sys, time, json
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from rasterizer_b200 import api, workloads as wl
dev = torch.device("cuda", 0)
ctx = api.Context(0)
ps = wl.load_scene("castle")
scene = api.Scene.from_prepared(ctx, ps)
w, h, n = 512, 256, 1024
mvps, poss = wl.probe_views(ps, n * 8, w, h)
mvps, poss = np.ascontiguousarray(mvps[::8]), np.ascontiguousarray(poss[::8])
words = (scene.n_boxes + 31) // 32
d_mvps, d_pos = torch.from_numpy(mvps).to(dev), torch.from_numpy(poss).to(dev)
d_vis = torch.zeros((n, words), dtype=torch.int32, device=dev)
b = api.ViewBatch()
b.width, b.height, b.nViews, b.flags = w, h, n, 0
b.mvps, b.camPos, b.visBits = d_mvps.data_ptr(), d_pos.data_ptr(), d_vis.data_ptr()
for _ in range(5): scene.render_views_raw(b, device=True)
torch.cuda.synchronize()
ts = []
for _ in range(50):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); scene.render_views_raw(b, device=True); t1 = time.perf_counter()
    ts.append((t1 - t0) * 1e3)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200): scene.render_views_raw(b, device=True)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(json.dumps({"submit_ms_median": sorted(ts)[25], "submit_ms_min": min(ts), "back_to_back_submit_ms": (t1 - t0) / 200 * 1e3, "back_to_back_total_ms": (t2 - t0) / 200 * 1e3, "cpus": os.cpu_count()}))
# per-step event times against the wall clock of the whole queue, for several queue depths
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
for reps in (10, 50, 200):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for a, z in ev:
        a.record(stream); scene.render_views_raw(b, device=True); z.record(stream)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    ms = [a.elapsed_time(z) for a, z in ev]
    gaps = [ev[i][1].elapsed_time(ev[i + 1][0]) for i in range(reps - 1)]
    print(json.dumps({"reps": reps, "step_ms_median": sorted(ms)[reps // 2], "step_ms_max": max(ms), "gap_ms_median": sorted(gaps)[len(gaps) // 2], "gap_ms_max": max(gaps),
                      "first_to_last_ms_per_step": ev[0][0].elapsed_time(ev[-1][1]) / reps, "wall_ms_per_step": (t1 - t0) * 1e3 / reps}))
