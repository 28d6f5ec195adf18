"""How the cluster kernel's warps wait for gate decisions (measurement build: make -C rasterizer_b200/csrc variant NAME=waitstats
DEFS=-DORZ_WAIT_STATS=1, then ORZ_LIB=.../variants/lib_waitstats.so python tools/wait_stats.py [case ...])."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rasterizer_b200 import api, camera as cam, workloads as wl
sys.argv = [sys.argv[0]] + (sys.argv[1:] or ["castle1024", "sponza256", "castle1", "probes1024"])
CASES = {"castle1024": ("castle", 1920, 1080, 1024, "path"), "sponza256": ("sponza", 1920, 1080, 256, "path"), "castle1": ("castle", 1920, 1080, 1, "default"),
         "sponza1": ("sponza", 1920, 1080, 1, "default"), "probes1024": ("castle", 512, 256, 1024, "probes")}
ctx = api.Context(0)
lib = api.lib()
lib.orz_debug_wait_stats.argtypes = [C.c_void_p]
scenes = {}
for name in sys.argv[1:]:
    sname, w, h, n, kind = CASES[name]
    if sname not in scenes:
        ps = wl.load_scene(sname); scenes[sname] = (ps, api.Scene.from_prepared(ctx, ps))
    ps, sc = scenes[sname]
    if kind == "default":
        c = ps.camera
        mvps = cam.view_projection(c["pos"], c["dir"], c["up"], c["fov"], w, h)[None].astype(np.float32); poss = np.array(c["pos"], np.float32)[None]
    else:
        mvps, poss = (wl.camera_path if kind == "path" else wl.probe_views)(ps, n, w, h)
    buf = (C.c_ulonglong * 8)()
    sc.render_views(w, h, mvps, cam_pos=poss, want=("vis",))
    lib.orz_debug_wait_stats(buf)   # warm-up run discarded
    out = sc.render_views(w, h, mvps, cam_pos=poss, want=("vis", "gate"))
    lib.orz_debug_wait_stats(buf)
    v = list(buf)
    n_occ = out["gate"].shape[1]
    print(json.dumps({"case": name, "views": n, "occluders": n_occ, "visible_per_view": float((out["gate"] & 1).sum() / n), "waits_per_view": v[0] / n,
                      "ended_visible_per_view": v[1] / n, "ended_invisible_per_view": v[2] / n, "decided_on_arrival_pct": 100.0 * v[3] / max(v[0], 1),
                      "mean_wait_us_visible": v[6] / max(v[1], 1) / 1965.0, "mean_wait_us_invisible": v[7] / max(v[2], 1) / 1965.0,
                      "warp_us_waited_per_view_visible": v[6] / n / 1965.0, "warp_us_waited_per_view_invisible": v[7] / n / 1965.0}))
