"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck): a few views of the
synthetic city through every kernel of the view-batch path and the per-call path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rasterizer_b200 import api, workloads as wl, camera as cam

ps = wl.synthetic_city(n_blocks=5)
ctx = api.Context(0)
sc = api.Scene.from_prepared(ctx, ps)
w, h = 320, 184
mvps, poss = wl.camera_path(ps, 6, w, h)
ctx.set_cluster_views(0)      # the one-CTA-per-view batch kernel
for gw in (1, 4, 8):
    ctx.set_group_warps(gw)
    out = sc.render_views(w, h, mvps, cam_pos=poss, want=("vis", "clip", "gate", "depth", "hiz", "quads"))
ctx.set_group_warps(0)
ctx.set_cluster_views(1024)   # the cluster path (speculative setup + dataflow cluster kernel), every cluster size
for cs in (0, 1, 2, 4, 8, 16):
    ctx.set_cluster_size(cs)
    out2 = sc.render_views(w, h, mvps, cam_pos=poss, want=("vis", "clip", "gate", "depth", "hiz", "quads"))
    assert all(np.array_equal(out[k], out2[k]) for k in out), cs
ctx.set_cluster_size(0)
sd = api.Scene.bake_on_device(ctx, ps.batches, ps.ref_min, ps.ref_max)   # device bake
sd.close()
out = sc.render_views(w, h, mvps[:2], cam_pos=poss[:2], flags=api.BATCH_NO_GATE | api.BATCH_FORCE_CLIPPED | api.BATCH_WIDE, want=("vis", "depth", "hiz"))
ctx.set_cluster_views(api.DEFAULT_CLUSTER_VIEWS)
ctx.set_tile_height(1, 1)     # 8 x 1 strips on the cluster path
out3 = sc.render_views(w, h, mvps, cam_pos=poss, want=("vis", "clip", "gate", "depth", "hiz", "quads"))
assert all(np.array_equal(out2[k], out3[k]) for k in out2)
ctx.set_tile_height(0, 1)
out = sc.render_views(w, h, mvps[:2], cam_pos=poss[:2], flags=api.BATCH_NO_GATE | api.BATCH_FORCE_CLIPPED, want=("vis", "depth", "hiz"))
# per-call path: the frame loop twice (the second frame runs on predicted query chains), both tile heights
r = api.Rasterizer(ctx, w, h)
n_occ = min(12, len(sc.packed_list))
occs = [api.Occluder(ctx, p, ps.ref_min, ps.ref_max) for p in sc.packed_list[:n_occ]]
for tile_h in (1, 4):
    ctx.set_tile_height(0, tile_h)
    for frame in range(2):
        r.clear(); r.setModelViewProjection(mvps[frame % len(mvps)])
        for i, o in enumerate(occs):
            vis, clip = r.queryVisibility(sc.bounds_min[i], sc.bounds_max[i])
            if vis:
                r.rasterize(o, clip)
ctx.set_tile_height(0, 1)
r.query_boxes(ps.quad_boxes()[:500]); r.readBackDepth(); r.download()
print("sanitize workload done", int(out["hiz"].sum()))
