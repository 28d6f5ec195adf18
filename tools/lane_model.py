"""CPU model of the cluster rasteriser's SIMD efficiency (round-2 planning), from the oracle
port's setup records -- no GPU needed.  For sampled views of a scene's camera path it runs the frame loop on
the port (gate decisions), sets up every quad of the visible occluders, and counts for every (occluder, tile of
8x4 blocks):
  records      primitives whose block rectangle meets the tile = warp steps of the tile-major walk today
  lanes        blocks of the tile inside those rectangles = useful lane-steps of the pixel work (upper bound:
               rectangle, not coverage)
  deepest      most rectangles stacked on one block of the tile = steps if every lane advanced on its own to
               its next record (section 9, first candidate)
  split_*      steps if the warp's 32 lanes were cut into 2 / 4 / 8 sub-tiles that each walk only the records
               meeting them, in lock step (steps = the busiest sub-tile's records)
usage: python tools/lane_model.py [castle|sponza|city] [views] [width height]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from oracle import port_oracle as po
from rasterizer_b200 import camera as cam, workloads as wl

TW, TH = 8, 4


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "castle"
    n_views = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    w, h = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (1920, 1080)
    ps = wl.load_scene(name)
    po.set_tables()
    baked = [po.bake(b, ps.ref_min, ps.ref_max) for b in ps.batches]
    packed = [b[0] for b in baked]
    centers, bmin, bmax = (np.stack([b[i] for b in baked]) for i in (1, 2, 3))
    mvps, poss = wl.camera_path(ps, 1024, w, h)
    pick = np.linspace(0, 1023, n_views).astype(int)
    port = po.PortRasterizer(w, h)
    tot = dict(records=0, lanes=0, deepest=0, tiles=0, prims=0, quads=0)
    splits = {"8x2": (8, 2), "4x4": (4, 4), "4x2": (4, 2), "2x2": (2, 2), "8x1": (8, 1)}
    split_steps = {k: 0 for k in splits}
    hist = np.zeros(33, np.int64)  # records by blocks-of-the-tile they span
    for v in pick:
        order = cam.front_to_back_order(centers, poss[v])
        gate, _ = port.frame(packed, bmin, bmax, ps.ref_min, ps.ref_max, mvps[v], order)
        for slot, o in enumerate(order):
            if not gate[slot] & 1:
                continue
            words = packed[o].reshape(-1, 4, 8)  # [group][vertex][quad in group]
            stack = {}   # tile -> per-block counts
            recs = {}    # tile -> records
            sub = {}     # tile -> {split: records per sub-tile}
            for g in range(words.shape[0]):
                for q in range(8):
                    p = port.setup_quad(words[g, :, q], ps.ref_min, ps.ref_max, bool(gate[slot] & 2))
                    tot["quads"] += 1
                    if p.mode == 0:
                        continue
                    tot["prims"] += 1
                    x0, y0, x1, y1 = p.minX, p.minY, p.minX + p.rangeX, p.minY + p.rangeY
                    for ty in range(y0 // TH, (y1 - 1) // TH + 1):
                        for tx in range(x0 // TW, (x1 - 1) // TW + 1):
                            cx0, cx1 = max(x0, tx * TW) - tx * TW, min(x1, tx * TW + TW) - tx * TW
                            cy0, cy1 = max(y0, ty * TH) - ty * TH, min(y1, ty * TH + TH) - ty * TH
                            key = (tx, ty)
                            if key not in stack:
                                stack[key] = np.zeros((TH, TW), np.int32)
                                recs[key] = 0
                                sub[key] = {k: np.zeros((TH // sh, TW // sw), np.int32) for k, (sw, sh) in splits.items()}
                            stack[key][cy0:cy1, cx0:cx1] += 1
                            recs[key] += 1
                            for k, (sw, sh) in splits.items():
                                sub[key][k][cy0 // sh:(cy1 - 1) // sh + 1, cx0 // sw:(cx1 - 1) // sw + 1] += 1
                            blocks = (cx1 - cx0) * (cy1 - cy0)
                            tot["lanes"] += blocks
                            hist[blocks] += 1
            for key, s in stack.items():
                tot["tiles"] += 1
                tot["records"] += recs[key]
                tot["deepest"] += int(s.max())
                for k in splits:
                    split_steps[k] += int(sub[key][k].max())
    port.close()
    res = dict(scene=name, views=int(n_views), width=w, height=h, **{k: int(x) for k, x in tot.items()})
    res["lanes_per_record"] = tot["lanes"] / max(tot["records"], 1)
    res["records_per_tile_visit"] = tot["records"] / max(tot["tiles"], 1)
    res["steps_today_over_per_lane_advance"] = tot["records"] / max(tot["deepest"], 1)
    res["steps_today_over_split"] = {k: tot["records"] / max(x, 1) for k, x in split_steps.items()}
    res["records_by_blocks_spanned"] = {str(i): int(c) for i, c in enumerate(hist) if c}
    print(json.dumps(res, indent=1))
    os.makedirs("profiles", exist_ok=True)
    json.dump(res, open(f"profiles/r1_lane_model_{name}_{w}x{h}.json", "w"), indent=1)


main()
