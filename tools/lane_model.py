"""CPU model of the cluster rasteriser's SIMD efficiency (round-2 planning), from the oracle
port's setup records -- no GPU needed.  For sampled views of a scene's camera path it runs the frame loop on
the port (gate decisions), sets up every quad of the visible occluders, and counts for every (occluder, tile of
8x4 blocks):
  records      primitives whose block rectangle meets the tile = warp steps of the tile-major walk today
  lanes        blocks of the tile inside those rectangles = useful lane-steps of the pixel work (upper bound:
               rectangle, not coverage)
  deepest      most rectangles stacked on one block of the tile = steps if every lane advanced on its own to
               its next record (section 9, first candidate)
  imbalance    tile t belongs to warp t mod (16 C) for the whole view: records per warp, busiest warp over mean,
               for the cluster sizes the batch (C = 2) and the single-view (C = 16) paths use
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
    tiles_x, tiles_y = (w // 8 + TW - 1) // TW, (h // 8 + TH - 1) // TH
    imbalance = {2: [], 16: []}
    patterns = {}   # alternative static tile -> warp maps at C = 2, and the greedy bound with the true costs
    rng = np.random.default_rng(1)
    t_all = np.arange(tiles_x * tiles_y)
    tx_all, ty_all = t_all % tiles_x, t_all // tiles_x
    maps = {"t mod 32 (today)": t_all % 32, "(tx + 5 ty) mod 32": (tx_all + 5 * ty_all) % 32, "(3 tx + 7 ty) mod 32": (3 * tx_all + 7 * ty_all) % 32,
            "(tx + 8 ty) mod 32": (tx_all + 8 * ty_all) % 32,
            "random": rng.permutation(t_all.size) % 32}
    for k in list(maps) + ["greedy by true cost", "greedy by speculative cost (all occluders in the frustum)",
                           "snake deal by speculative cost", "snake deal by true cost"]:
        patterns[k] = []
    empty = po.PortRasterizer(w, h)  # never drawn into: its gate answers "in the frustum" only
    for v in pick:
        tile_work = np.zeros(tiles_x * tiles_y, np.int64)
        tile_guess = np.zeros(tiles_x * tiles_y, np.int64)
        order = cam.front_to_back_order(centers, poss[v])
        gate, _ = port.frame(packed, bmin, bmax, ps.ref_min, ps.ref_max, mvps[v], order)
        empty.clear()
        empty.set_mvp(mvps[v])
        for slot, o in enumerate(order):
            if not gate[slot] & 1:
                g0 = empty.query(bmin[o], bmax[o])
                if g0 & 1:  # set up speculatively by k_setup_views, rejected later by the gate: costs nothing in the walk
                    wd = packed[o].reshape(-1, 4, 8)
                    for g in range(wd.shape[0]):
                        for q in range(8):
                            p = empty.setup_quad(wd[g, :, q], ps.ref_min, ps.ref_max, bool(g0 & 2))
                            if p.mode:
                                for ty in range(p.minY // TH, min((p.minY + p.rangeY - 1) // TH + 1, tiles_y)):
                                    for tx in range(p.minX // TW, min((p.minX + p.rangeX - 1) // TW + 1, tiles_x)):
                                        tile_guess[ty * tiles_x + tx] += 1
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
                if key[0] < tiles_x and key[1] < tiles_y:
                    tile_work[key[1] * tiles_x + key[0]] += recs[key]
                tot["deepest"] += int(s.max())
                for k in splits:
                    split_steps[k] += int(sub[key][k].max())
        for k, m in maps.items():
            per_warp = np.bincount(m, weights=tile_work, minlength=32)
            patterns[k].append(float(per_warp.max() / max(per_warp.mean(), 1e-9)))
        bins = np.zeros(32)
        for cost in np.sort(tile_work)[::-1]:
            bins[np.argmin(bins)] += cost
        patterns["greedy by true cost"].append(float(bins.max() / max(bins.mean(), 1e-9)))
        tile_guess += tile_work  # what the setup kernel knows: records of every occluder in the frustum
        bins, true_bins, fill = np.zeros(32), np.zeros(32), np.zeros(32, np.int64)
        for t in np.argsort(-tile_guess, kind="stable"):
            open_bins = np.where(fill < 32, bins, np.inf)  # a warp keeps at most 32 tiles
            b = int(np.argmin(open_bins))
            bins[b] += tile_guess[t]
            true_bins[b] += tile_work[t]
            fill[b] += 1
        patterns["greedy by speculative cost (all occluders in the frustum)"].append(float(true_bins.max() / max(true_bins.mean(), 1e-9)))
        for label, costs in (("snake deal by speculative cost", tile_guess), ("snake deal by true cost", tile_work)):
            rank = np.argsort(-costs, kind="stable")      # tiles by falling cost, dealt 0..31, 31..0, 0..31, ...
            pos = np.arange(rank.size)
            warp_of = np.where((pos // 32) % 2 == 0, pos % 32, 31 - pos % 32)
            per_warp = np.bincount(warp_of, weights=tile_work[rank], minlength=32)
            patterns[label].append(float(per_warp.max() / max(per_warp.mean(), 1e-9)))
        for c in imbalance:
            warps = 16 * c
            per_warp = np.bincount(np.arange(tile_work.size) % warps, weights=tile_work, minlength=warps)
            imbalance[c].append(float(per_warp.max() / max(per_warp.mean(), 1e-9)))
    port.close()
    empty.close()
    res = dict(scene=name, views=int(n_views), width=w, height=h, **{k: int(x) for k, x in tot.items()})
    res["lanes_per_record"] = tot["lanes"] / max(tot["records"], 1)
    res["records_per_tile_visit"] = tot["records"] / max(tot["tiles"], 1)
    res["steps_today_over_per_lane_advance"] = tot["records"] / max(tot["deepest"], 1)
    res["steps_today_over_split"] = {k: tot["records"] / max(x, 1) for k, x in split_steps.items()}
    res["busiest_warp_over_mean"] = {f"C={c}": dict(mean=float(np.mean(x)), worst=float(np.max(x))) for c, x in imbalance.items()}
    res["busiest_warp_over_mean_by_tile_map_C2"] = {k: dict(mean=float(np.mean(x)), worst=float(np.max(x))) for k, x in patterns.items()}
    res["records_by_blocks_spanned"] = {str(i): int(c) for i, c in enumerate(hist) if c}
    print(json.dumps(res, indent=1))
    os.makedirs("profiles", exist_ok=True)
    json.dump(res, open(f"profiles/r1_lane_model_{name}_{w}x{h}.json", "w"), indent=1)


main()
