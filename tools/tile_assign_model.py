"""CPU model of tile -> warp assignments for the cluster rasteriser (cost-aware tile ownership).
For sampled views of a camera path it computes, from the oracle port's setup records,
  true[t]   records of the occluders the gate lets through that meet tile t   (what the walk really costs)
  guess[t]  records of EVERY occluder in the frustum that meet tile t         (what k_setup_views knows before the walk)
and prices the busiest warp over the mean for: the fixed map t mod W, a serial greedy (largest tile to the least
loaded warp, at most K tiles per warp) and the ROUND-MATCHED deal the GPU uses (k_assign_tiles: tiles sorted by falling
guess; in round r the next W tiles go one to each warp, the heaviest tile to the warp with the smallest load so far).
usage: python tools/tile_assign_model.py [castle|sponza] [views] [width height]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from oracle import port_oracle as po
from rasterizer_b200 import camera as cam, workloads as wl

TW, TH = 8, 4


def tile_costs(name, n_views, w, h):
    ps = wl.load_scene(name)
    po.set_tables()
    baked = [po.bake(b, ps.ref_min, ps.ref_max) for b in ps.batches]
    packed = [b[0] for b in baked]
    centers, bmin, bmax = (np.stack([b[i] for b in baked]) for i in (1, 2, 3))
    mvps, poss = wl.camera_path(ps, 1024, w, h)
    pick = np.linspace(0, 1023, n_views).astype(int)
    port, empty = po.PortRasterizer(w, h), po.PortRasterizer(w, h)
    tiles_x, tiles_y = (w // 8 + TW - 1) // TW, (h // 8 + TH - 1) // TH
    out = []
    for v in pick:
        true = np.zeros(tiles_x * tiles_y, np.int64)
        guess = np.zeros(tiles_x * tiles_y, np.int64)
        true_b = np.zeros(tiles_x * tiles_y, np.int64)   # with the block weight: 12 + blocks of the rectangle on the tile
        guess_b = np.zeros(tiles_x * tiles_y, np.int64)
        order = cam.front_to_back_order(centers, poss[v])
        gate, _ = port.frame(packed, bmin, bmax, ps.ref_min, ps.ref_max, mvps[v], order)
        empty.clear()
        empty.set_mvp(mvps[v])
        for slot, o in enumerate(order):
            g0 = empty.query(bmin[o], bmax[o])
            if not g0 & 1:
                continue
            vis = bool(gate[slot] & 1)
            wd = packed[o].reshape(-1, 4, 8)
            for g in range(wd.shape[0]):
                for q in range(8):
                    p = empty.setup_quad(wd[g, :, q], ps.ref_min, ps.ref_max, bool(g0 & 2))
                    if not p.mode:
                        continue
                    x0, y0, x1, y1 = p.minX, p.minY, p.minX + p.rangeX, p.minY + p.rangeY
                    for ty in range(y0 // TH, min((y1 - 1) // TH + 1, tiles_y)):
                        for tx in range(x0 // TW, min((x1 - 1) // TW + 1, tiles_x)):
                            t = ty * tiles_x + tx
                            blocks = (min(x1, tx * TW + TW) - max(x0, tx * TW)) * (min(y1, ty * TH + TH) - max(y0, ty * TH))
                            guess[t] += 1
                            guess_b[t] += 12 + blocks
                            if vis:
                                true[t] += 1
                                true_b[t] += 12 + blocks
        out.append((true, guess, true_b, guess_b))
    port.close()
    empty.close()
    return out, tiles_x * tiles_y


def ratio(per_warp):
    return float(per_warp.max() / max(per_warp.mean(), 1e-9))


def fixed(true, W):
    return ratio(np.bincount(np.arange(true.size) % W, weights=true, minlength=W))


def greedy(true, guess, W, K):
    bins, tb, fill = np.zeros(W), np.zeros(W), np.zeros(W, np.int64)
    for t in np.argsort(-guess, kind="stable"):
        b = int(np.argmin(np.where(fill < K, bins, np.inf)))
        bins[b] += guess[t]; tb[b] += true[t]; fill[b] += 1
    return ratio(tb)


def round_matched(true, guess, W):
    rank = np.argsort(-guess, kind="stable")
    bins, tb = np.zeros(W), np.zeros(W)
    for r0 in range(0, rank.size, W):
        tiles = rank[r0:r0 + W]                      # falling cost
        warps = np.argsort(bins, kind="stable")      # rising load
        for t, b in zip(tiles, warps):
            bins[b] += guess[t]; tb[b] += true[t]
    return ratio(tb)


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "castle"
    n_views = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    w, h = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (1920, 1080)
    costs, n_tiles = tile_costs(name, n_views, w, h)
    res = dict(scene=name, views=n_views, width=w, height=h, tiles=n_tiles, busiest_warp_over_mean={})
    for C in (1, 2, 4, 16):
        W = 16 * C
        K = (n_tiles + W - 1) // W
        rows = {}
        for label, f in (("t mod W (fixed)", lambda tr, gu, tb, gb: fixed(tr, W)),
                         ("greedy, speculative records", lambda tr, gu, tb, gb: greedy(tr, gu, W, K)),
                         ("round matched, speculative records", lambda tr, gu, tb, gb: round_matched(tr, gu, W)),
                         ("round matched, speculative 12 + blocks (priced by 12 + blocks)", lambda tr, gu, tb, gb: round_matched(tb, gb, W)),
                         ("fixed (priced by 12 + blocks)", lambda tr, gu, tb, gb: fixed(tb, W)),
                         ("round matched, true records", lambda tr, gu, tb, gb: round_matched(tr, tr, W))):
            x = [f(*c) for c in costs]
            rows[label] = dict(mean=float(np.mean(x)), worst=float(np.max(x)))
        res["busiest_warp_over_mean"][f"C={C} ({W} warps, {K} tiles each)"] = rows
    print(json.dumps(res, indent=1))
    os.makedirs("profiles", exist_ok=True)
    json.dump(res, open(f"profiles/r2_tile_assign_model_{name}_{w}x{h}.json", "w"), indent=1)


main()
