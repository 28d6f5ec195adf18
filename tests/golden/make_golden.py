"""Generates the committed golden vectors by running the UNMODIFIED reference (oracle/_ref, built
from /root/reference by oracle/Makefile) in the development container.

The reference's results depend on the host CPU's rcpps / rsqrtps, so the tables of
the generating host (Intel Xeon) are stored next to the vectors; the oracle port and the CUDA path
reproduce the vectors on ANY host once those tables are installed.
Usage: python tests/golden/make_golden.py   (needs oracle/_ref/libref_oracle.so)"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import port_oracle as po  # noqa: E402
from oracle import ref_oracle as ro  # noqa: E402
from rasterizer_b200 import camera as cam  # noqa: E402
from rasterizer_b200 import workloads as wl  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def views_for(ps, n, w, h):
    c = ps.camera
    m0 = cam.view_projection(c["pos"], c["dir"], c["up"], c["fov"], w, h)
    mv, pp = wl.camera_path(ps, n - 1, w, h)
    return np.concatenate([m0[None], mv]), np.concatenate([np.array(c["pos"], np.float32)[None], pp])


def capture(ps, s, w, h, n_views, keep_arrays):
    mvps, poss = views_for(ps, n_views, w, h)
    orders = wl.orders_for(s.centers, poss)
    boxes = ps.quad_boxes()
    r = ro.RefRasterizer(w, h)
    out = {"mvps": mvps, "poss": poss, "orders": orders}
    for v in range(n_views):
        gate, quads = r.frame(s, mvps[v], orders[v])
        q = r.query_boxes(boxes)
        out[f"gate{v}"] = gate
        out[f"quads{v}"] = np.array([quads], np.int64)
        out[f"vis{v}"] = np.packbits(q & 1, bitorder="little")
        out[f"clip{v}"] = np.packbits((q >> 1) & 1, bitorder="little")
        out[f"depth_sha{v}"] = np.frombuffer(bytes.fromhex(sha(r.depth())), np.uint8)
        out[f"hiz_sha{v}"] = np.frombuffer(bytes.fromhex(sha(r.hiz())), np.uint8)
        if keep_arrays:
            out[f"hiz{v}"] = r.hiz()
        out[f"image_sha{v}"] = np.frombuffer(bytes.fromhex(sha(r.readback())), np.uint8)
    r.close()
    return out


def main():
    np.save(os.path.join(HERE, "host_rcp_table.npy"), po.probe_host_rcp(11))
    np.save(os.path.join(HERE, "host_rsqrt_table.npy"), po.probe_host_rsqrt(10))
    np.save(os.path.join(HERE, "edge_mask_table.npy"), ro.RefRasterizer(64, 64).lut())
    # synthetic city: fully reproducible anywhere (scene generated from a seed)
    ps = wl.synthetic_city()
    s = ro.RefScene.from_batches(ps.batches, ps.ref_min, ps.ref_max)
    g = capture(ps, s, 640, 360, 4, keep_arrays=True)
    g["packed_sha"] = np.stack([np.frombuffer(bytes.fromhex(sha(s.packed(i))), np.uint8) for i in range(s.n_occluders)])
    np.savez_compressed(os.path.join(HERE, "city_640x360.npz"), **g)
    s.close()
    # near-clipped soup through rasterize<true>, no gate
    ps = wl.synthetic_soup(4096, cube=60.0)
    s = ro.RefScene.from_batches(ps.batches, ps.ref_min, ps.ref_max)
    c = ps.camera
    r = ro.RefRasterizer(640, 360)
    m = cam.view_projection(c["pos"], c["dir"], c["up"], c["fov"], 640, 360)
    order = cam.front_to_back_order(s.centers, c["pos"])
    r.submit_all(s, m, order, True)
    np.savez_compressed(os.path.join(HERE, "soup4096_640x360_clipped.npz"), mvp=m, order=order, hiz=r.hiz(),
                        depth_sha=np.frombuffer(bytes.fromhex(sha(r.depth())), np.uint8))
    r.close(); s.close()
    # the reference's own scenes (hashes only; the scene data is not redistributable in git)
    for name, size, n in (("Castle", (1920, 1080), 3), ("Castle", (512, 256), 4), ("Sponza", (1920, 1080), 2)):
        if not (ro.scene_available(name) and wl.have_scene(name.lower())):
            continue
        ps = wl.load_scene(name.lower())
        s = ro.RefScene.load(name)
        g = capture(ps, s, size[0], size[1], n, keep_arrays=False)
        np.savez_compressed(os.path.join(HERE, f"{name.lower()}_{size[0]}x{size[1]}.npz"), **g)
        s.close()
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
