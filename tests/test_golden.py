"""Golden vectors produced by the UNMODIFIED reference (tests/golden/make_golden.py): the oracle
port (CPU) and the CUDA path (GPU) must reproduce them with the generating host's rcpps / rsqrtps
tables installed, whatever CPU the test runs on."""
import hashlib
import os

import numpy as np
import pytest

from oracle import port_oracle as po
from rasterizer_b200 import api
from rasterizer_b200 import camera as cam
from rasterizer_b200 import workloads as wl

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RCP = np.load(os.path.join(G, "host_rcp_table.npy"))
RSQRT = np.load(os.path.join(G, "host_rsqrt_table.npy"))
LUT = np.load(os.path.join(G, "edge_mask_table.npy"))


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


def scene_for(key):
    name = key.split("_")[0]
    if name == "city":
        return wl.synthetic_city()
    if not wl.have_scene(name):
        pytest.skip(f"prepared scene {name} not present")
    return wl.load_scene(name)


CASES = ["city_640x360", "castle_1920x1080", "castle_512x256", "sponza_1920x1080"]


def test_edge_mask_table_golden():
    assert np.array_equal(po.build_lut(), LUT)
    assert np.array_equal(api.edge_mask_table(), LUT)


@pytest.mark.parametrize("case", CASES)
def test_oracle_port_reproduces_reference_vectors(case):
    g = np.load(os.path.join(G, case + ".npz"))
    ps = scene_for(case)
    w, h = (int(x) for x in case.split("_")[1].split("x"))
    po.set_tables(RCP, RSQRT)
    try:
        baked = [po.bake(b, ps.ref_min, ps.ref_max) for b in ps.batches]
        if "packed_sha" in g:
            assert all(np.array_equal(sha(b[0]), g["packed_sha"][i]) for i, b in enumerate(baked))
        packed = [b[0] for b in baked]
        bmin, bmax = np.stack([b[2] for b in baked]), np.stack([b[3] for b in baked])
        boxes = ps.quad_boxes()
        port = po.PortRasterizer(w, h, LUT)
        n_views = g["mvps"].shape[0] if case.startswith("city") else 2  # the port is slow on 130k-box scenes
        for v in range(n_views):
            gate, quads = port.frame(packed, bmin, bmax, ps.ref_min, ps.ref_max, g["mvps"][v], g["orders"][v])
            assert np.array_equal(gate, g[f"gate{v}"]) and quads == int(g[f"quads{v}"][0])
            assert np.array_equal(sha(port.hiz()), g[f"hiz_sha{v}"])
            assert np.array_equal(sha(port.depth()), g[f"depth_sha{v}"])
            if f"hiz{v}" in g:
                assert np.array_equal(port.hiz(), g[f"hiz{v}"])
            if len(boxes) <= 30000:
                q = port.query_boxes(boxes)
                assert np.array_equal(np.packbits(q & 1, bitorder="little"), g[f"vis{v}"])
                assert np.array_equal(np.packbits((q >> 1) & 1, bitorder="little"), g[f"clip{v}"])
            assert np.array_equal(sha(port.readback()), g[f"image_sha{v}"])
        port.close()
    finally:
        po.set_tables()


def test_oracle_port_soup_clipped_golden():
    g = np.load(os.path.join(G, "soup4096_640x360_clipped.npz"))
    ps = wl.synthetic_soup(4096, cube=60.0)
    po.set_tables(RCP, RSQRT)
    try:
        port = po.PortRasterizer(640, 360, LUT)
        port.clear(); port.set_mvp(g["mvp"])
        baked = [po.bake(b, ps.ref_min, ps.ref_max)[0] for b in ps.batches]
        for o in g["order"]:
            port.rasterize(baked[o], ps.ref_min, ps.ref_max, True)
        assert np.array_equal(port.hiz(), g["hiz"])
        assert np.array_equal(sha(port.depth()), g["depth_sha"])
        port.close()
    finally:
        po.set_tables()


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_cuda_path_reproduces_reference_vectors(case):
    g = np.load(os.path.join(G, case + ".npz"))
    ps = scene_for(case)
    w, h = (int(x) for x in case.split("_")[1].split("x"))
    ctx = api.Context(0)
    ctx.set_rcp_table(RCP)
    api.set_rsqrt_table(RSQRT)
    try:
        sc = api.Scene.from_prepared(ctx, ps)
        if "packed_sha" in g:
            assert all(np.array_equal(sha(p), g["packed_sha"][i]) for i, p in enumerate(sc.packed_list))
        n = g["mvps"].shape[0]
        out = sc.render_views(w, h, g["mvps"], orders=g["orders"], want=("vis", "clip", "gate", "depth", "hiz", "quads"))
        out2 = sc.render_views(w, h, g["mvps"], cam_pos=g["poss"], want=("vis", "gate"))  # order computed on the GPU
        nb = sc.n_boxes
        for v in range(n):
            assert np.array_equal(out["gate"][v], g[f"gate{v}"]) and out["quads"][v] == int(g[f"quads{v}"][0])
            assert np.array_equal(out2["gate"][v], g[f"gate{v}"])
            assert np.array_equal(sha(out["hiz"][v]), g[f"hiz_sha{v}"])
            assert np.array_equal(sha(out["depth"][v]), g[f"depth_sha{v}"])
            vis = np.packbits(api.unpack_bits(out["vis"][v:v + 1], nb)[0], bitorder="little")
            clip = np.packbits(api.unpack_bits(out["clip"][v:v + 1], nb)[0], bitorder="little")
            assert np.array_equal(vis, g[f"vis{v}"])
            assert np.array_equal(clip, g[f"clip{v}"])
            assert np.array_equal(out2["vis"][v], out["vis"][v])
        sc.close()
    finally:
        api.set_rsqrt_table(None)
        ctx.close()
