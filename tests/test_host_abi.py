"""CPU-side checks of the product's host logic and of the C-ABI boundary: the library loads and
exports every symbol include/orz.h declares, host bake / edge-mask table / rcpps probe agree with
the oracle, compute entry points fail loudly without a device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from oracle import port_oracle as po
from rasterizer_b200 import api
from rasterizer_b200 import workloads as wl

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "orz.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(orz_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 30
    lib = C.CDLL(api.LIB_PATH)
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing
    assert names == set(api.EXPORTS), (names ^ set(api.EXPORTS))


def test_no_cpu_fallback():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    with pytest.raises(api.OrzError, match="no CUDA device"):
        api.Context(0)


def test_edge_mask_table_and_rcp_probe_match_oracle():
    assert np.array_equal(api.edge_mask_table(), po.build_lut())
    table, bits, exact = api.probe_host_rcp()
    assert exact, "the rcpps exponent/special-value model does not hold on this CPU"
    assert np.array_equal(table, po.probe_host_rcp(bits))


def test_host_bake_matches_oracle():
    po.set_tables()
    scenes = [wl.synthetic_city(), wl.synthetic_soup(1024, cube=40.0)]
    if wl.have_scene("castle"):
        scenes.append(wl.load_scene("castle"))
    for ps in scenes:
        for b in ps.batches[:: max(1, len(ps.batches) // 25)]:
            got, want = api.bake(b, ps.ref_min, ps.ref_max), po.bake(b, ps.ref_min, ps.ref_max)
            for g, w in zip(got, want):
                assert np.array_equal(g.view(np.uint32), w.view(np.uint32))
    # a batch whose padding quads are collapsed (NaN normals, Main.cpp:91-94) keeps the same order
    b = wl.synthetic_city().batches[-1]
    assert np.array_equal(api.bake(b, scenes[0].ref_min, scenes[0].ref_max)[0], po.bake(b, scenes[0].ref_min, scenes[0].ref_max)[0])


def test_table_rsqrt_bake_is_host_independent():
    """With the rsqrtps table installed the bake no longer depends on the CPU it runs on."""
    ps = wl.synthetic_city()
    t = po.probe_host_rsqrt(10)
    want = [api.bake(b, ps.ref_min, ps.ref_max)[0] for b in ps.batches[:4]]
    api.set_rsqrt_table(t)
    try:
        got = [api.bake(b, ps.ref_min, ps.ref_max)[0] for b in ps.batches[:4]]
    finally:
        api.set_rsqrt_table(None)
    assert all(np.array_equal(a, b) for a, b in zip(got, want))


def test_dropin_cpp_program_compiles_and_links():
    out = os.path.join(ROOT, "tests", "_build", "dropin_frame")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-mavx2", "-Wno-ignored-attributes", "-I", os.path.join(ROOT, "rasterizer_b200", "csrc", "dropin"),
                           "-o", out, os.path.join(ROOT, "tests", "dropin_frame.cpp"), "-L", os.path.join(ROOT, "rasterizer_b200"),
                           "-lrasterizer_b200", "-Wl,-rpath," + os.path.join(ROOT, "rasterizer_b200")])
    assert os.path.exists(out)
