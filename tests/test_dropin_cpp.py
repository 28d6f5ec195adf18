"""The C++ drop-in classes (rasterizer_b200/csrc/dropin: same names and signatures as the
reference's Occluder / Rasterizer) driven by an application-style program, checked against the oracle."""
import os
import subprocess

import numpy as np
import pytest

from oracle import port_oracle as po
from rasterizer_b200 import api
from rasterizer_b200 import camera as cam
from rasterizer_b200 import workloads as wl

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name,size", [("city", (640, 360)), ("castle", (1920, 1080))])
def test_cpp_frame_loop(tmp_path, name, size):
    if name != "city" and not wl.have_scene(name):
        pytest.skip("prepared scene missing")
    exe = os.path.join(ROOT, "tests", "_build", "dropin_frame")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-mavx2", "-Wno-ignored-attributes", "-I", os.path.join(ROOT, "rasterizer_b200", "csrc", "dropin"),
                           "-o", exe, os.path.join(ROOT, "tests", "dropin_frame.cpp"), "-L", os.path.join(ROOT, "rasterizer_b200"),
                           "-lrasterizer_b200", "-Wl,-rpath," + os.path.join(ROOT, "rasterizer_b200")])
    ps = wl.load_scene(name)
    w, h = size
    scene_file = wl.prepared_path(name) if name != "city" else str(tmp_path / "city.orzscn")
    if name == "city":
        ps.save(scene_file)
    po.set_tables()
    baked = [po.bake(b, ps.ref_min, ps.ref_max) for b in ps.batches]
    c = ps.camera
    mvp = cam.view_projection(c["pos"], c["dir"], c["up"], c["fov"], w, h)
    order = cam.front_to_back_order(np.stack([b[1] for b in baked]), c["pos"])
    mvp.tofile(tmp_path / "mvp.bin"); order.tofile(tmp_path / "order.bin")
    log = subprocess.check_output([exe, scene_file, str(w), str(h), str(tmp_path / "mvp.bin"), str(tmp_path / "order.bin"), str(tmp_path / "out.bin")],
                                  env=dict(os.environ, ORZ_FRAME_REPS="50"), text=True)
    print(log)  # frame_ms: the per-call frame loop as an unchanged application runs it
    raw = np.fromfile(tmp_path / "out.bin", np.uint8)
    n, blocks = len(order), (w // 8) * (h // 8)
    gate = raw[:n]
    hiz = raw[n:n + 2 * blocks].view(np.uint16)
    depth = raw[n + 2 * blocks:n + 2 * blocks + 2 * w * h].view(np.uint16)
    image = raw[n + 2 * blocks + 2 * w * h:]
    port = po.PortRasterizer(w, h)
    want_gate, _ = port.frame([b[0] for b in baked], np.stack([b[2] for b in baked]), np.stack([b[3] for b in baked]), ps.ref_min, ps.ref_max, mvp, order)
    assert np.array_equal(gate, want_gate)
    assert np.array_equal(hiz, port.hiz())
    assert np.array_equal(depth, port.depth())
    assert np.array_equal(image, port.readback())
    port.close()


def test_cpp_preparation_on_the_gpu(tmp_path):
    """tests/dropin_prepare.cpp (Main.cpp:86-128 against the reference's class names): with the drop-in
    SurfaceAreaHeuristic batching on the GPU the baked scene equals the one batched on the host."""
    from oracle import ref_oracle as ro

    if not ro.scene_available("Castle"):
        pytest.skip("no Castle data under oracle/_ref/scenes")
    exe = os.path.join(ROOT, "tests", "_build", "dropin_prepare")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-mavx2", "-msse4.1", "-Wno-ignored-attributes", "-I", os.path.join(ROOT, "rasterizer_b200", "csrc", "dropin"),
                           "-o", exe, os.path.join(ROOT, "tests", "dropin_prepare.cpp"), "-L", os.path.join(ROOT, "rasterizer_b200"),
                           "-lrasterizer_b200", "-Wl,-rpath," + os.path.join(ROOT, "rasterizer_b200")])
    d = os.path.join(ro.SCENE_DIR, "Castle")
    files = [os.path.join(d, "IndexBuffer.bin"), os.path.join(d, "VertexBuffer.bin")]
    subprocess.check_call([exe, *files, str(tmp_path / "host.bin")], env=dict(os.environ, ORZ_PREP_ON_HOST="1"))
    subprocess.check_call([exe, *files, str(tmp_path / "gpu.bin")], env={k: v for k, v in os.environ.items() if k != "ORZ_PREP_ON_HOST"})
    host, gpu = np.fromfile(tmp_path / "host.bin", np.uint32), np.fromfile(tmp_path / "gpu.bin", np.uint32)
    assert host[0] == 76 and np.array_equal(host, gpu)
