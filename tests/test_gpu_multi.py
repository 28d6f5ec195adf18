"""Multi-GPU view-batch path on real devices, both kinds of caller:
  * one process per GPU (torchrun, NCCL): every rank renders its share of one camera path and the per-view visibility
    bitmasks are gathered -- through torch.distributed (host-buffer flavour) and through the product's own C-ABI
    communicator (orz_comm_create / orz_gather_bits[_overlapped]); contiguous and round-robin dealing;
  * one C++ process driving every visible GPU through include/orz.h alone (tests/comm_gather.cpp, orz_comm_create_all).
The gathered result must equal the single-GPU run bit for bit.  The torchrun test needs two GPUs; the C++ program
runs on however many are visible (one included: a communicator of one rank)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from rasterizer_b200 import api, workloads as wl, distributed as D
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ps = wl.synthetic_city()
ctx = api.Context(local)
sc = api.Scene.from_prepared(ctx, ps)
mvps, poss = wl.camera_path(ps, 37, 640, 360)
dev = torch.device("cuda", local)
res = {}
res["contiguous"] = D.render_views_sharded(sc, 640, 360, mvps, cam_pos=poss, device=dev)
res["interleaved"] = D.render_views_sharded(sc, 640, 360, mvps, cam_pos=poss, device=dev, balance="interleaved")
comm = D.make_comm(ctx)
res["cabi"] = D.render_views_gathered(sc, comm, 640, 360, mvps, poss)
res["cabi_overlapped"] = D.render_views_gathered(sc, comm, 640, 360, mvps, poss, overlapped=True)
# every rank must hold the same gathered buffer
for k, v in res.items():
    mine = v.to(dev).contiguous()
    ref = mine.clone()
    dist.broadcast(ref, src=0)
    assert torch.equal(mine, ref), (k, rank)
if rank == 0:
    np.savez(sys.argv[2], **{k: v.cpu().numpy().view(np.uint32) for k, v in res.items()})
dist.barrier()
comm.close(); sc.close(); ctx.close()
dist.destroy_process_group()
'''


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    return port


def test_two_gpu_gather_equals_single_gpu(tmp_path):
    import torch

    n_gpus = torch.cuda.device_count()
    if n_gpus < 2:
        pytest.skip("needs two GPUs")
    from rasterizer_b200 import api, workloads as wl

    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    ps = wl.synthetic_city()
    ctx = api.Context(0)
    sc = api.Scene.from_prepared(ctx, ps)
    mvps, poss = wl.camera_path(ps, 37, 640, 360)
    want = sc.render_views(640, 360, mvps, cam_pos=poss, want=("vis",))["vis"]
    sc.close(); ctx.close()
    for world in sorted({2, n_gpus}):
        out = tmp_path / f"bits{world}.npz"
        subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
                               "--master-port", str(_free_port()), str(script), ROOT, str(out)], timeout=900)
        got = np.load(out)
        for k in ("contiguous", "interleaved", "cabi", "cabi_overlapped"):
            assert np.array_equal(got[k], want), (world, k)


def test_cpp_caller_gathers_over_every_visible_gpu(tmp_path):
    """tests/comm_gather.cpp: no Python, no torch in the data path -- contexts, scenes, renders and the NCCL all-gather
    all go through include/orz.h from one C++ process (the reference's kind of caller, Main.cpp:88,127,181-206)."""
    import torch

    from rasterizer_b200 import api, workloads as wl

    exe = os.path.join(ROOT, "tests", "_build", "comm_gather")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include", "-o", exe,
                           os.path.join(ROOT, "tests", "comm_gather.cpp"), "-L", os.path.join(ROOT, "rasterizer_b200"), "-lrasterizer_b200",
                           "-L", "/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath," + os.path.join(ROOT, "rasterizer_b200")])
    ps = wl.synthetic_city()
    ctx = api.Context(0)
    sc = api.Scene.from_prepared(ctx, ps)
    w, h, n = 640, 360, 45
    mvps, poss = wl.camera_path(ps, n, w, h)
    want = sc.render_views(w, h, mvps, cam_pos=poss, want=("vis",))["vis"]
    baked = tmp_path / "city.orzbake"
    sc.save(str(baked))
    sc.close(); ctx.close()
    mvps.tofile(tmp_path / "mvps.bin"); poss.tofile(tmp_path / "pos.bin")
    env = dict(os.environ)
    nccl = os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "nccl", "lib", "libnccl.so.2")
    if not os.path.exists("/usr/lib/x86_64-linux-gnu/libnccl.so.2") and os.path.exists(nccl):
        env["ORZ_NCCL_LIB"] = os.path.abspath(nccl)
    counts = sorted({1, torch.cuda.device_count()})
    for gpus in counts:
        out = tmp_path / f"out{gpus}.bin"
        subprocess.check_call([exe, str(baked), str(w), str(h), str(tmp_path / "mvps.bin"), str(tmp_path / "pos.bin"), str(out), str(gpus)],
                              env=env, timeout=600)
        got = np.fromfile(out, np.uint32).reshape(n, -1)
        assert np.array_equal(got, want), gpus
