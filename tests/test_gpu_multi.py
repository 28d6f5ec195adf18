"""Multi-GPU view-batch path on real devices: two ranks (one per GPU, NCCL) render their slices of
one camera path and all-gather the visibility bitmasks; the result must equal the single-GPU run.
Skipped when fewer than two GPUs are visible."""
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
bits = D.render_views_sharded(sc, 640, 360, mvps, cam_pos=poss, device=torch.device("cuda", local))
if rank == 0:
    np.save(sys.argv[2], bits.cpu().numpy().view(np.uint32))
dist.barrier()
sc.close(); ctx.close()
dist.destroy_process_group()
'''


def test_two_gpu_gather_equals_single_gpu(tmp_path):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from rasterizer_b200 import api, workloads as wl

    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    out = tmp_path / "bits.npy"
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                           "--master-port", str(port), str(script), ROOT, str(out)], timeout=600)
    got = np.load(out)
    ps = wl.synthetic_city()
    ctx = api.Context(0)
    sc = api.Scene.from_prepared(ctx, ps)
    mvps, poss = wl.camera_path(ps, 37, 640, 360)
    want = sc.render_views(640, 360, mvps, cam_pos=poss, want=("vis",))["vis"]
    assert np.array_equal(got, want)
    sc.close(); ctx.close()
