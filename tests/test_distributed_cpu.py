"""N>1 host logic on CPU: world_size-2 gloo process group; slices cover every view exactly once and
the all-gather reassembles per-view bitmasks in view order (the GPU render itself is covered by
the -m gpu tests; here each rank fabricates its slice's bits deterministically)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from rasterizer_b200 import distributed as D  # noqa: E402


def _fake_bits(view_ids, words):
    v = np.asarray(view_ids, np.uint32)[:, None]
    return ((v * np.uint32(2654435761)) ^ (np.arange(words, dtype=np.uint32)[None, :] * np.uint32(40503))).astype(np.uint32)


def _worker(rank, world, port, n_views, words, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    start, stop, per = D.view_slice(n_views, rank, world)
    local = np.zeros((per, words), np.uint32)
    local[: stop - start] = _fake_bits(range(start, stop), words)
    out = D.all_gather_bits(torch.from_numpy(local.view(np.int32)), n_views)
    q.put((rank, out.numpy().view(np.uint32).copy()))
    dist.barrier()
    dist.destroy_process_group()


class _FakeScene:
    """Stands in for api.Scene on CPU: `render_views` fabricates the bits of the views it is handed
    (identified by the first matrix element), so the sharding + gather logic runs without a GPU."""

    n_boxes = 5 * 32

    def render_views(self, width, height, mvps, flags=0, want=("vis",), **kw):
        return {"vis": _fake_bits(np.asarray(mvps)[:, 0].astype(np.uint32), 5)}


def _worker_sharded(rank, world, port, n_views, balance, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mvps = np.zeros((n_views, 16), np.float32)
    mvps[:, 0] = np.arange(n_views)
    out = D.render_views_sharded(_FakeScene(), 64, 64, mvps, cam_pos=np.zeros((n_views, 3), np.float32), balance=balance)
    q.put((rank, out.numpy().view(np.uint32).copy()))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("n_views", [8, 13, 1])
def test_slices_and_gather_world2(n_views):
    world, words = 2, 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_views, words, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = _fake_bits(range(n_views), words)
    for r in range(world):
        assert np.array_equal(got[r], want)


def test_view_slices_partition():
    for n in (0, 1, 7, 8, 1024, 8191):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                a, b, per = D.view_slice(n, r, world)
                assert 0 <= b - a <= per
                seen += list(range(a, b))
            assert seen == list(range(n))


@pytest.mark.parametrize("balance", ["contiguous", "interleaved"])
@pytest.mark.parametrize("n_views", [9, 2, 1])
def test_render_views_sharded_world2(n_views, balance):
    """Both ways of dealing views to ranks return every view's bits in view order on every rank."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_sharded, args=(r, world, port, n_views, balance, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = _fake_bits(range(n_views), 5)
    for r in range(world):
        assert np.array_equal(got[r], want), (balance, r)
