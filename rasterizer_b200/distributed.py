"""Multi-GPU side of the view-batch path: views are independent, so the view list is
cut into contiguous equal slices (tail padded), every rank renders its slice on its own GPU with
the static scene replicated, and ONE collective -- an all-gather of the per-view visibility
bitmasks -- assembles the result.  Depth / HiZ stay on the GPU that produced them.  One process per
GPU; torch.distributed is only the plumbing (NCCL over NVLink on the GPUs, gloo in CPU tests)."""
from __future__ import annotations

import numpy as np


def view_slice(n_views: int, rank: int, world: int):
    """Contiguous slice of rank `rank` -> (start, stop, per_rank); every rank owns `per_rank` rows
    of the gathered buffer, the last ranks may own fewer real views (padding rows are zero)."""
    per = (n_views + world - 1) // world
    start = min(rank * per, n_views)
    stop = min(start + per, n_views)
    return start, stop, per


def all_gather_bits(local_bits, n_views: int):
    """local_bits: torch tensor [per_rank, words] (int32, rows beyond the rank's real views zero),
    on the device the process group's backend expects.  Returns [n_views, words] on every rank."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return local_bits[:n_views]
    out = torch.empty((world,) + tuple(local_bits.shape), dtype=local_bits.dtype, device=local_bits.device)
    dist.all_gather_into_tensor(out.view(-1), local_bits.contiguous().view(-1))
    return out.view(world * local_bits.shape[0], -1)[:n_views]


def interleaved_rows(n_views: int, rank: int, world: int):
    """Views rank, rank + world, ... -> (indices, per_rank).  Camera paths are coherent, so contiguous
    slices differ in work by up to 1.7x (Castle orbit, tools/slice_balance.py); interleaving gives
    every rank the same mix."""
    per = (n_views + world - 1) // world
    return np.arange(rank, n_views, world), per


def make_comm(ctx):
    """The product's own NCCL communicator (C ABI: orz_comm_create) for this rank's context.  torch.distributed is
    only the rendezvous: rank 0's 128-byte NCCL id travels through broadcast_object_list."""
    import torch.distributed as dist

    from . import api

    rank, world = dist.get_rank(), dist.get_world_size()
    box = [api.Comm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    return api.Comm(ctx, world, rank, box[0])


def render_views_gathered(scene, comm, width, height, mvps, cam_pos, flags=0, overlapped=False):
    """Device-resident flavour of render_views_sharded through the C ABI only: views dealt round-robin, rendered with
    orz_render_views_device into a torch buffer, gathered with orz_gather_bits (NCCL all-gather on the context stream,
    or beside it when `overlapped`).  Returns a torch int32 tensor [n_views, words] on the GPU, identical on all ranks."""
    import torch

    from . import api

    rank, world = comm.rank, comm.n_ranks
    dev = torch.device("cuda", scene.ctx.device)
    mvps = np.ascontiguousarray(mvps, np.float32).reshape(-1, 16)
    n = mvps.shape[0]
    idx, per = interleaved_rows(n, rank, world)
    words = (scene.n_boxes + 31) // 32
    d_mvp = torch.from_numpy(np.ascontiguousarray(mvps[idx])).to(dev)
    d_pos = torch.from_numpy(np.ascontiguousarray(np.asarray(cam_pos, np.float32).reshape(-1, 3)[idx])).to(dev)
    local = torch.zeros((per, words), dtype=torch.int32, device=dev)
    gathered = torch.empty((world, per, words), dtype=torch.int32, device=dev)
    torch.cuda.synchronize(dev)
    if idx.size:
        b = api.ViewBatch()
        b.width, b.height, b.nViews, b.flags = width, height, int(idx.size), flags
        b.mvps, b.camPos, b.visBits = d_mvp.data_ptr(), d_pos.data_ptr(), local.data_ptr()
        scene.render_views_raw(b, device=True)
    comm.gather_bits(local.data_ptr(), per * words, gathered.data_ptr(), overlapped=overlapped)
    comm.synchronize()
    return gathered.transpose(0, 1).reshape(per * world, words)[:n].contiguous()  # row r * per + i holds view i * world + r


def render_views_sharded(scene, width, height, mvps, cam_pos=None, orders=None, flags=0, device=None, balance="contiguous"):
    """Render rank's slice of `mvps` through the C ABI and gather every view's visibility bitmask.
    Returns a torch int32 tensor [n_views, words] identical on all ranks.  balance="interleaved"
    deals the views round-robin instead of in contiguous slices (same result, even work)."""
    import torch
    import torch.distributed as dist

    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    mvps = np.ascontiguousarray(mvps, np.float32).reshape(-1, 16)
    n = mvps.shape[0]
    if balance == "interleaved":
        idx, per = interleaved_rows(n, rank, world)
        words = (scene.n_boxes + 31) // 32
        local = np.zeros((per, words), np.uint32)
        if idx.size:
            kw = dict(orders=np.asarray(orders)[idx]) if orders is not None else dict(cam_pos=np.asarray(cam_pos)[idx])
            local[: idx.size] = scene.render_views(width, height, mvps[idx], flags=flags, want=("vis",), **kw)["vis"]
        t = torch.from_numpy(local.view(np.int32))
        if device is not None:
            t = t.to(device)
        g = all_gather_bits(t, per * world)  # row r * per + i holds view i * world + r
        return g.view(world, per, -1).transpose(0, 1).reshape(per * world, -1)[:n].contiguous()
    start, stop, per = view_slice(n, rank, world)
    words = (scene.n_boxes + 31) // 32
    local = np.zeros((per, words), np.uint32)
    if stop > start:
        kw = dict(orders=np.asarray(orders)[start:stop]) if orders is not None else dict(cam_pos=np.asarray(cam_pos)[start:stop])
        out = scene.render_views(width, height, mvps[start:stop], flags=flags, want=("vis",), **kw)
        local[: stop - start] = out["vis"]
    t = torch.from_numpy(local.view(np.int32))
    if device is not None:
        t = t.to(device)
    return all_gather_bits(t, n)
