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
