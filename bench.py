#!/usr/bin/env python
"""Benchmark of the occlusion-culling hot path (BASELINE.json: views/sec at 1920x1080, Castle).

A step = one pass of the hot path over one batch of camera views: for every view the frame loop of
Main.cpp:181-206 (clear, setMVP, front-to-back gate query + rasterize<clip> per occluder) plus
queryVisibility for every occludee box; outputs per view: depth, HiZ, visibility bits.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our CUDA path
  python bench.py --impl reference [...]                          the reference's own CPU code
                                                                  (oracle/_ref, all host threads)
Under torchrun every rank renders its own share of the camera path (weak scaling of BASELINE config 3, the
headline line) and the per-view visibility bitmasks are gathered with NCCL through the product's C ABI
(orz_gather_bits); rank 0 prints ONE JSON line.  The same line carries, under "config5", BASELINE config 5 -- 8 192
visibility probes at 512x256 PARTITIONED over the N GPUs (strong scaling) -- and, at N = 1, Sponza / single-view /
config-4 extras with the reference's CPU time beside each.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from rasterizer_b200 import workloads as wl  # noqa: E402

WORKLOADS = {
    # name: (scene, width, height, views per GPU, camera set)
    "castle_1080p_path": ("castle", 1920, 1080, 1024, "path"),      # BASELINE configs[2]
    "castle_512x256_probes": ("castle", 512, 256, 8192, "probes"),  # BASELINE configs[4] (8192 probes partitioned over the GPUs)
    "sponza_1080p_path": ("sponza", 1920, 1080, 256, "path"),
    "city_640x360_path": ("city", 640, 360, 256, "path"),
}
PARTITIONED = {"castle_512x256_probes"}  # total views fixed, split over the ranks (strong scaling); the others: per-GPU views fixed


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="castle_1080p_path", choices=sorted(WORKLOADS))
    ap.add_argument("--views", type=int, default=0, help="views per GPU (default: the workload's)")
    ap.add_argument("--group-warps", type=int, default=0)
    ap.add_argument("--cluster-views", type=int, default=-1, help="largest batch that takes the cluster-per-view path (0 = batch kernel only; default: library's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="only the headline workload (no config5 / Sponza / single-view / config-4 objects)")
    ap.add_argument("--no-config4", action="store_true", help="skip the 5 M-quad soup (its scene takes ~20 s to generate and bake)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--parity-views", type=int, default=32, help="views of the timed batch compared with oracle/_ref outside the timed region")
    return ap.parse_args()


def pick_scene(name):
    if name != "city" and not wl.have_scene(name):
        return wl.load_scene("city"), f"synthetic city scene ({name} data not present on this box)"
    return wl.load_scene(name), "synthetic camera path over the reference's %s scene, random-free" % name.capitalize() if name != "city" else "synthetic city scene"


def make_views(ps, kind, n_total, w, h):
    return wl.camera_path(ps, n_total, w, h) if kind == "path" else wl.probe_views(ps, n_total, w, h)


def views_per_gpu(workload, world, override=0):
    if override:
        return override
    n = WORKLOADS[workload][3]
    return n // world if workload in PARTITIONED else n


def make_config(workload, ps, w, h, n_views, world, with_targets, group_warps=0):
    """One description of the workload for BOTH arms (the driver compares the two dicts)."""
    blocks = (w // 8) * (h // 8)
    return {"workload": workload, "scene": ps.name, "width": w, "height": h, "views_per_gpu": n_views, "occluders": len(ps.batches),
            "quads": ps.n_quads, "occludees": ps.n_quads,
            "outputs": "depth+HiZ+visibility bits per view (HBM resident)" if with_targets else "visibility bits per view",
            "l2": f"working set per step {n_views * (2 * w * h + 2 * blocks) / 1e6:.0f} MB of per-view depth+HiZ, larger than the 126 MB L2" if with_targets
                  else "L2 flushed by construction: every view clears and rewrites its scratch target",
            "group_warps": group_warps or "auto",
            "parallelism": f"views dealt round-robin to {world} GPU(s), NCCL all-gather of bitmasks (C ABI: orz_gather_bits)"}


def algorithmic_bytes(n_views, quads_submitted, n_occ, n_boxes, w, h, with_targets=True):
    """bytes that must move per view = 16 B per quad handed to rasterize + 32 B per gate
    box + depth and HiZ written once + 32 B per occludee box read + 1 bit per occludee + the matrix."""
    blocks = (w // 8) * (h // 8)
    per_view = 32 * n_occ + 32 * n_boxes + (n_boxes + 7) // 8 + 64 + ((2 * w * h + 2 * blocks) if with_targets else 0)
    return int(16 * int(quads_submitted) + n_views * per_view)


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.t0 = time.perf_counter()
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={device}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 8:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------
# the reference's own CPU implementation (oracle/_ref: the unmodified sources) on this box's host cores
class CpuReference:
    """Scene baked by the reference + one reference Rasterizer per thread, both built ONCE (the edge-mask table of a
    Rasterizer alone costs 0.5 s, Rasterizer.cpp:547-604), then timed as often as wanted."""

    def __init__(self, ps, w, h, n_threads):
        from oracle import ref_oracle as ro

        self.ro, self.ps, self.w, self.h, self.n_threads = ro, ps, w, h, n_threads
        self.kind = "reference" if ro.available() else "port"
        if self.kind == "reference":
            self.scene = ro.RefScene.from_batches(ps.batches, ps.ref_min, ps.ref_max)
            self.pool = ro.RefPool(w, h, n_threads)
            self.boxes = ps.quad_boxes()

    def close(self):
        if self.kind == "reference":
            self.pool.close(); self.scene.close()

    def time(self, mvps, poss, seconds=1.0, reps=None):
        """-> dict(value views/s, ...) over `mvps` repeated until about `seconds` of wall clock are filled."""
        if self.kind != "reference":
            return self._time_port(mvps, poss)
        orders = wl.orders_for(self.scene.centers, poss)
        if reps is None:
            wall, _ = self.pool.bench(self.scene, mvps, orders, self.boxes, 1)  # calibration pass (also warms caches and threads)
            reps = max(1, int(round(seconds / max(wall, 1e-4))))
        wall, out = self.pool.bench(self.scene, mvps, orders, self.boxes, reps)
        n = mvps.shape[0] * reps
        return dict(value=n / wall, unit="views/s", cores=self.n_threads, kind="reference", reps=reps, wall_s=wall,
                    sample=f"{mvps.shape[0]} views of the same camera set x {reps} reps, frame loop + {self.boxes.shape[0]} occludee queries per view, "
                           f"oracle/_ref (unmodified reference, g++ -O2 -mavx2 -mfma), one Rasterizer per thread",
                    frame_ms_per_view=1e3 * out[0] / n, query_ms_per_view=1e3 * out[1] / n,
                    mquads_per_s=out[2] / wall / 1e6, queries_per_s=self.boxes.shape[0] * n / wall)

    def _time_port(self, mvps, poss):
        # reference binary absent: the scalar port (much slower, single thread)
        from oracle import port_oracle as po

        ps = self.ps
        po.set_tables()
        baked = [po.bake(b, ps.ref_min, ps.ref_max) for b in ps.batches]
        packed = [b[0] for b in baked]
        centers, bmin, bmax = (np.stack([b[i] for b in baked]) for i in (1, 2, 3))
        boxes = ps.quad_boxes()
        port = po.PortRasterizer(self.w, self.h)
        t0 = time.perf_counter()
        n = 0
        for v in range(min(4, mvps.shape[0])):
            order = wl.orders_for(centers, poss[v:v + 1])[0]
            port.frame(packed, bmin, bmax, ps.ref_min, ps.ref_max, mvps[v], order)
            port.query_boxes(boxes)
            n += 1
        wall = time.perf_counter() - t0
        return dict(value=n / wall, unit="views/s", cores=1, kind="port", sample=f"{n} views, scalar C port (oracle/oracle_port.c)")


def run_reference(args):
    """Reference arm: the unmodified reference on ALL host threads, same workload description as our arm; every step a
    bounded sample (256 views of the same camera set, repeated to about one second)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    scene_name, w, h, _, kind = WORKLOADS[args.workload]
    ps, data = pick_scene(scene_name)
    n_views = views_per_gpu(args.workload, args.gpus, args.views)
    mvps, poss = make_views(ps, kind, n_views * args.gpus, w, h)
    sample = min(256, mvps.shape[0])
    idx = np.linspace(0, mvps.shape[0] - 1, sample).astype(int)
    threads = os.cpu_count() or 1
    cpu = CpuReference(ps, w, h, threads)
    with_targets = args.workload not in PARTITIONED
    res = cpu.time(mvps[idx], poss[idx], 1.0)  # calibrates the repetitions; untimed
    reps = res.get("reps")
    values = []
    for i in range(args.warmup + args.steps):
        res = cpu.time(mvps[idx], poss[idx], reps=reps)
        if i >= args.warmup:
            values.append(res["value"])
    cpu.close()
    value = float(np.mean(values))
    line = {
        "impl": "reference", "metric": "views_per_sec", "value": value, "unit": "views/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sample * (reps or 1) / value, "higher_is_better": True,
        "scaling": "strong" if args.workload in PARTITIONED else "weak", "vs_baseline": None,
        "dtype": "f32+u16", "data": data,
        "config": make_config(args.workload, ps, w, h, n_views, args.gpus, with_targets, args.group_warps),
        "cpu_baseline": {"value": value, "unit": "views/s", "cores": threads, "kind": res["kind"], "sample": res["sample"]},
        "e2e": {"value": value, "unit": "views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "mquads_per_sec": res.get("mquads_per_s"), "queries_per_sec": res.get("queries_per_s"),
        "step_values": values, "spread": float((max(values) - min(values)) / value) if values else None,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
class Rig:
    """Per-process state of our arm: torch plumbing (device tensors, events, the rendezvous), one context, the
    product's NCCL communicator."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist

        from rasterizer_b200 import api
        from rasterizer_b200 import distributed as D

        self.torch, self.dist, self.api, self.args = torch, dist, api, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (the product has no CPU path)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.ctx = api.Context(self.local)
        if args.group_warps:
            self.ctx.set_group_warps(args.group_warps)
        if args.cluster_views >= 0:
            self.ctx.set_cluster_views(args.cluster_views)
        self.stream = torch.cuda.ExternalStream(self.ctx.stream, device=self.dev)
        self.comm = D.make_comm(self.ctx) if self.world > 1 else None

    def sync_all(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max_over_ranks(self, x, dtype=None):
        t = self.torch.tensor([x], device=self.dev, dtype=dtype or self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def close(self):
        if self.comm:
            self.comm.close()
        self.ctx.close()
        if self.world > 1:
            self.dist.destroy_process_group()


def measure(rig, workload, n_views, steps, warmup, want_cpu_check=0, sample_clocks=True):
    """Time one workload on this rank's share of the views.  Returns a dict (rank 0's is the one printed):
    device-resident value, end-to-end value through host buffers, launches, clocks, parity sample."""
    torch, api, args = rig.torch, rig.api, rig.args
    rank, world, dev, ctx, stream = rig.rank, rig.world, rig.dev, rig.ctx, rig.stream
    scene_name, w, h, _, kind = WORKLOADS[workload]
    ps, data = pick_scene(scene_name)
    blocks = (w // 8) * (h // 8)
    scene = api.Scene.from_prepared(ctx, ps)
    n_boxes, n_occ, words = scene.n_boxes, scene.n_occluders, (scene.n_boxes + 31) // 32
    # views dealt round-robin (rank, rank + world, ...): camera paths are coherent, contiguous slices of the Castle
    # orbit differ in work by up to 1.7x (tools/slice_balance.py), which would measure the heaviest arc, not scaling
    mvps_all, poss_all = make_views(ps, kind, n_views * world, w, h)
    mvps, poss = np.ascontiguousarray(mvps_all[rank::world]), np.ascontiguousarray(poss_all[rank::world])
    with_targets = workload not in PARTITIONED  # probes: visibility bits only (depth stays scratch)

    d_mvps = torch.from_numpy(mvps).to(dev)
    d_pos = torch.from_numpy(poss).to(dev)
    d_vis = [torch.zeros((n_views, words), dtype=torch.int32, device=dev) for _ in range(2)]  # double buffered: the gather of step i
    d_all = [torch.zeros((world, n_views, words), dtype=torch.int32, device=dev) for _ in range(2)] if world > 1 else None  # overlaps step i + 1
    d_quads = torch.zeros(n_views, dtype=torch.int32, device=dev)
    d_depth = torch.empty((n_views, blocks * 64), dtype=torch.int16, device=dev) if with_targets else None
    d_hiz = torch.empty((n_views, blocks), dtype=torch.int16, device=dev) if with_targets else None

    def device_batch(k):
        b = api.ViewBatch()
        b.width, b.height, b.nViews, b.flags = w, h, n_views, 0
        b.mvps, b.camPos = d_mvps.data_ptr(), d_pos.data_ptr()  # order computed on the GPU (Main.cpp:185-190)
        b.visBits, b.quadsSubmitted = d_vis[k].data_ptr(), d_quads.data_ptr()
        if with_targets:
            b.depth, b.hiz = d_depth.data_ptr(), d_hiz.data_ptr()
        return b

    dbatch = [device_batch(0), device_batch(1)]

    def step(i, overlapped=True):
        k = i & 1
        if world > 1:
            rig.comm.join_older()  # the gather that last read this buffer pair (step i - 2) has finished
        scene.render_views_raw(dbatch[k], device=True)
        if world > 1:
            rig.comm.gather_bits(d_vis[k].data_ptr(), n_views * words, d_all[k].data_ptr(), overlapped=overlapped)

    def join():
        if world > 1:
            rig.comm.join()

    # ---- device-resident timing (value): inputs already in HBM
    for i in range(warmup):
        step(i)
    join()
    rig.sync_all()
    quads_submitted = int(d_quads.sum().item())
    sampler = ClockSampler(rig.local) if sample_clocks else None
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    ev_end = torch.cuda.Event(enable_timing=True)
    rig.sync_all()
    launches1 = ctx.launch_count
    for i, (a, m) in enumerate(ev):
        a.record(stream)
        step(i)
        m.record(stream)
    join()  # the last gather is inside the timed region
    ev_end.record(stream)
    rig.sync_all()
    launches = ctx.launch_count - launches1
    kern_ms = [a.elapsed_time(m) for a, m in ev]
    starts = [ev[0][0].elapsed_time(a) for a, _ in ev] + [ev[0][0].elapsed_time(ev_end)]
    step_ms = [starts[i + 1] - starts[i] for i in range(steps)]
    total_ms = rig.max_over_ranks(starts[-1])
    vis_ref = d_vis[(steps - 1) & 1].cpu().numpy().view(np.uint32).copy()

    # ---- multi-GPU correctness of the gathered buffer (outside the timed region): rank 0 re-renders a sample of
    # the OTHER ranks' views on its own GPU and compares with what the all-gather delivered
    gather_checked = 0
    if world > 1:
        got_all = d_all[(steps - 1) & 1].cpu().numpy().view(np.uint32)  # [rank][row][words]
        assert np.array_equal(got_all[rank], vis_ref), "all-gather: own rows differ from the local result"
        if rank == 0:
            rows = np.linspace(0, n_views - 1, min(8, n_views)).astype(int)
            for r in range(1, world):
                mv = np.ascontiguousarray(mvps_all[r::world][rows]); pp = np.ascontiguousarray(poss_all[r::world][rows])
                want = scene.render_views(w, h, mv, cam_pos=pp, want=("vis",))["vis"]
                assert np.array_equal(got_all[r][rows], want), f"all-gather: rank {r}'s rows differ from a single-GPU render"
                gather_checked += len(rows)

    # ---- end-to-end timing: host (pinned) inputs -> C ABI -> host visibility bits, every step
    h_mvps = torch.from_numpy(mvps).pin_memory()
    h_pos = torch.from_numpy(poss).pin_memory()
    h_vis = torch.zeros((n_views, words), dtype=torch.int32).pin_memory()
    h_all = torch.zeros((world, n_views, words), dtype=torch.int32).pin_memory() if world > 1 else None
    hb = api.ViewBatch()
    hb.width, hb.height, hb.nViews, hb.flags = w, h, n_views, api.BATCH_TARGETS_ON_DEVICE
    hb.mvps, hb.camPos, hb.visBits = h_mvps.data_ptr(), h_pos.data_ptr(), h_vis.data_ptr()
    if with_targets:
        hb.depth, hb.hiz = d_depth.data_ptr(), d_hiz.data_ptr()

    h_all2 = [h_all, torch.zeros_like(h_all).pin_memory()] if world > 1 else None  # two result buffers: step i's download runs beside step i + 1
    copy_stream = torch.cuda.Stream(device=dev) if world > 1 else None
    gathered = [torch.cuda.Event() for _ in range(2)] if world > 1 else None    # ctx stream: gather of this buffer pair has been joined
    downloaded = [torch.cuda.Event() for _ in range(2)] if world > 1 else None  # copy stream: its rows are in host memory

    def e2e_download(k):
        """the gathered bitmasks of buffer pair k -> host, on the copy stream, as soon as the gather has finished"""
        rig.comm.join()
        gathered[k].record(stream)
        copy_stream.wait_event(gathered[k])
        with torch.cuda.stream(copy_stream):
            h_all2[k].copy_(d_all[k], non_blocking=True)
        downloaded[k].record(copy_stream)

    def e2e_step(i):
        if world == 1:
            scene.render_views_raw(hb, device=False)   # orz_render_views: H2D matrices+positions, kernels, D2H bits, sync
            return
        # N GPUs: the same copies around the device entry + the collective, two steps in flight (what a consumer of a
        # stream of batches does): H2D, kernels, all-gather on the communicator's stream, then the download of ALL ranks'
        # rows on a copy stream while the next step's kernels run
        k = i & 1
        with torch.cuda.stream(stream):
            d_mvps.copy_(h_mvps, non_blocking=True); d_pos.copy_(h_pos, non_blocking=True)
        scene.render_views_raw(dbatch[k], device=True)
        if i >= 1:
            e2e_download(k ^ 1)                        # step i - 1's gather ran beside this step's kernels
        if i >= 2:
            stream.wait_event(downloaded[k])           # the rows step i - 2 left in this buffer pair are on the host (long since)
        rig.comm.gather_bits(d_vis[k].data_ptr(), n_views * words, d_all[k].data_ptr(), overlapped=True)

    def e2e_finish(n):
        if world > 1 and n:
            e2e_download((n - 1) & 1)
            copy_stream.synchronize()
        stream.synchronize()

    for i in range(warmup):
        e2e_step(i)
    e2e_finish(warmup)
    rig.sync_all()
    t0 = time.perf_counter()
    for i in range(steps):
        e2e_step(i)
    e2e_finish(steps)                                  # every step's result is in host memory inside the timed region
    rig.sync_all()
    e2e_s = rig.max_over_ranks(time.perf_counter() - t0)
    if world == 1:
        assert np.array_equal(h_vis.numpy().view(np.uint32), vis_ref), "e2e and device-resident paths disagree"
    else:
        for hb_k in h_all2[:min(2, steps)]:
            assert np.array_equal(hb_k.numpy().view(np.uint32)[rank], vis_ref), "e2e and device-resident paths disagree"
    clocks = None
    if sampler:
        # nvidia-smi samples every 100 ms and the timed regions are tens of ms long: keep the same load running (untimed)
        # until the sampler has seen at least half a second of it
        while time.perf_counter() - sampler.t0 < 0.6:
            scene.render_views_raw(dbatch[0], device=True)
            torch.cuda.synchronize()
        clocks = sampler.stop()

    # ---- parity sample against the unmodified reference on the host (outside the timed region)
    parity = None
    if want_cpu_check and rank == 0:
        from oracle import ref_oracle as ro

        if ro.available():
            rows = np.linspace(0, n_views - 1, min(want_cpu_check, n_views)).astype(int)
            ref = ro.RefScene.from_batches(ps.batches, ps.ref_min, ps.ref_max)
            orders = wl.orders_for(ref.centers, poss[rows])
            kw = {}
            if with_targets:
                kw = dict(depth=d_depth[torch.from_numpy(rows).to(dev)].cpu().numpy(), hiz=d_hiz[torch.from_numpy(rows).to(dev)].cpu().numpy())
            mism = ro.check_views(ref, w, h, mvps[rows], orders, ps.quad_boxes(), vis=vis_ref[rows], **kw)
            ref.close()
            assert not mism.any(), "bench: the timed batch differs from oracle/_ref: " + ro.describe_mismatch(mism)
            parity = {"views": int(len(rows)), "result": "bit-exact", "compared": ("depth, HiZ, " if with_targets else "") + "visibility bits",
                      "against": "oracle/_ref (unmodified reference) on this box's host"}
        else:
            parity = {"views": 0, "result": "oracle/_ref not present on this box"}

    total_views = n_views * world
    kern = float(np.mean(kern_ms))
    res = dict(workload=workload, ps=ps, data=data, w=w, h=h, n_views=n_views, with_targets=with_targets, n_occ=n_occ, n_boxes=n_boxes,
               value=total_views * steps / (total_ms / 1e3), ms_per_step=total_ms / steps, kernel_ms=kern, step_ms=step_ms,
               e2e_value=total_views * steps / e2e_s, h2d=int(mvps.nbytes + poss.nbytes), d2h=int((h_all if world > 1 else h_vis).numel() * 4),
               launches=int(launches), clocks=clocks, quads_submitted=quads_submitted, parity=parity, gather_checked=gather_checked,
               mvps=mvps, poss=poss, scene=scene)
    return res


def single_view(rig, name, w, h, cpu_ms=True):
    """BASELINE configs[0]/[1]: ONE view (the scene's default camera) through the host entry point: matrix in, bits out."""
    from rasterizer_b200 import camera as cam

    if not wl.have_scene(name):
        return None
    ps = wl.load_scene(name)
    scene = rig.api.Scene.from_prepared(rig.ctx, ps)
    c = ps.camera
    one_mvp = cam.view_projection(c["pos"], c["dir"], c["up"], c["fov"], w, h)[None]
    one_pos = np.array(c["pos"], np.float32)[None]
    for _ in range(5):
        scene.render_views(w, h, one_mvp, cam_pos=one_pos, want=("vis",))
    t = []
    for _ in range(40):
        t0 = time.perf_counter()
        scene.render_views(w, h, one_mvp, cam_pos=one_pos, want=("vis",))
        t.append(time.perf_counter() - t0)
    out = {"scene": name, "ms": float(np.median(t)) * 1e3, "ms_min": float(np.min(t)) * 1e3,
           "what": "one view, default camera: host matrix in -> frame loop + all occludee queries -> host bits out (wall clock, includes H2D/D2H and launches; median of 40)"}
    scene.close()
    if cpu_ms:
        cpu = CpuReference(ps, w, h, 1)
        cb = cpu.time(one_mvp, one_pos, 1.0)
        cpu.close()
        if "frame_ms_per_view" in cb:
            out["reference_ms_one_core"] = cb["frame_ms_per_view"] + cb["query_ms_per_view"]
    return out


def config4(rig, n_quads=5_000_000, w=3840, h=2160):
    """BASELINE configs[3]: 5 M-quad soup (10 M triangles), 3840x2160, camera inside the geometry, every batch through
    rasterize<true>, no gate.  One view per step; the reference on one host core beside it."""
    from rasterizer_b200 import camera as cam

    torch, api = rig.torch, rig.api
    ps = wl.synthetic_soup(n_quads)
    boxes = ps.quad_boxes()[::97]
    scene = api.Scene.bake_on_device(rig.ctx, ps.batches, ps.ref_min, ps.ref_max, boxes)
    c = ps.camera
    dirs = ((0.0, 0.0, 1.0), (0.6, -0.2, 0.7), (-0.5, 0.3, -0.8), (0.1, 0.9, 0.2))
    mvps = np.stack([cam.view_projection(c["pos"], d, c["up"], c["fov"], w, h) for d in dirs]).astype(np.float32)
    poss = np.zeros((len(dirs), 3), np.float32)
    flags = api.BATCH_NO_GATE | api.BATCH_FORCE_CLIPPED
    blocks = (w // 8) * (h // 8)
    dev = rig.dev
    d_mvps, d_pos = torch.from_numpy(mvps).to(dev), torch.from_numpy(poss).to(dev)
    words = (scene.n_boxes + 31) // 32
    d_vis = torch.zeros((1, words), dtype=torch.int32, device=dev)
    d_depth = torch.empty((1, blocks * 64), dtype=torch.int16, device=dev)
    d_hiz = torch.empty((1, blocks), dtype=torch.int16, device=dev)

    def batch(v):
        b = api.ViewBatch()
        b.width, b.height, b.nViews, b.flags = w, h, 1, flags
        b.mvps, b.camPos = d_mvps[v:v + 1].data_ptr(), d_pos[v:v + 1].data_ptr()
        b.visBits, b.depth, b.hiz = d_vis.data_ptr(), d_depth.data_ptr(), d_hiz.data_ptr()
        return b

    bs = [batch(v) for v in range(len(dirs))]
    for b in bs:
        scene.render_views_raw(b, device=True)
    torch.cuda.synchronize()
    sampler = ClockSampler(rig.local)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(3 * len(bs))]
    n0 = rig.ctx.launch_count
    for i, (a, z) in enumerate(ev):
        a.record(rig.stream)
        scene.render_views_raw(bs[i % len(bs)], device=True)
        z.record(rig.stream)
    torch.cuda.synchronize()
    launches = rig.ctx.launch_count - n0
    ms = [a.elapsed_time(z) for a, z in ev]
    while time.perf_counter() - sampler.t0 < 0.6:
        scene.render_views_raw(bs[0], device=True)
        torch.cuda.synchronize()
    clocks = sampler.stop()
    per_view = float(np.mean(ms))
    alg = 16 * ps.n_quads + 2 * w * h + 2 * blocks + 32 * scene.n_boxes + 64
    peak, _ = measured_peak()
    out = {"workload": f"soup {ps.n_quads} quads ({2 * ps.n_quads} triangles), {w}x{h}, no gate, rasterize<true>", "ms_per_view": per_view,
           "views_per_s": 1e3 / per_view, "mquads_per_s": ps.n_quads / per_view / 1e3, "ms_each": ms, "launches_per_view": launches / len(ev), "clocks": clocks,
           "roofline": {"bound": "hbm", "algorithmic_bytes_per_view": alg, "achieved": alg / (per_view / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": alg / (per_view / 1e3) / 1e9 / peak}}
    if not rig.args.no_cpu_baseline:
        from oracle import ref_oracle as ro

        if ro.available():
            ref = ro.RefScene.from_batches(ps.batches, ps.ref_min, ps.ref_max)
            orders = wl.orders_for(ref.centers, poss)
            r = ro.RefRasterizer(w, h)
            t = []
            for v in range(len(dirs)):
                t0 = time.perf_counter()
                r.submit_all(ref, mvps[v], orders[v], True, zero_depth=False)
                t.append(time.perf_counter() - t0)
            out["reference_ms_per_view_one_core"] = float(np.mean(t)) * 1e3
            out["speedup_vs_one_core"] = out["reference_ms_per_view_one_core"] / per_view
            # parity of the timed views (outside the timed region)
            got = scene.render_views(w, h, mvps[:2], orders=orders[:2], flags=flags, want=("depth", "hiz", "vis"))
            mism = ro.check_views(ref, w, h, mvps[:2], orders[:2], boxes, mode=3, depth=got["depth"], hiz=got["hiz"], vis=got["vis"])
            assert not mism.any(), "bench config 4 differs from oracle/_ref: " + ro.describe_mismatch(mism)
            out["parity"] = {"views": 2, "result": "bit-exact", "compared": "depth, HiZ, visibility bits"}
            r.close(); ref.close()
    scene.close()
    return out


def run_ours(args):
    rig = Rig(args)
    rank, world = rig.rank, rig.world
    n_views = views_per_gpu(args.workload, world, args.views)
    res = measure(rig, args.workload, n_views, args.steps, args.warmup, want_cpu_check=args.parity_views)
    ps, w, h = res["ps"], res["w"], res["h"]
    line = None
    if rank == 0:
        peak, peak_src = measured_peak()
        alg = algorithmic_bytes(n_views, res["quads_submitted"], res["n_occ"], res["n_boxes"], w, h, res["with_targets"])
        kern = res["kernel_ms"]
        achieved = alg / (kern / 1e3) / 1e9
        traffic = ncu_traffic()
        cluster_limit = args.cluster_views if args.cluster_views >= 0 else 16384
        cluster_path = n_views <= cluster_limit and (w // 8) * (h // 8) <= 65536
        line = {
            "metric": "views_per_sec", "value": res["value"], "unit": "views/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "strong" if args.workload in PARTITIONED else "weak",
            "vs_baseline": None, "dtype": "f32+u16", "data": res["data"],
            "config": make_config(args.workload, ps, w, h, n_views, world, res["with_targets"], args.group_warps),
            "mquads_per_sec": res["quads_submitted"] * world / (kern / 1e3) / 1e6,
            "queries_per_sec": res["n_boxes"] * n_views * world / (kern / 1e3),
            "kernel_ms_per_step": kern, "step_ms": res["step_ms"],
            "e2e": {"value": res["e2e_value"], "unit": "views/s", "h2d_bytes_per_step": res["h2d"], "d2h_bytes_per_step": res["d2h"],
                    "note": "depth/HiZ are written to HBM per view and stay device-resident; the caller-visible result is the bitmask"
                            + ("; N GPUs: every rank downloads ALL ranks' gathered rows each step, the download of step i (copy stream) and its "
                               "all-gather (communicator stream) run beside the kernels of step i + 1, everything on the host before the clock stops"
                               if world > 1 else "")},
            "gpu_launches": res["launches"],
            "clocks": res["clocks"],
            "parity_checked_views": (res["parity"] or {}).get("views", 0), "parity": res["parity"],
            "gather_checked_rows": res["gather_checked"],
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (traffic or {}).get("dram_bytes_per_launch"), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg,
                         "kernel": ("one step = k_prepare_views + k_sort_views + k_setup_views (speculative setup of every occluder in the frustum) + "
                                    "k_raster_views_cluster (one thread-block cluster per view, dataflow gates, tile-major depth) + k_query_views; "
                                    if cluster_path else
                                    "one step = k_prepare_views + k_sort_views + 4 x (k_render_views<GW> + k_query_views), the four cost-sorted sub-batches overlapped on four streams; ")
                                   + "duration = CUDA events around the step on the context stream",
                         "traffic_note": (traffic or {}).get("note"),
                         "note": "issue/latency bound by design (SURVEY 8d): HBM is not the limiter; ncu per-launch counters in profiles/"},
        }
        if not args.no_cpu_baseline and world == 1:
            sample = min(128, n_views)
            idx = np.linspace(0, n_views - 1, sample).astype(int)
            cpu = CpuReference(ps, w, h, 1)
            cb = cpu.time(res["mvps"][idx], res["poss"][idx], args.cpu_seconds)
            cpu.close()
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
            line["cpu_baseline"].update({k: cb[k] for k in ("frame_ms_per_view", "query_ms_per_view", "mquads_per_s", "queries_per_s") if k in cb})
    res["scene"].close()

    if not args.no_extras:
        # ---- BASELINE config 5: 8 192 probes at 512x256 partitioned over the ranks (strong scaling), bits only
        if args.workload != "castle_512x256_probes" and wl.have_scene("castle"):
            n5 = views_per_gpu("castle_512x256_probes", world)
            r5 = measure(rig, "castle_512x256_probes", n5, max(args.steps, 10), args.warmup, want_cpu_check=16)
            if rank == 0:
                line["config5"] = {"workload": "castle_512x256_probes", "total_views": n5 * world, "views_per_gpu": n5, "scaling": "strong",
                                   "views_per_s": r5["value"], "ms_per_step": r5["ms_per_step"], "kernel_ms": r5["kernel_ms"], "step_ms": r5["step_ms"],
                                   "e2e_views_per_s": r5["e2e_value"], "gpu_launches": r5["launches"], "clocks": r5["clocks"], "parity": r5["parity"],
                                   "gather_checked_rows": r5["gather_checked"],
                                   "note": "visibility bits only; gather of step i (orz_gather_bits_overlapped) runs beside step i + 1, the last one inside the timed region"}
            r5["scene"].close()
        if rank == 0 and world == 1:
            # ---- single views (BASELINE configs 0 and 1) and the Sponza path, reference on one core beside each
            line["single_view"] = single_view(rig, "castle", 1920, 1080, not args.no_cpu_baseline)
            sv2 = single_view(rig, "sponza", 1920, 1080, not args.no_cpu_baseline)
            if sv2:
                line["single_view_sponza"] = sv2
            if wl.have_scene("sponza") and args.workload != "sponza_1080p_path":
                rs = measure(rig, "sponza_1080p_path", 256, 5, 3, want_cpu_check=8, sample_clocks=False)
                line["sponza_1080p_256_views"] = {"views_per_s": rs["value"], "ms_per_step": rs["ms_per_step"], "e2e_views_per_s": rs["e2e_value"], "parity": rs["parity"]}
                if not args.no_cpu_baseline:
                    idx = np.linspace(0, 255, 32).astype(int)
                    cpu = CpuReference(rs["ps"], 1920, 1080, 1)
                    cb = cpu.time(rs["mvps"][idx], rs["poss"][idx], 4.0)
                    cpu.close()
                    line["sponza_1080p_256_views"]["reference_views_per_s_one_core"] = cb["value"]
                rs["scene"].close()
            if not args.no_config4:
                line["config4"] = config4(rig)
    if rank == 0:
        print(json.dumps(line), flush=True)
    rig.close()


if __name__ == "__main__":
    # the contract is ONE JSON line on stdout: libraries (NCCL prints its version banner there when NCCL_DEBUG is
    # set) write to fd 1 directly, so fd 1 is pointed at stderr and the line goes to the original stdout
    import faulthandler

    faulthandler.enable()  # a crash inside a native library still leaves a Python traceback on stderr
    _real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = _real_stdout
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
