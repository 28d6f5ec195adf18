#!/usr/bin/env python
"""Benchmark of the occlusion-culling hot path (BASELINE.json: views/sec at 1920x1080, Castle).

A step = one pass of the hot path over one batch of camera views: for every view the frame loop of
Main.cpp:181-206 (clear, setMVP, front-to-back gate query + rasterize<clip> per occluder) plus
queryVisibility for every occludee box; outputs per view: depth, HiZ, visibility bits.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our CUDA path
  python bench.py --impl reference [...]                          the reference's own CPU code
                                                                  (oracle/_ref, all host threads)
Under torchrun every rank renders its own slice of the camera path (weak scaling) and the
per-view visibility bitmasks are gathered with NCCL; rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import ctypes as C
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
    "castle_512x256_probes": ("castle", 512, 256, 8192, "probes"),  # BASELINE configs[4] (per-GPU slice of 8192/N)
    "sponza_1080p_path": ("sponza", 1920, 1080, 256, "path"),
    "city_640x360_path": ("city", 640, 360, 256, "path"),
}


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
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    return ap.parse_args()


def pick_scene(name):
    if name != "city" and not wl.have_scene(name):
        return wl.load_scene("city"), f"synthetic city scene ({name} data not present on this box)"
    return wl.load_scene(name), "synthetic camera path over the reference's %s scene, random-free" % name.capitalize() if name != "city" else "synthetic city scene"


def make_views(ps, kind, n_total, w, h):
    return wl.camera_path(ps, n_total, w, h) if kind == "path" else wl.probe_views(ps, n_total, w, h)


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
def cpu_reference(ps, w, h, mvps, poss, n_threads, seconds, want_kind="reference"):
    """Time the reference's own CPU implementation (oracle/_ref: unmodified sources) on host cores.
    Returns dict(value views/s, queries/s, mquads/s, kind, cores, sample)."""
    from oracle import ref_oracle as ro

    if ro.available():
        s = ro.RefScene.from_batches(ps.batches, ps.ref_min, ps.ref_max)
        boxes = ps.quad_boxes()
        orders = wl.orders_for(s.centers, poss)
        # calibrate with one pass, then repeat to fill the time budget
        wall, out = ro.bench_views(s, w, h, mvps, orders, boxes, n_threads, 1)
        reps = max(1, int(seconds / max(wall, 1e-3)))
        wall, out = ro.bench_views(s, w, h, mvps, orders, boxes, n_threads, reps)
        n = mvps.shape[0] * reps
        res = dict(value=n / wall, unit="views/s", cores=n_threads, kind="reference",
                   sample=f"{mvps.shape[0]} views of the same camera set x {reps} reps, frame loop + {boxes.shape[0]} occludee queries per view, "
                          f"oracle/_ref (unmodified reference, g++ -O2 -mavx2 -mfma), one Rasterizer per thread",
                   frame_ms_per_view=1e3 * out[0] / n, query_ms_per_view=1e3 * out[1] / n,
                   mquads_per_s=out[2] / wall / 1e6, queries_per_s=boxes.shape[0] * n / wall)
        s.close()
        return res
    # reference binary absent: fall back to the scalar port (much slower, single thread)
    from oracle import port_oracle as po

    po.set_tables()
    baked = [po.bake(b, ps.ref_min, ps.ref_max) for b in ps.batches]
    packed = [b[0] for b in baked]
    centers, bmin, bmax = (np.stack([b[i] for b in baked]) for i in (1, 2, 3))
    boxes = ps.quad_boxes()
    port = po.PortRasterizer(w, h)
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
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    scene_name, w, h, views, kind = WORKLOADS[args.workload]
    ps, data = pick_scene(scene_name)
    n_views = args.views or views
    mvps, poss = make_views(ps, kind, n_views * args.gpus, w, h)
    sample = min(256, n_views)
    idx = np.linspace(0, mvps.shape[0] - 1, sample).astype(int)
    threads = os.cpu_count() or 1
    per_step = []
    res = None
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        res = cpu_reference(ps, w, h, mvps[idx], poss[idx], threads, 0.0)
        if i >= args.warmup:
            per_step.append((time.perf_counter() - t0, res["value"]))
    value = float(np.mean([v for _, v in per_step]))
    line = {
        "impl": "reference", "metric": "views_per_sec", "value": value, "unit": "views/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sample / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32+u16", "data": data,
        "config": {"workload": args.workload, "scene": ps.name, "width": w, "height": h, "views_per_step": sample,
                   "occluders": len(ps.batches), "quads": ps.n_quads, "occludees": ps.n_quads},
        "cpu_baseline": {"value": value, "unit": "views/s", "cores": threads, "kind": res["kind"], "sample": res["sample"]},
        "e2e": {"value": value, "unit": "views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "mquads_per_sec": res.get("mquads_per_s"), "queries_per_sec": res.get("queries_per_s"),
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    from rasterizer_b200 import api

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    scene_name, w, h, views, kind = WORKLOADS[args.workload]
    ps, data = pick_scene(scene_name)
    n_views = args.views or views
    if args.workload == "castle_512x256_probes" and not args.views:
        n_views = views // world  # config 5: 8192 probes partitioned over the GPUs
    blocks = (w // 8) * (h // 8)

    ctx = api.Context(local)
    if args.group_warps:
        ctx.set_group_warps(args.group_warps)
    if args.cluster_views >= 0:
        ctx.set_cluster_views(args.cluster_views)
    cluster_limit = args.cluster_views if args.cluster_views >= 0 else 1024
    cluster_path = n_views <= cluster_limit and blocks <= 65536
    scene = api.Scene.from_prepared(ctx, ps)
    n_boxes, n_occ, words = scene.n_boxes, scene.n_occluders, (scene.n_boxes + 31) // 32
    # views dealt round-robin (rank, rank + world, ...): camera paths are coherent, contiguous slices of the Castle
    # orbit differ in work by up to 1.7x (tools/slice_balance.py), which would measure the heaviest arc, not scaling
    mvps_all, poss_all = make_views(ps, kind, n_views * world, w, h)
    mvps, poss = np.ascontiguousarray(mvps_all[rank::world]), np.ascontiguousarray(poss_all[rank::world])
    with_targets = args.workload != "castle_512x256_probes"  # probes: visibility bits only (depth stays scratch)

    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    d_mvps = torch.from_numpy(mvps).to(dev)
    d_pos = torch.from_numpy(poss).to(dev)
    d_vis = torch.zeros((n_views, words), dtype=torch.int32, device=dev)
    d_quads = torch.zeros(n_views, dtype=torch.int32, device=dev)
    d_depth = torch.empty((n_views, blocks * 64), dtype=torch.int16, device=dev) if with_targets else None
    d_hiz = torch.empty((n_views, blocks), dtype=torch.int16, device=dev) if with_targets else None
    d_all = torch.zeros((world, n_views, words), dtype=torch.int32, device=dev) if world > 1 else None

    def batch(host: bool, h_mvps=None, h_pos=None, h_vis=None):
        b = api.ViewBatch()
        b.width, b.height, b.nViews = w, h, n_views
        b.flags = api.BATCH_TARGETS_ON_DEVICE if host else 0
        b.mvps = h_mvps.data_ptr() if host else d_mvps.data_ptr()
        b.camPos = h_pos.data_ptr() if host else d_pos.data_ptr()  # order computed on the GPU (Main.cpp:185-190)
        b.visBits = h_vis.data_ptr() if host else d_vis.data_ptr()
        b.quadsSubmitted = None if host else d_quads.data_ptr()
        if with_targets:
            b.depth, b.hiz = d_depth.data_ptr(), d_hiz.data_ptr()
        return b

    def gather():
        if world > 1:
            with torch.cuda.stream(stream):
                dist.all_gather_into_tensor(d_all.view(-1), d_vis.view(-1))

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident timing (value): inputs already in HBM
    dbatch = batch(False)
    launches0 = ctx.launch_count
    for _ in range(args.warmup):
        scene.render_views_raw(dbatch, device=True)
        gather()
    sync_all()
    quads_submitted = int(d_quads.sum().item())
    sampler = ClockSampler(local)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches1 = ctx.launch_count
    sync_all()
    for a, m, z in ev:
        a.record(stream)
        scene.render_views_raw(dbatch, device=True)
        m.record(stream)
        gather()
        z.record(stream)
    sync_all()
    launches = ctx.launch_count - launches1
    step_ms = [a.elapsed_time(z) for a, m, z in ev]
    kern_ms = [a.elapsed_time(m) for a, m, z in ev]
    total_ms = ev[0][0].elapsed_time(ev[-1][2])
    t = torch.tensor([total_ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    vis_ref = d_vis.cpu().numpy().copy()

    # ---- end-to-end timing: host (pinned) inputs -> C ABI -> host visibility bits, every step
    h_mvps = torch.from_numpy(mvps).pin_memory()
    h_pos = torch.from_numpy(poss).pin_memory()
    h_vis = torch.zeros((n_views, words), dtype=torch.int32).pin_memory()
    hbatch = batch(True, h_mvps, h_pos, h_vis)
    for _ in range(args.warmup):
        scene.render_views_raw(hbatch, device=False)
        gather()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        scene.render_views_raw(hbatch, device=False)   # H2D matrices+positions, kernel, D2H bits, sync
        gather()
    sync_all()
    e2e_s = time.perf_counter() - t0
    # nvidia-smi samples every 100 ms and the timed regions are tens of ms long: keep the same load running (untimed)
    # until the sampler has seen at least half a second of it
    while time.perf_counter() - sampler.t0 < 0.6:
        scene.render_views_raw(dbatch, device=True)
        torch.cuda.synchronize()
    clocks = sampler.stop()
    t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    assert np.array_equal(h_vis.numpy(), vis_ref), "e2e and device-resident paths disagree"

    if rank == 0:
        total_views = n_views * world
        value = total_views * args.steps / (total_ms / 1e3)
        e2e_value = total_views * args.steps / e2e_s
        peak, peak_src = measured_peak()
        alg = algorithmic_bytes(n_views, quads_submitted, n_occ, n_boxes, w, h, with_targets)
        kern = float(np.mean(kern_ms))
        achieved = alg / (kern / 1e3) / 1e9
        traffic = ncu_traffic()
        line = {
            "metric": "views_per_sec", "value": value, "unit": "views/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak" if args.workload != "castle_512x256_probes" else "strong",
            "vs_baseline": None, "dtype": "f32+u16", "data": data,
            "config": {"workload": args.workload, "scene": ps.name, "width": w, "height": h, "views_per_gpu": n_views, "occluders": n_occ,
                       "quads": int(scene.quads_per_occluder.sum()), "occludees": n_boxes,
                       "outputs": "depth+HiZ+visibility bits per view (HBM resident)" if with_targets else "visibility bits per view",
                       "l2": f"working set per step {n_views * (2 * w * h + 2 * blocks) / 1e6:.0f} MB of per-view depth+HiZ, larger than the 126 MB L2" if with_targets
                             else "L2 flushed by construction: every view clears and rewrites its scratch target",
                       "group_warps": args.group_warps or "auto", "parallelism": f"views dealt round-robin to {world} GPU(s), NCCL all-gather of bitmasks"},
            "mquads_per_sec": quads_submitted * world / (kern / 1e3) / 1e6,
            "queries_per_sec": n_boxes * total_views / (kern / 1e3),
            "kernel_ms_per_step": kern, "step_ms": step_ms,
            "e2e": {"value": e2e_value, "unit": "views/s", "h2d_bytes_per_step": int(mvps.nbytes + poss.nbytes), "d2h_bytes_per_step": int(h_vis.numel() * 4),
                    "note": "depth/HiZ are written to HBM per view and stay device-resident; the caller-visible result is the bitmask"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (traffic or {}).get("dram_bytes_per_launch"), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg,
                         "kernel": ("one step = k_prepare_views + k_sort_views + k_setup_views (speculative setup of every occluder in the frustum) + "
                                    "k_raster_views_cluster (one thread-block cluster per view, dataflow gates, tile-major register-resident depth) + k_query_views; "
                                    if cluster_path else
                                    "one step = k_prepare_views + k_sort_views + 4 x (k_render_views<GW> + k_query_views), the four cost-sorted sub-batches overlapped on four streams; ")
                                   + "duration = CUDA events around the step on the context stream",
                         "traffic_note": (traffic or {}).get("note"),
                         "note": "issue/latency bound by design (SURVEY 8d): HBM is not the limiter; ncu per-launch counters in profiles/"},
        }
        # BASELINE configs[0]/[1]: ONE view (the scene's default camera) through the host entry point: matrix in, bits out
        c = ps.camera
        from rasterizer_b200 import camera as cam
        one_mvp = cam.view_projection(c["pos"], c["dir"], c["up"], c["fov"], w, h)[None]
        one_pos = np.array(c["pos"], np.float32)[None]
        for _ in range(3):
            scene.render_views(w, h, one_mvp, cam_pos=one_pos, want=("vis",))
        t0 = time.perf_counter()
        for _ in range(20):
            scene.render_views(w, h, one_mvp, cam_pos=one_pos, want=("vis",))
        line["single_view"] = {"ms": (time.perf_counter() - t0) / 20 * 1e3, "what": "one view, default camera: host matrix in -> frame loop + all occludee queries -> host bits out (wall clock, includes H2D/D2H and launches)"}
        if not args.no_cpu_baseline and world == 1:
            cb1 = cpu_reference(ps, w, h, one_mvp, one_pos, 1, 1.0)
            line["single_view"]["reference_ms_one_core"] = cb1["frame_ms_per_view"] + cb1["query_ms_per_view"]
        if not args.no_cpu_baseline and world == 1:
            sample = min(128, n_views)
            idx = np.linspace(0, n_views - 1, sample).astype(int)
            cb = cpu_reference(ps, w, h, mvps[idx], poss[idx], 1, args.cpu_seconds)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
            line["cpu_baseline"].update({k: cb[k] for k in ("frame_ms_per_view", "query_ms_per_view", "mquads_per_s", "queries_per_s") if k in cb})
        print(json.dumps(line), flush=True)
    scene.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    # the contract is ONE JSON line on stdout: libraries (NCCL prints its version banner there when NCCL_DEBUG is
    # set) write to fd 1 directly, so fd 1 is pointed at stderr and the line goes to the original stdout
    _real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = _real_stdout
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
