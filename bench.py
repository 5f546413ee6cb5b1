#!/usr/bin/env python3
"""bench.py — headline benchmark of the hot path (contract in the task statement / DESIGN.md §Measurement).

Workload (BASELINE.json metric "Mrays/s (primary + 1 GI bounce) on castle.vox", configs[2]):
castle stand-in, 1920x1080, primary + 1 diffuse GI bounce, 4 spp per GPU.  A step = one frame.
Rays counted = primary + bounce rays actually traced; sun-shadow (NEE) rays are traced but reported
separately.  N GPUs: screen tiles (32x32) are dealt round-robin to ranks, the tree is
replicated, one all-gather of the compact per-rank tile buffers + an un-permute kernel finish the
frame.  Scaling is WEAK by default (spp = 4*N, i.e. every GPU traces as many rays as the N=1 run);
--scaling strong keeps the N=1 frame and splits it.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 2|3]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

--impl reference times the CPU port of the reference semantics (the oracle, all host threads) on a
bounded sample of the same workload: the reference's own path needs Vulkan RT hardware + a Rust
nightly toolchain and cannot be built or run here.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # BASELINE.json configs[1..3]
    1: dict(width=1920, height=1080, spp=1, max_bounces=0, name="castle 1920x1080 primary-only 1spp"),
    2: dict(width=1920, height=1080, spp=4, max_bounces=1, name="castle 1920x1080 primary+1 GI bounce 4spp"),
    3: dict(width=3840, height=2160, spp=8, max_bounces=2, name="castle 3840x2160 primary+2 GI bounces 8spp"),
    # BASELINE.json configs[4]: synthetic 4096^3-domain VDB (1 % of the leaf slots filled), random-origin ray batch
    5: dict(rays=20_000_000, name="synthetic 4096^3 VDB 1% fill, random ray batch"),
}
SEED = 2


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes(stats, pixels_written, rays_in=0, hits_out=0):
    """B = 32*R_in + 16*R_out + 8*W + 4*P + 5*H + 16*pixels."""
    return 32 * rays_in + 16 * hits_out + 8 * stats["mask_words"] + 4 * stats["child_ptrs"] + 5 * stats["hits_shaded"] + 16 * pixels_written


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def castle_setup():
    from dust_b200 import host, scenes
    data, label = scenes.load_asset("castle")
    W = dict(scenes.CASTLE_CAMERA)
    sky = host.sunlight_bake(turbidity=1.0, albedo=(0.2, 0.2, 0.2), direction=host.castle_sun_direction(0.0))
    return data, label, W, sky


# --------------------------------------------------------------------------------------- reference arm

def oracle_scene(data):
    from oracle import pyoracle as po
    ov = po.VoxFile(data)
    models = {m: ov.load_model(m) for m in sorted({m for m, _ in ov.instances})}
    sc = po.Scene([(mat, models[m].tree._h, models[m].materials, ov.palette) for m, mat in ov.instances])
    sc._models, sc._vox = models, ov
    return sc


def cpu_sample_rows(height, n_rows):
    """Rows spread over the frame so the sample sees sky, terrain and castle in proportion."""
    n_rows = max(1, min(height, n_rows))
    return [int((i + 0.5) * height / n_rows) for i in range(n_rows)]


def run_cpu_rows(osc, cam, sky, cfg, rows, spp):
    """Times the oracle on `rows` (ONE o_render_rows call: all host threads share the rows). Returns (rays, seconds)."""
    t0 = time.perf_counter()
    r = osc.render(cam, sky, cfg["width"], cfg["height"], spp=spp, max_bounces=cfg["max_bounces"], seed=SEED, hits=False, rows=rows)
    dt = time.perf_counter() - t0
    return r["stats"]["primary_rays"] + r["stats"]["bounce_rays"], dt


def reference_arm(args, cfg):
    from dust_b200 import host
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    data, label, camkw, sky = castle_setup()
    cam = host.camera_look_at(width=cfg["width"], height=cfg["height"], **camkw)
    osc = oracle_scene(data)
    cores = os.cpu_count() or 1
    spp = cfg["spp"]
    # calibrate: how many rows fit in ~4 s per step
    rays, dt = run_cpu_rows(osc, cam, sky, cfg, cpu_sample_rows(cfg["height"], 4), spp)
    rate = rays / max(dt, 1e-6)
    per_row = rays / 4
    n_rows = int(max(4, min(cfg["height"], 4.0 * rate / max(per_row, 1))))
    rows = cpu_sample_rows(cfg["height"], n_rows)
    for _ in range(args.warmup):
        run_cpu_rows(osc, cam, sky, cfg, rows[: max(4, n_rows // 8)], spp)
    tot_rays, tot_t = 0, 0.0
    for _ in range(args.steps):
        r, t = run_cpu_rows(osc, cam, sky, cfg, rows, spp)
        tot_rays += r; tot_t += t
    value = tot_rays / tot_t / 1e6
    sample = f"{n_rows} of {cfg['height']} rows (evenly spread) of the same frame per step, {spp} spp, all host threads"
    line = {
        "impl": "reference", "metric": "Mrays/s (primary + 1 GI bounce) on castle.vox", "value": value, "unit": "Mrays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(cfg, label, args, spp),
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CPU port of the reference semantics (oracle/); the reference's own Vulkan-RT path cannot run here",
    }
    print(json.dumps(line), flush=True)


def reference_arm_rays(args, cfg):
    """--impl reference --config 5: the oracle port on all host threads, each step a bounded sample of the ray batch."""
    from dust_b200 import host, scenes
    from oracle import pyoracle as po
    if int(os.environ.get("RANK", "0")) != 0:
        return
    dom = host.SynthDomain(seed=5, fill=0.01, domain_leaves=1024)
    desc = dom.describe()
    otree = po.TreeView(desc)
    ident = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], np.float32)
    osc = po.Scene([(ident, otree._h, dom.materials, dom.palette)])
    n = int(args.rays or cfg["rays"])
    rays = scenes.ray_batch(n, dom.extent, seed=5)
    n_cal = min(n, 200_000)
    t0 = time.perf_counter()
    osc.trace_rays(rays[:n_cal])
    rate = n_cal / max(time.perf_counter() - t0, 1e-6)
    n_step = int(max(n_cal, min(n, 4.0 * rate)))                          # ~4 s of CPU work per step
    for _ in range(args.warmup):
        osc.trace_rays(rays[: max(n_cal, n_step // 8)])
    tot_t = 0.0
    for _ in range(args.steps):
        t0 = time.perf_counter()
        osc.trace_rays(rays[:n_step])
        tot_t += time.perf_counter() - t0
    value = args.steps * n_step / tot_t / 1e6
    cores = os.cpu_count() or 1
    sample = f"first {n_step} of {n} rays per step, all host threads"
    print(json.dumps({
        "impl": "reference", "metric": "Mrays/s, random-origin ray batch on the synthetic 4096^3 VDB", "value": value, "unit": "Mrays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"BASELINE configs[4]: {cfg['name']}", "rays_per_gpu": n, "leaves": dom.num_leaves},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CPU port of the reference semantics (oracle/); the reference itself has no ray traversal outside the Vulkan driver"}), flush=True)


def workload_config(cfg, label, args, spp):
    return {"workload": f"BASELINE configs[{args.config}]: {cfg['name']}", "scene": label, "width": cfg["width"], "height": cfg["height"],
            "spp_per_frame": spp, "max_bounces": cfg["max_bounces"], "camera": "examples/castle.rs:126-128", "sun": "zenith (castle.rs:260-265 scroll 0)",
            "sharding": f"32x32 tiles round-robin over {args.gpus} rank(s), tree replicated", "l2": "flushed between timed steps (256 MiB write)",
            "rays_counted": "primary + bounce (sun-shadow rays traced, reported in shadow_rays_per_step)"}


# -------------------------------------------------------------------------------------------- our arm

def ours(args, cfg):
    import torch
    import torch.distributed as dist
    from dust_b200 import host
    from dust_b200.capi import DUST_FMT_RGBA16F
    from dust_b200.loader import load_vox
    from dust_b200.pipeline import Context, resolve_shards, shard_layout

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W, H = cfg["width"], cfg["height"]
    spp = cfg["spp"] * (world if args.scaling == "weak" else 1)

    if world > 1:                                  # rank 0 generates (and caches) the stand-in asset once
        if rank == 0:
            castle_setup()
        dist.barrier()
    data, label, camkw, sky = castle_setup()
    cam = host.camera_look_at(width=W, height=H, **camkw)
    ctx = Context(local)
    t0 = time.perf_counter()
    loaded = load_vox(ctx, data)                                   # trees built on the host (mirror of the Rust crates), then uploaded
    upload_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    loaded_dev = load_vox(ctx, data, device_build=True)            # same trees built on the device from the voxel lists
    ctx.synchronize()
    device_build_s = time.perf_counter() - t0
    del loaded_dev
    scene = loaded.scene
    ctx.use_torch_stream()
    if world > 1:
        # the data-path collective is the library's own entry point (ncclAllGather on the ctx stream);
        # torch.distributed only carries the 128-byte id, the barriers and the max-over-ranks of the timings
        uid = [Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(uid[0], rank, world)

    tiles, per = shard_layout(W, H, world)
    if world > 1:
        shard_buf = torch.empty((per, 4), dtype=torch.float32, device="cuda")
        gathered = torch.empty((world, per, 4), dtype=torch.float32, device="cuda")
        frame = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
    else:
        frame = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    kw = dict(spp=spp, max_bounces=cfg["max_bounces"], seed=SEED)

    def step(frame_index=0, stats=False):
        if world > 1:
            scene.trace_device(cam, sky, W, H, shard_buf, shard_index=rank, shard_count=world, frame_index=frame_index, stats=stats, **kw)
            ctx.allgather_framebuffer(shard_buf, gathered)
            resolve_shards(ctx, W, H, world, gathered, frame)
        else:
            scene.trace_device(cam, sky, W, H, frame, frame_index=frame_index, stats=stats, **kw)

    # ray / visit counts of exactly this workload (one untimed pass with the counting kernel variant)
    step(stats=True)
    torch.cuda.synchronize()
    st = ctx.stats()
    my_pixels = W * H if world == 1 else per
    counts = torch.tensor([st["primary_rays"], st["bounce_rays"], st["shadow_rays"], st["mask_words"], st["child_ptrs"], st["hits_shaded"],
                           algorithmic_bytes(st, my_pixels)], dtype=torch.float64, device="cuda")
    rank_counts = counts.clone()
    if world > 1:
        dist.all_reduce(counts)
    tot = [float(v) for v in counts.tolist()]
    rays_per_step = tot[0] + tot[1]

    for _ in range(max(3, args.warmup)):
        step()
    torch.cuda.synchronize()
    launches0 = ctx.stats()["kernel_launches"]
    clocks = ClockSampler(local)
    clocks.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    for k in range(args.steps):
        flush.fill_(k & 0xFF)                      # evict L2 (126 MB) between timed steps; not inside the event pair
        if world > 1:
            dist.barrier()
        ev[k][0].record()
        if world > 1:
            scene.trace_device(cam, sky, W, H, shard_buf, shard_index=rank, shard_count=world, **kw)
            ev[k][2].record()
            ctx.allgather_framebuffer(shard_buf, gathered)
            resolve_shards(ctx, W, H, world, gathered, frame)
        else:
            scene.trace_device(cam, sky, W, H, frame, **kw)
            ev[k][2].record()
        ev[k][1].record()
    torch.cuda.synchronize()
    launches = ctx.stats()["kernel_launches"] - launches0
    # K short steps can end before nvidia-smi's first 100 ms tick: keep the same load running (untimed) until
    # the sampler has seen at least ~0.8 s of it, so the clocks line describes the load that was timed
    t_load = time.perf_counter()
    while time.perf_counter() - t_load < 0.8:
        step()
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clk = clocks.stop()
    step_ms = sum(a.elapsed_time(b) for a, b, _ in ev)
    kern_ms = sum(a.elapsed_time(c) for a, _, c in ev) / args.steps      # the trace kernel alone (same stream)

    # ---- per-kernel view of the same call, live (not under a profiler): DUST_TRACE_TIMING drops an event behind every launch;
    # one STATS pass gives what each launch added to the visit counters, K timed passes give the device time of each launch
    tgt = shard_buf if world > 1 else frame
    skw = dict(shard_index=rank, shard_count=world) if world > 1 else {}
    scene.trace_device(cam, sky, W, H, tgt, stats=True, timing=True, **skw, **kw)
    per_launch = ctx.launch_info()
    for rec in per_launch:
        rec["ms"] = 0.0
    for k in range(args.steps):
        flush.fill_(k & 0xFF)
        scene.trace_device(cam, sky, W, H, tgt, timing=True, **skw, **kw)
        for rec, t_rec in zip(per_launch, ctx.launch_info()):
            assert rec["kernel"] == t_rec["kernel"]
            rec["ms"] += t_rec["ms"] / args.steps
    t = torch.tensor([step_ms, kern_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    step_ms_max, kern_ms_max = [float(v) for v in t.tolist()]
    ms_per_step = step_ms_max / args.steps
    value = rays_per_step / (ms_per_step * 1e-3) / 1e6

    # ---- e2e: the public host-buffer call (dust_trace, DUST_MEM_HOST): kernel + D2H of the RGBA16F frame into pinned memory
    ctx.set_stream(None)
    n_out = W * H if world == 1 else per
    pinned = torch.empty((n_out, 4), dtype=torch.float16).pin_memory()
    out_np = pinned.numpy()
    e2e_kw = dict(spp=spp, max_bounces=cfg["max_bounces"], seed=SEED, fmt=DUST_FMT_RGBA16F, out=out_np, shard_index=rank, shard_count=world)
    scene.trace(cam, sky, W, H, **e2e_kw)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        scene.trace(cam, sky, W, H, frame_index=0, **e2e_kw)
        float(out_np[0, 0])                                       # host read of the result
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = rays_per_step / (float(te.item()) / args.steps) / 1e6
    h2d = 192 + 224 + 36                                           # DustCamera + DustSky + DustTraceConfig, sent as launch parameters
    d2h = n_out * 8

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline: algorithmic bytes of THIS rank's frame / the device time of the kernels of one dust_trace call
    # (one launch for primary-only frames; the wavefront pipeline's launches for frames with bounces — events on
    # the launching stream bracket all of them; per-kernel shares: profiles/).
    peak, peak_src = measured_peak()
    alg_bytes = float(rank_counts[6].item())
    achieved = alg_bytes / (kern_ms_max * 1e-3) / 1e9
    traffic, traffic_note = None, None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            traffic = tj.get("config%d_dram_bytes_per_step" % args.config)
            traffic_note = tj.get("how")
        except Exception:
            traffic = None
    launches_per_step = launches / max(1, args.steps) - (1 if world > 1 else 0)     # resolve_kernel follows the all-gather
    pipeline = "frame_kernel (megakernel, 1 launch)" if cfg["max_bounces"] == 0 else \
        "wavefront pipeline: wf_trace, wf_shade, wf_cull, wf_trace_pt, wf_resolve (%d launches)" % launches_per_step
    kernels, total_ms = [], sum(r["ms"] for r in per_launch) or 1.0
    for r in per_launch:
        b = 8 * r["mask_words"] + 4 * r["child_ptrs"] + 5 * r["hits_shaded"] + (16 * my_pixels if r["kernel"] in ("wf_resolve", "frame_kernel") else 0)
        kernels.append({"kernel": r["kernel"], "ms": round(r["ms"], 4), "share": round(r["ms"] / total_ms, 4), "algorithmic_bytes": b,
                        "achieved_GBps": round(b / (r["ms"] * 1e-3) / 1e9, 1) if r["ms"] > 0 else None,
                        "rays": r["primary_rays"] + r["bounce_rays"] + r["shadow_rays"]})
    dom = max(kernels, key=lambda k_: k_["ms"]) if kernels else None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "dominant_kernel": None if dom is None else dict(dom, frac=(dom["achieved_GBps"] or 0.0) / peak,
                                                                 note="device time between the CUDA events behind consecutive launches of a live dust_trace call (DUST_TRACE_TIMING), mean of the timed steps; bytes = 8 W + 4 P + 5 H of that launch"),
                "kernels": kernels,
                "kernel": "dust_trace = " + pipeline, "kernel_ms": kern_ms_max, "algorithmic_bytes_per_launch": alg_bytes,
                "bytes_per_ray": alg_bytes / max(1.0, float(rank_counts[0].item() + rank_counts[1].item())), "peak_source": peak_src,
                "traffic_source": traffic_note,
                "note": "scene (tree %.1f MB) is L2-resident after first touch: the traversal kernels are instruction-issue bound "
                        "(70-80 %% issue slots busy), not HBM-bound; see profiles/" % (loaded.device_bytes / 1e6)}

    # ---- CPU baseline beside it: oracle port on all host threads, bounded sample of the same frame
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        osc = oracle_scene(data)
        cores = os.cpu_count() or 1
        rows = cpu_sample_rows(H, 8)
        r, dt = run_cpu_rows(osc, cam, sky, cfg, rows, spp)
        n_rows = int(max(8, min(H, 15.0 * (r / max(dt, 1e-6)) / max(r / 8, 1))))
        rows = cpu_sample_rows(H, n_rows)
        r, dt = run_cpu_rows(osc, cam, sky, cfg, rows, spp)
        cpu = {"value": r / dt / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
               "sample": f"{n_rows} of {H} rows (evenly spread) of the same frame, {spp} spp, {dt:.1f} s"}

    line = {
        "metric": "Mrays/s (primary + 1 GI bounce) on castle.vox", "value": value, "unit": "Mrays/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(cfg, label, args, spp),
        "rays_per_step": rays_per_step, "shadow_rays_per_step": tot[2],
        "scene_stats": {"voxels": loaded.num_voxels, "leaves": loaded.num_leaves, "instances": len(loaded.instances),
                        "device_tree_bytes": loaded.device_bytes, "upload_s": upload_s,
                        "device_build_s": device_build_s},
        "roofline": roofline, "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "dust_trace(DUST_MEM_HOST, RGBA16F) -> pinned host frame" + ("" if world == 1 else " (every rank downloads its own shard; no gather on this path)")},
        "gpu_launches": int(launches), "clocks": clk,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def rays_config(args, cfg):
    """--config 5: the out-of-L2 case.  A step = one dust_trace_rays call over the whole batch (device buffers);
    e2e = the same call with host buffers (H2D of the rays, D2H of the hits inside the timed region)."""
    import torch
    from dust_b200 import host, scenes
    from dust_b200.pipeline import Context, Scene, HIT_DTYPE
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = int(args.rays or cfg["rays"])
    t0 = time.perf_counter()
    dom = host.SynthDomain(seed=5, fill=0.01, domain_leaves=1024)
    ctx = Context(local)
    dtree = ctx.upload_tree(dom.describe())
    ident = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], np.float32)
    scene = Scene(ctx, [(ident, dtree)])
    build_s = time.perf_counter() - t0
    rays = scenes.ray_batch(n, dom.extent, seed=5 + rank)            # weak scaling: every rank its own batch, tree replicated
    ctx.use_torch_stream()
    d_rays = torch.from_numpy(rays.view(np.float32).reshape(-1, 8)).cuda()
    d_hits = torch.empty((n, 4), dtype=torch.int32, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    scene.trace_rays_device(d_rays, d_hits, n, stats=True)
    torch.cuda.synchronize()
    st = ctx.stats()
    alg = algorithmic_bytes(st, 0, rays_in=n, hits_out=n)
    for _ in range(max(3, args.warmup)):
        scene.trace_rays_device(d_rays, d_hits, n)
    torch.cuda.synchronize()
    launches0 = ctx.stats()["kernel_launches"]
    clocks = ClockSampler(local)
    clocks.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    for k in range(args.steps):
        flush.fill_(k & 0xFF)
        ev[k][0].record()
        scene.trace_rays_device(d_rays, d_hits, n)
        ev[k][1].record()
    torch.cuda.synchronize()
    launches = ctx.stats()["kernel_launches"] - launches0
    t_load = time.perf_counter()
    while time.perf_counter() - t_load < 0.8:                       # see ours(): enough nvidia-smi samples under the same load
        scene.trace_rays_device(d_rays, d_hits, n)
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clk = clocks.stop()
    ms = sum(a.elapsed_time(b) for a, b in ev) / args.steps
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    # e2e through the host-buffer call
    ctx.set_stream(None)
    n_e2e = n
    rays_pin = torch.from_numpy(rays.view(np.uint8).reshape(-1)).pin_memory()           # pinned host memory: the copies of
    hits_pin = torch.empty(n_e2e * 16, dtype=torch.uint8).pin_memory()                   # consecutive chunks overlap the kernels
    rays_host = rays_pin.numpy().view(rays.dtype)
    hits_host = hits_pin.numpy().view(HIT_DTYPE)
    scene.trace_rays(rays_host, out=hits_host)
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 3))
    for _ in range(e2e_steps):
        scene.trace_rays(rays_host, out=hits_host)
        int(hits_host["instance"][0])
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    e2e_same = bool(np.array_equal(hits_host.view(np.uint32).reshape(-1, 4), d_hits.cpu().numpy().view(np.uint32)))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    peak, peak_src = measured_peak()
    achieved = alg / (ms * 1e-3) / 1e9
    hit_frac = float((d_hits[:, 1] != -1).float().mean().item())
    # ---- CPU baseline beside it: the oracle port on all host threads over the first rays of the same batch,
    # and (it is the checker, after all) a bit-for-bit comparison of those rays with the GPU's answers
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import pyoracle as po
        desc = dom.describe()
        otree = po.TreeView(desc)                                   # a view of the same host tree (kept alive with the scene)
        osc = po.Scene([(ident, otree._h, dom.materials, dom.palette)])
        n_cal = min(n, 100_000)
        t0 = time.perf_counter()
        osc.trace_rays(rays[:n_cal])
        rate = n_cal / max(time.perf_counter() - t0, 1e-6)
        n_cpu = int(max(n_cal, min(n, 12.0 * rate)))
        t0 = time.perf_counter()
        o_hits, _ = osc.trace_rays(rays[:n_cpu])
        dt = time.perf_counter() - t0
        same = bool(np.array_equal(o_hits.view(np.uint32).reshape(-1, 4), d_hits[:n_cpu].cpu().numpy().view(np.uint32)))
        cpu = {"value": n_cpu / dt / 1e6, "unit": "Mrays/s", "cores": os.cpu_count() or 1, "kind": "port",
               "sample": f"first {n_cpu} rays of the same batch, {dt:.1f} s", "gpu_hits_bit_exact_on_sample": same}
    line = {
        "metric": "Mrays/s, random-origin ray batch on the synthetic 4096^3 VDB", "value": world * n / (ms * 1e-3) / 1e6, "unit": "Mrays/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"BASELINE configs[4]: {cfg['name']}", "rays_per_gpu": n, "leaves": dom.num_leaves, "tree_device_bytes": dtree.device_bytes,
                   "hit_fraction": hit_frac, "l2": "flushed between timed steps (256 MiB write); the tree (%.0f MB) and the rays exceed L2" % (dtree.device_bytes / 1e6),
                   "build_s": build_s},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                     "kernel": "wf_trace_pt<4> (persistent lanes, %d launch(es))" % (launches // args.steps), "kernel_ms": ms,
                     "algorithmic_bytes_per_launch": alg, "bytes_per_ray": alg / n, "peak_source": peak_src},
        "cpu_baseline": cpu,
        "e2e": {"value": n_e2e / e2e_s / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": n_e2e * 32, "d2h_bytes_per_step": n_e2e * 16,
                "api": "dust_trace_rays(DUST_MEM_HOST), %d rays per call from / to pinned host memory, 2 Mi-ray chunks on two streams" % n_e2e,
                "hits_equal_device_path": e2e_same},
        "gpu_launches": int(launches), "clocks": clk,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--rays", type=int, default=0, help="--config 5: rays per GPU (default 2e7; BASELINE names 1e8)")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.config == 5:
        if args.impl == "reference":
            reference_arm_rays(args, cfg)
        else:
            rays_config(args, cfg)
        return
    if args.impl == "reference":
        reference_arm(args, cfg)
    else:
        ours(args, cfg)


if __name__ == "__main__":
    main()
