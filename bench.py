#!/usr/bin/env python3
"""bench.py — headline benchmark of the hot path (contract in the task statement / DESIGN.md §6).

Headline workload = BASELINE.json configs[2] (metric "Mrays/s (primary + 1 GI bounce) on castle.vox"): castle stand-in,
1920x1080, primary + 1 diffuse GI bounce, 4 spp per GPU.  A step = one frame.  Rays counted = primary + bounce rays
actually traced; sun-shadow (NEE) rays are traced but reported separately.  N GPUs: 32x32 screen tiles are
dealt round-robin to the ranks, the tree is replicated, ONE all-gather of the compact per-rank tile buffers + an un-permute
kernel finish the frame.

The one JSON line carries, beside the contract's keys:
  value / scaling   WEAK scaling (spp = 4 N: every GPU traces as many rays as the N=1 run)
  strong            N > 1: the same N=1 frame (4 spp) split over the N ranks — value, ms, where the time goes
  parity            N = 1: the CPU frame rendered for cpu_baseline compared with the GPU frame (hit ids / t bit-exact, radiance rel.)
  sharded_equals_whole   N > 1: the NCCL-gathered frame against the same frame rendered whole on rank 0, bit for bit
  e2e               N = 1: dust_trace_async -> pinned host frame, 3 frames in flight; N > 1: trace + all-gather + un-permute + D2H
                    of the FULL frame on rank 0
  extra             records of the other BASELINE configs the driver cannot ask for by flag: configs[3] (3840x2160, 2 bounces,
                    8 spp; split over the ranks) and configs[4] (10^8 pcg rays on the 4096^3 synthetic VDB; split over the ranks),
                    each with value, roofline, e2e, parity; and (N = 1) "gi": one frame of the reference's own GI (radiance cache +
                    surfels) at the reference's sizes, with its parity against the oracle

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 1|2|3|4] [--no-extras]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

--config takes the INDEX into BASELINE.json `configs` (configs[0] is the CPU-only teapot case; 5 is accepted as an alias of 4
for round-1 command lines).  --impl reference times the CPU port of the reference semantics (the oracle, all host threads) on
a bounded sample of the same workload: the reference's own path needs Vulkan RT hardware + a Rust nightly toolchain and
cannot be built or run here.
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
    # index into BASELINE.json `configs`
    1: dict(width=1920, height=1080, spp=1, max_bounces=0, name="castle 1920x1080 primary-only 1spp"),
    2: dict(width=1920, height=1080, spp=4, max_bounces=1, name="castle 1920x1080 primary+1 GI bounce 4spp"),
    3: dict(width=3840, height=2160, spp=8, max_bounces=2, name="castle 3840x2160 primary+2 GI bounces 8spp"),
    4: dict(rays=100_000_000, name="synthetic 4096^3 VDB 1% fill, random-origin ray batch 1e8 (pcg)"),
}
SEED = 2
RAY_SEED = 5
METRIC = "Mrays/s (primary + 1 GI bounce) on castle.vox"


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes(stats, pixels_written, rays_in=0, hits_out=0, strict=False):
    """B = 32*R_in + 16*R_out + 8*W + 4*P + 5*H + 16*pixels.  strict: W by the letter of §8d (node words tested +
    mid nodes entered + leaves entered); default: every 64-bit word the walk consults, block masks and cached nodes included."""
    w = stats["mask_words_strict"] if strict else stats["mask_words"]
    return 32 * rays_in + 16 * hits_out + 8 * w + 4 * stats["child_ptrs"] + 5 * stats["hits_shaded"] + 16 * pixels_written


def traffic_record(key):
    """dram__bytes of one step from the committed ncu capture (profiles/traffic.json), or (None, None)."""
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        tj = json.load(open(tp))
        return tj.get(key), tj.get("how")
    except Exception:
        return None, None


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


def workload_config(cfg, cfg_id, label, world, spp, scaling):
    return {"workload": f"BASELINE.json configs[{cfg_id}]: {cfg['name']}", "scene": label, "width": cfg["width"], "height": cfg["height"],
            "spp_per_frame": spp, "max_bounces": cfg["max_bounces"], "camera": "examples/castle.rs:126-128", "sun": "zenith (castle.rs:260-265 scroll 0)",
            "sharding": f"32x32 tiles round-robin over {world} rank(s), tree replicated ({scaling})", "l2": "flushed between timed steps (256 MiB write)",
            "rays_counted": "primary + bounce (sun-shadow rays traced, reported in shadow_rays_per_step)"}


# --------------------------------------------------------------------------------------- reference arm (CPU port)

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


def run_cpu_rows(osc, cam, sky, cfg, rows, spp, hits=False):
    """Times the oracle on `rows` (ONE o_render_rows call: all host threads share the rows). Returns (rays, seconds, result)."""
    t0 = time.perf_counter()
    r = osc.render(cam, sky, cfg["width"], cfg["height"], spp=spp, max_bounces=cfg["max_bounces"], seed=SEED, hits=hits, rows=rows)
    dt = time.perf_counter() - t0
    return r["stats"]["primary_rays"] + r["stats"]["bounce_rays"], dt, r


def reference_arm(args, cfg, cfg_id):
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
    rays, dt, _ = run_cpu_rows(osc, cam, sky, cfg, cpu_sample_rows(cfg["height"], 4), spp)
    rate = rays / max(dt, 1e-6)
    per_row = rays / 4
    n_rows = int(max(4, min(cfg["height"], 4.0 * rate / max(per_row, 1))))
    rows = cpu_sample_rows(cfg["height"], n_rows)
    for _ in range(args.warmup):
        run_cpu_rows(osc, cam, sky, cfg, rows[: max(4, n_rows // 8)], spp)
    tot_rays, tot_t = 0, 0.0
    for _ in range(args.steps):
        r, t, _ = run_cpu_rows(osc, cam, sky, cfg, rows, spp)
        tot_rays += r; tot_t += t
    value = tot_rays / tot_t / 1e6
    sample = f"{n_rows} of {cfg['height']} rows (evenly spread) of the same frame per step, {spp} spp, all host threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mrays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(cfg, cfg_id, label, args.gpus, spp, "weak"),
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CPU port of the reference semantics (oracle/: scalar C, instances visited front to back by box entry time); "
                "the reference's own Vulkan-RT path cannot run here",
    }
    print(json.dumps(line), flush=True)


def reference_arm_rays(args, cfg, cfg_id):
    """--impl reference --config 4: the oracle port on all host threads, each step a bounded sample of the ray batch."""
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
    n_cal = min(n, 200_000)
    cal = scenes.ray_batch_pcg(n_cal, dom.extent, seed=RAY_SEED)
    t0 = time.perf_counter()
    osc.trace_rays(cal)
    rate = n_cal / max(time.perf_counter() - t0, 1e-6)
    n_step = int(max(n_cal, min(n, 4.0 * rate)))                          # ~4 s of CPU work per step
    rays = scenes.ray_batch_pcg(n_step, dom.extent, seed=RAY_SEED)
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
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"BASELINE.json configs[{cfg_id}]: {cfg['name']}", "rays": n, "leaves": dom.num_leaves},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CPU port of the reference semantics (oracle/); the reference itself has no ray traversal outside the Vulkan driver"}), flush=True)


# -------------------------------------------------------------------------------------------- our arm

class Env:
    def __init__(self, args):
        import torch
        self.torch = torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
        torch.cuda.set_device(self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier(self):
        if self.dist:
            self.dist.barrier()

    def allmax(self, vals):
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device="cuda")
        if self.dist:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]

    def allsum(self, vals):
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device="cuda")
        if self.dist:
            self.dist.all_reduce(t)
        return [float(v) for v in t.tolist()]

    def close(self):
        if self.dist:
            self.dist.destroy_process_group()


def load_castle(env):
    """The castle scene on this rank's GPU (host build + upload, and the device build timed beside it) + the NCCL communicator."""
    from dust_b200.loader import load_vox
    from dust_b200.pipeline import Context
    if env.world > 1:                                  # rank 0 generates (and caches) the stand-in asset once
        if env.rank == 0:
            castle_setup()
        env.barrier()
    data, label, camkw, sky = castle_setup()
    ctx = Context(env.local)
    t0 = time.perf_counter()
    loaded = load_vox(ctx, data)                                   # trees built on the host (mirror of the Rust crates), then uploaded
    upload_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    loaded_dev = load_vox(ctx, data, device_build=True)            # same trees built on the device from the voxel lists
    ctx.synchronize()
    device_build_s = time.perf_counter() - t0
    del loaded_dev
    ctx.use_torch_stream()
    if env.world > 1:
        # the data-path collective is the library's own entry point (ncclAllGather on the ctx stream);
        # torch.distributed only carries the 128-byte id, the barriers and the max-over-ranks of the timings
        uid = [Context.comm_unique_id() if env.rank == 0 else None]
        env.dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(uid[0], env.rank, env.world)
    return dict(ctx=ctx, loaded=loaded, scene=loaded.scene, data=data, label=label, camkw=camkw, sky=sky,
                upload_s=upload_s, device_build_s=device_build_s)


def radiance_rel_err(a, b):
    """max over pixels of |a - b| relative to the pixel's own magnitude (floor: 1e-3 of the frame mean) — the tests' measure."""
    a, b = a[..., :3].astype(np.float64), b[..., :3].astype(np.float64)
    scale = np.maximum(np.abs(b).max(axis=-1, keepdims=True), 1e-3 * max(np.abs(b).mean(), 1e-30))
    return float((np.abs(a - b) / scale).max())


def castle_record(env, S, cfg, cfg_id, scaling, steps, warmup, *, kernels=True, cpu=True, e2e=True, clocks=True, max_cpu_s=15.0):
    """One castle workload on `world` GPUs -> the record (rank 0) or None.  scaling: 'weak' (spp x world) | 'strong' (the N=1 frame split)."""
    torch = env.torch
    from dust_b200 import host
    from dust_b200.capi import DUST_FMT_RGBA16F
    from dust_b200.pipeline import resolve_shards, shard_layout
    world, rank = env.world, env.rank
    ctx, scene, sky = S["ctx"], S["scene"], S["sky"]
    W, H = cfg["width"], cfg["height"]
    spp = cfg["spp"] * (world if scaling == "weak" else 1)
    cam = host.camera_look_at(width=W, height=H, **S["camkw"])
    ctx.use_torch_stream()
    tiles, per = shard_layout(W, H, world)
    frame = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
    if world > 1:
        shard_buf = torch.empty((per, 4), dtype=torch.float32, device="cuda")
        gathered = torch.empty((world, per, 4), dtype=torch.float32, device="cuda")
    kw = dict(spp=spp, max_bounces=cfg["max_bounces"], seed=SEED)
    skw = dict(shard_index=rank, shard_count=world) if world > 1 else {}
    tgt = shard_buf if world > 1 else frame

    def step(stats=False):
        scene.trace_device(cam, sky, W, H, tgt, stats=stats, **skw, **kw)
        if world > 1:
            ctx.allgather_framebuffer(shard_buf, gathered)
            resolve_shards(ctx, W, H, world, gathered, frame)

    # ray / visit counts of exactly this workload (one untimed pass with the counting kernel variant)
    step(stats=True)
    torch.cuda.synchronize()
    st = ctx.stats()
    my_pixels = W * H if world == 1 else per
    mine = [st["primary_rays"], st["bounce_rays"], st["shadow_rays"], st["mask_words"], st["child_ptrs"], st["hits_shaded"],
            algorithmic_bytes(st, my_pixels), algorithmic_bytes(st, my_pixels, strict=True)]
    tot = env.allsum(mine)
    rays_per_step = tot[0] + tot[1]

    for _ in range(max(3, warmup)):
        step()
    torch.cuda.synchronize()
    launches0 = ctx.stats()["kernel_launches"]
    sampler = ClockSampler(env.local) if clocks else None
    if sampler:
        sampler.start()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(steps)]
    env.barrier()
    torch.cuda.synchronize()
    for k in range(steps):
        env.flush.fill_(k & 0xFF)                  # evict L2 (126 MB) between timed steps; not inside the event pair
        env.barrier()
        ev[k][0].record()
        scene.trace_device(cam, sky, W, H, tgt, **skw, **kw)
        ev[k][1].record()
        if world > 1:
            ctx.allgather_framebuffer(shard_buf, gathered)
            ev[k][2].record()
            resolve_shards(ctx, W, H, world, gathered, frame)
        ev[k][3].record()
    torch.cuda.synchronize()
    launches = ctx.stats()["kernel_launches"] - launches0
    if sampler:
        # K short steps can end before nvidia-smi's first 100 ms tick: keep the same load running (untimed) until
        # the sampler has seen at least ~0.8 s of it, so the clocks line describes the load that was timed
        t_load = time.perf_counter()
        while time.perf_counter() - t_load < 0.8:
            step()
            torch.cuda.synchronize()
    env.barrier()
    clk = sampler.stop() if sampler else None
    step_ms = sum(e[0].elapsed_time(e[3]) for e in ev) / steps
    kern_ms = sum(e[0].elapsed_time(e[1]) for e in ev) / steps           # the trace kernels alone (same stream)
    gather_ms = sum(e[1].elapsed_time(e[2]) for e in ev) / steps if world > 1 else 0.0
    resolve_ms = sum(e[2].elapsed_time(e[3]) for e in ev) / steps if world > 1 else 0.0
    # max over ranks of the step; the parts are reported for the slowest-tracing rank's view (max of each)
    step_ms_max, kern_ms_max, gather_ms_max, resolve_ms_max, kern_ms_min = env.allmax([step_ms, kern_ms, gather_ms, resolve_ms, -kern_ms])
    kern_ms_min = -kern_ms_min
    value = rays_per_step / (step_ms_max * 1e-3) / 1e6

    # ---- per-kernel view of the same call, live (not under a profiler): DUST_TRACE_TIMING drops an event behind every launch;
    # one STATS pass gives what each launch added to the visit counters, K timed passes give the device time of each launch
    per_launch = []
    if kernels:
        scene.trace_device(cam, sky, W, H, tgt, stats=True, timing=True, **skw, **kw)
        per_launch = ctx.launch_info()
        for rec in per_launch:
            rec["ms"] = 0.0
        for k in range(steps):
            env.flush.fill_(k & 0xFF)
            scene.trace_device(cam, sky, W, H, tgt, timing=True, **skw, **kw)
            for rec, t_rec in zip(per_launch, ctx.launch_info()):
                assert rec["kernel"] == t_rec["kernel"]
                rec["ms"] += t_rec["ms"] / steps

    # ---- sharded == whole (N > 1): the gathered frame of the LAST step against the same frame traced whole on rank 0
    sharded_equals_whole = None
    if world > 1:
        step()
        torch.cuda.synchronize()
        if rank == 0:
            whole = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
            scene.trace_device(cam, sky, W, H, whole, **kw)
            torch.cuda.synchronize()
            sharded_equals_whole = bool(torch.equal(whole.view(torch.int32), frame.view(torch.int32)))
            del whole
        env.barrier()

    # ---- e2e through the public host-buffer API
    e2e_rec = None
    if e2e:
        ctx.set_stream(None)
        if world == 1:
            # dust_trace_async: kernels on the context's stream, the download of frame k on the copy stream while frame k+1 traces
            # (the reference keeps 3 frames in flight the same way); every frame's result is read on the host before its buffer is reused
            pins = [torch.empty((W * H, 4), dtype=torch.float16).pin_memory() for _ in range(3)]
            outs = [p.numpy() for p in pins]
            e2e_kw = dict(spp=spp, max_bounces=cfg["max_bounces"], seed=SEED, fmt=DUST_FMT_RGBA16F)
            for b in range(3):                                     # warm every slot (each owns a staging buffer, allocated at first use)
                scene.frame_wait(scene.trace_async(cam, sky, W, H, outs[b], **e2e_kw))
            acc = 0.0
            t0 = time.perf_counter()
            pending = []
            for k in range(steps):
                pending.append((scene.trace_async(cam, sky, W, H, outs[k % 3], **e2e_kw), k % 3))
                if len(pending) == 3:
                    tk, b = pending.pop(0)
                    scene.frame_wait(tk)
                    acc += float(outs[b][0, 0])                    # host read of the result
            for tk, b in pending:
                scene.frame_wait(tk)
                acc += float(outs[b][0, 0])
            e2e_s = time.perf_counter() - t0
            d2h = W * H * 8
            api = "dust_trace_async(DUST_MEM_HOST, RGBA16F) -> pinned host frames, 3 in flight, dust_frame_wait + host read per frame"
            # the synchronous call, for comparison (kernels, then the download, then the host continues)
            out1 = outs[0]
            scene.trace(cam, sky, W, H, out=out1, **e2e_kw)
            t0 = time.perf_counter()
            for k in range(steps):
                scene.trace(cam, sky, W, H, out=out1, **e2e_kw)
                acc += float(out1[0, 0])
            sync_s = time.perf_counter() - t0
        else:
            # N ranks: trace the shard, all-gather, un-permute, and rank 0 downloads the FULL frame into pinned memory
            ctx.use_torch_stream()
            pin = torch.empty((H, W, 4), dtype=torch.float32).pin_memory() if rank == 0 else None
            step(); torch.cuda.synchronize(); env.barrier()
            t0 = time.perf_counter()
            acc = 0.0
            for k in range(steps):
                step()
                if rank == 0:
                    pin.copy_(frame, non_blocking=True)
                torch.cuda.synchronize()
                if rank == 0:
                    acc += float(pin[0, 0, 0])
            e2e_s = time.perf_counter() - t0
            sync_s = None
            d2h = W * H * 16
            api = "dust_trace(shard, device) + dust_allgather_framebuffer + dust_resolve_shards on every rank, D2H of the full RGBA32F frame into pinned memory on rank 0"
        e2e_s_max = env.allmax([e2e_s])[0]
        e2e_rec = {"value": rays_per_step / (e2e_s_max / steps) / 1e6, "unit": "Mrays/s",
                   "h2d_bytes_per_step": 192 + 224 + 36,           # DustCamera + DustSky + DustTraceConfig, sent as launch parameters
                   "d2h_bytes_per_step": d2h, "api": api}
        if sync_s is not None:
            e2e_rec["synchronous_call_value"] = rays_per_step / (sync_s / steps) / 1e6
        ctx.use_torch_stream()

    if rank != 0:
        return None

    # ---- roofline: algorithmic bytes of THIS rank's frame / the device time of the kernels of one dust_trace call
    peak, peak_src = measured_peak()
    alg_bytes, alg_strict = float(mine[6]), float(mine[7])
    achieved = alg_bytes / (kern_ms_max * 1e-3) / 1e9
    traffic, traffic_note = traffic_record("config%d_dram_bytes_per_step" % cfg_id) if world == 1 else (None, None)
    n_trace_launches = launches / max(1, steps) - (1 if world > 1 else 0)           # resolve_kernel follows the all-gather
    pipeline = "frame_kernel (megakernel, 1 launch)" if cfg["max_bounces"] == 0 else \
        "wavefront pipeline: wf_tile_cull, wf_sun_grid, wf_trace, wf_shade, wf_cull, wf_trace_pt, wf_trace_slow, wf_resolve (%d launches)" % n_trace_launches
    klist, total_ms = [], sum(r["ms"] for r in per_launch) or 1.0
    for r in per_launch:
        px = 16 * my_pixels if r["kernel"] in ("wf_resolve", "frame_kernel") else 0
        b = 8 * r["mask_words"] + 4 * r["child_ptrs"] + 5 * r["hits_shaded"] + px
        bs = 8 * r["mask_words_strict"] + 4 * r["child_ptrs"] + 5 * r["hits_shaded"] + px
        klist.append({"kernel": r["kernel"], "ms": round(r["ms"], 4), "share": round(r["ms"] / total_ms, 4), "algorithmic_bytes": b,
                      "algorithmic_bytes_strict": bs,
                      "achieved_GBps": round(b / (r["ms"] * 1e-3) / 1e9, 1) if r["ms"] > 0 else None,
                      "rays": r["primary_rays"] + r["bounce_rays"] + r["shadow_rays"]})
    dom = max(klist, key=lambda k_: k_["ms"]) if klist else None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "strict": {"algorithmic_bytes_per_launch": alg_strict, "achieved": alg_strict / (kern_ms_max * 1e-3) / 1e9,
                           "frac": alg_strict / (kern_ms_max * 1e-3) / 1e9 / peak,
                           "note": "W by the letter of SURVEY.md 8d: node words tested + mid nodes entered + leaves entered; the headline count also has "
                                   "the block-mask words and the words of a node already in registers"},
                "dominant_kernel": None if dom is None else dict(
                    dom, frac=(dom["achieved_GBps"] or 0.0) / peak,
                    frac_strict=(dom["algorithmic_bytes_strict"] / (dom["ms"] * 1e-3) / 1e9 / peak) if dom["ms"] > 0 else None,
                    note="device time between the CUDA events behind consecutive launches of a live dust_trace call (DUST_TRACE_TIMING), mean of the timed steps; bytes = 8 W + 4 P + 5 H of that launch"),
                "kernels": klist,
                "kernel": "dust_trace = " + pipeline, "kernel_ms": kern_ms_max, "algorithmic_bytes_per_launch": alg_bytes,
                "bytes_per_ray": alg_bytes / max(1.0, float(mine[0] + mine[1])), "peak_source": peak_src,
                "traffic_source": traffic_note,
                "note": "scene (tree %.1f MB) is L2-resident after first touch: the traversal kernels are bound by the ALU pipe "
                        "(integer / compare / select), not by HBM; see profiles/" % (S["loaded"].device_bytes / 1e6)}

    # ---- CPU baseline beside it (N = 1): the oracle port on all host threads over a bounded sample of the same frame — and, since
    # those pixels exist anyway, the parity check of the frame the numbers above were measured on
    cpu_rec, parity = None, None
    if world == 1 and cpu:
        osc = oracle_scene(S["data"])
        cores = os.cpu_count() or 1
        rows = cpu_sample_rows(H, 8)
        r, dt, _ = run_cpu_rows(osc, cam, sky, cfg, rows, spp)
        n_rows = int(max(8, min(H, max_cpu_s * (r / max(dt, 1e-6)) / max(r / 8, 1))))
        rows = cpu_sample_rows(H, n_rows)
        r, dt, o = run_cpu_rows(osc, cam, sky, cfg, rows, spp, hits=True)
        cpu_rec = {"value": r / dt / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
                   "sample": f"{n_rows} of {H} rows (evenly spread) of the same frame, {spp} spp, {dt:.1f} s",
                   "note": "scalar C port, instances visited front to back by box entry time (no SIMD): a checker first, a baseline second"}
        ctx.set_stream(None)
        g = scene.trace(cam, sky, W, H, spp=spp, max_bounces=cfg["max_bounces"], seed=SEED, hits=True)
        ctx.use_torch_stream()
        ra = np.asarray(rows)
        ids_ok = all(bool(np.array_equal(g[k][ra], o[k][ra])) for k in ("hit_instance", "hit_leaf", "hit_voxel"))
        t_ok = bool(np.array_equal(g["hit_t"][ra].view(np.uint32), o["hit_t"][ra].view(np.uint32)))
        rel = radiance_rel_err(g["radiance"][ra], o["radiance"][ra])
        same_as_timed = bool(np.array_equal(g["radiance"].view(np.uint32), frame.cpu().numpy().view(np.uint32)))
        parity = {"rows": n_rows, "pixels": int(n_rows * W), "hit_pixels": int((o["hit_instance"][ra] != 0xFFFFFFFF).sum()),
                  "hits_bit_exact": ids_ok and t_ok, "radiance_max_rel": rel, "radiance_tolerance": 1e-4,
                  "ok": bool(ids_ok and t_ok and rel <= 1e-4),
                  "host_call_frame_equals_timed_device_frame": same_as_timed,
                  "against": "oracle/ (CPU port of the traversal semantics, DESIGN.md 2), same camera, seed and samples"}

    rec = {
        "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": steps, "warmup": max(3, warmup), "ms_per_step": step_ms_max,
        "scaling": scaling, "config": workload_config(cfg, cfg_id, S["label"], world, spp, scaling),
        "rays_per_step": rays_per_step, "shadow_rays_per_step": tot[2],
        "roofline": roofline, "cpu_baseline": cpu_rec, "parity": parity, "e2e": e2e_rec, "gpu_launches": int(launches), "clocks": clk,
    }
    if world > 1:
        rec["sharded_equals_whole"] = sharded_equals_whole
        rec["step_parts_ms"] = {"trace_max": kern_ms_max, "trace_min": kern_ms_min, "allgather_max": gather_ms_max, "resolve_max": resolve_ms_max,
                                "limiter": max((("tile imbalance + trace", kern_ms_max), ("ncclAllGather", gather_ms_max), ("resolve_kernel", resolve_ms_max)),
                                               key=lambda kv: kv[1])[0],
                                "tile_imbalance": (kern_ms_max - kern_ms_min) / max(kern_ms_max, 1e-9),
                                "note": "allgather_max includes waiting for the slowest rank's trace (the collective starts when every rank arrived)"}
    return rec


def rays_record(env, n_total, steps, warmup, *, cpu=True, e2e=True, clocks=True, cfg_id=4):
    """BASELINE.json configs[4]: the out-of-L2 case.  The 10^8-ray batch is split contiguously over the ranks (tree replicated, no
    collective).  A step = one dust_trace_rays call over this rank's slice (device buffers); e2e = the same call with host buffers
    (H2D of the rays, D2H of the hits inside the timed region)."""
    torch = env.torch
    from dust_b200 import host, scenes
    from dust_b200.pipeline import Context, Scene, HIT_DTYPE, RAY_DTYPE
    world, rank = env.world, env.rank
    cfg = CONFIGS[4]
    per_rank = (n_total + world - 1) // world
    first = rank * per_rank
    n = max(0, min(per_rank, n_total - first))
    t0 = time.perf_counter()
    dom = host.SynthDomain(seed=5, fill=0.01, domain_leaves=1024)
    ctx = Context(env.local)
    dtree = ctx.upload_tree(dom.describe())
    ident = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], np.float32)
    scene = Scene(ctx, [(ident, dtree)])
    build_s = time.perf_counter() - t0
    ctx.use_torch_stream()
    # rays first .. first+n-1 of the seeded pcg batch, generated on the device in slices (same bits as the numpy generator)
    d_rays = torch.empty((n, 8), dtype=torch.float32, device="cuda")
    for a in range(0, n, 1 << 24):
        b = min(n, a + (1 << 24))
        d_rays[a:b] = scenes.ray_batch_pcg(b - a, dom.extent, seed=RAY_SEED, first=first + a, device="cuda")
    d_hits = torch.empty((n, 4), dtype=torch.int32, device="cuda")
    scene.trace_rays_device(d_rays, d_hits, n, stats=True)
    torch.cuda.synchronize()
    st = ctx.stats()
    alg, alg_strict = algorithmic_bytes(st, 0, rays_in=n, hits_out=n), algorithmic_bytes(st, 0, rays_in=n, hits_out=n, strict=True)
    for _ in range(max(3, warmup)):
        scene.trace_rays_device(d_rays, d_hits, n)
    torch.cuda.synchronize()
    launches0 = ctx.stats()["kernel_launches"]
    sampler = ClockSampler(env.local) if clocks else None
    if sampler:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    env.barrier()
    torch.cuda.synchronize()
    for k in range(steps):
        env.flush.fill_(k & 0xFF)
        env.barrier()
        ev[k][0].record()
        scene.trace_rays_device(d_rays, d_hits, n)
        ev[k][1].record()
    torch.cuda.synchronize()
    launches = ctx.stats()["kernel_launches"] - launches0
    if sampler:
        t_load = time.perf_counter()
        while time.perf_counter() - t_load < 0.8:                   # enough nvidia-smi samples under the same load
            scene.trace_rays_device(d_rays, d_hits, n)
            torch.cuda.synchronize()
    env.barrier()
    clk = sampler.stop() if sampler else None
    ms = sum(a.elapsed_time(b) for a, b in ev) / steps
    ms_max = env.allmax([ms])[0]
    # ---- e2e through the host-buffer call: this rank's slice from / to pinned host memory
    e2e_rec = None
    hits_dev = d_hits.cpu().numpy().view(np.uint32)
    if e2e:
        ctx.set_stream(None)
        rays_pin = torch.empty((n, 8), dtype=torch.float32).pin_memory()          # pinned host memory: the copies of
        rays_pin.copy_(d_rays)                                                    # consecutive chunks overlap the kernels
        hits_pin = torch.empty(n * 16, dtype=torch.uint8).pin_memory()
        rays_host = rays_pin.numpy().reshape(-1).view(RAY_DTYPE)
        hits_host = hits_pin.numpy().view(HIT_DTYPE)
        scene.trace_rays(rays_host, out=hits_host)
        env.barrier()
        t0 = time.perf_counter()
        e2e_steps = max(1, min(steps, 3))
        for _ in range(e2e_steps):
            scene.trace_rays(rays_host, out=hits_host)
            int(hits_host["instance"][0])
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        e2e_s_max = env.allmax([e2e_s])[0]
        e2e_same = bool(np.array_equal(hits_host.view(np.uint32).reshape(-1, 4), hits_dev))
        e2e_rec = {"value": n_total / e2e_s_max / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": n * 32, "d2h_bytes_per_step": n * 16,
                   "api": "dust_trace_rays(DUST_MEM_HOST), %d rays per rank and call from / to pinned host memory, 2 Mi-ray chunks on two streams" % n,
                   "hits_equal_device_path": e2e_same}
        ctx.use_torch_stream()
    env.barrier()
    if rank != 0:
        return None
    peak, peak_src = measured_peak()
    achieved = alg / (ms_max * 1e-3) / 1e9
    hit_frac = float((d_hits[:, 1] != -1).float().mean().item())
    traffic, traffic_note = traffic_record("config4_dram_bytes_per_1e8_rays")
    if traffic is not None:
        traffic = traffic * (n / 1e8)
    # ---- CPU baseline beside it: the oracle port on all host threads over the first rays of the same batch,
    # and (it is the checker, after all) a bit-for-bit comparison of those rays with the GPU's answers
    cpu_rec, parity = None, None
    if cpu:
        from oracle import pyoracle as po
        desc = dom.describe()
        otree = po.TreeView(desc)                                   # a view of the same host tree (kept alive with the scene)
        osc = po.Scene([(ident, otree._h, dom.materials, dom.palette)])
        n_cal = min(n, 100_000)
        cal = scenes.ray_batch_pcg(n_cal, dom.extent, seed=RAY_SEED, first=first)
        t0 = time.perf_counter()
        osc.trace_rays(cal)
        rate = n_cal / max(time.perf_counter() - t0, 1e-6)
        n_cpu = int(max(n_cal, min(n, 10.0 * rate)))
        host_rays = scenes.ray_batch_pcg(n_cpu, dom.extent, seed=RAY_SEED, first=first)
        gen_same = bool(np.array_equal(host_rays.view(np.uint32).reshape(-1, 8), d_rays[:n_cpu].cpu().numpy().view(np.uint32)))
        t0 = time.perf_counter()
        o_hits, _ = osc.trace_rays(host_rays)
        dt = time.perf_counter() - t0
        same = bool(np.array_equal(o_hits.view(np.uint32).reshape(-1, 4), hits_dev[:n_cpu]))
        cpu_rec = {"value": n_cpu / dt / 1e6, "unit": "Mrays/s", "cores": os.cpu_count() or 1, "kind": "port",
                   "sample": f"first {n_cpu} rays of the same batch, {dt:.1f} s"}
        parity = {"rays": n_cpu, "hits_bit_exact": same, "ok": bool(same and gen_same), "host_and_device_ray_generators_agree": gen_same,
                  "against": "oracle/ on the first rays of the batch (instance, leaf, voxel, t compared as bits)"}
    return {
        "metric": "Mrays/s, random-origin ray batch on the synthetic 4096^3 VDB", "value": n_total / (ms_max * 1e-3) / 1e6, "unit": "Mrays/s",
        "n_gpus": world, "steps": steps, "warmup": max(3, warmup), "ms_per_step": ms_max, "scaling": "strong",
        "config": {"workload": f"BASELINE.json configs[{cfg_id}]: {cfg['name']}", "rays": n_total, "rays_per_gpu": n, "leaves": dom.num_leaves,
                   "tree_device_bytes": dtree.device_bytes, "hit_fraction": hit_frac,
                   "rays_from": "pcg(pcg(i + pcg(seed)) + k) streams (headers/spatial_hash.glsl:105-111), seed %d; rank r traces rays [r n/N, (r+1) n/N)" % RAY_SEED,
                   "l2": "flushed between timed steps (256 MiB write); the tree (%.0f MB) and the rays exceed L2" % (dtree.device_bytes / 1e6),
                   "build_s": build_s},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "traffic_source": traffic_note,
                     "strict": {"algorithmic_bytes_per_launch": alg_strict, "achieved": alg_strict / (ms_max * 1e-3) / 1e9,
                                "frac": alg_strict / (ms_max * 1e-3) / 1e9 / peak},
                     "kernel": "wf_trace_pt<4> (persistent lanes, %d launch(es))" % (launches // max(1, steps)), "kernel_ms": ms_max,
                     "algorithmic_bytes_per_launch": alg, "bytes_per_ray": alg / max(n, 1), "peak_source": peak_src,
                     "note": "ncu (profiles/r02_config4_*): the walk of the 4096^3 domain mostly stays in the L1-resident upper levels; the kernel "
                             "is bound by the ALU pipe, not by DRAM latency"},
        "cpu_baseline": cpu_rec, "parity": parity, "e2e": e2e_rec, "gpu_launches": int(launches), "clocks": clk,
    }


def gi_record(env, S, steps, warmup, *, cpu=True):
    """The reference's own GI at its own sizes: 1920x1080, 1 sample per pixel per frame, radiance cache of
    32 Mi entries, 720 x 480 surfels (standard.rs:334-358).  A step = one dust_trace_gi frame (primary + ambient-occlusion +
    final-gather + surfel passes + the deterministic insert) on a warm cache, device-resident illuminance image.  N = 1 only: the
    cache is one sequentially-consistent table; sharding it is out of scope."""
    torch = env.torch
    from dust_b200 import host
    from dust_b200.pipeline import RadianceCache
    ctx, scene, sky = S["ctx"], S["scene"], S["sky"]
    W, H = 1920, 1080
    CAP, POOL = 32 * 1024 * 1024, 720 * 480
    cam = host.camera_look_at(width=W, height=H, **S["camkw"])
    cache = RadianceCache(ctx, CAP, POOL)
    ctx.use_torch_stream()
    img = torch.zeros((H, W, 4), dtype=torch.float16, device="cuda")
    # frame 0 on the fresh cache through the host call: statistics, and the frame the parity check below compares
    ctx.set_stream(None)
    first = cache.gi_frame(scene, cam, sky, W, H, seed=SEED, frame_index=0)
    t0_table, t0_pool = cache.download() if cpu else (None, None)
    ctx.use_torch_stream()
    frame_index = 1
    for _ in range(max(3, warmup) + 4):                              # a few frames until the pool is full and the cache answers
        cache.gi_frame_device(scene, cam, sky, W, H, img, seed=SEED, frame_index=frame_index); frame_index += 1
    torch.cuda.synchronize()
    launches0 = ctx.stats()["kernel_launches"]
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for k in range(steps):
        env.flush.fill_(k & 0xFF)
        ev[k][0].record()
        cache.gi_frame_device(scene, cam, sky, W, H, img, seed=SEED, frame_index=frame_index); frame_index += 1
        ev[k][1].record()
    torch.cuda.synchronize()
    launches = ctx.stats()["kernel_launches"] - launches0
    ms = sum(a.elapsed_time(b) for a, b in ev) / steps
    ctx.set_stream(None)
    last = cache.gi_frame(scene, cam, sky, W, H, seed=SEED, frame_index=frame_index)          # statistics of a warm frame
    st = last["stats"]
    hit_px = int(np.isfinite(last["gbuffer"]["depth"]).sum())
    closest = W * H + hit_px + st["far_hits"] + st["far_misses"] + st["surfel_rays"]
    rec = {"metric": "Mrays/s, one frame of the reference's GI (primary + near + far + surfel rays) on castle.vox", "value": closest / (ms * 1e-3) / 1e6,
           "unit": "Mrays/s", "ms_per_step": ms, "frames_per_s": 1e3 / ms, "steps": steps, "n_gpus": 1,
           "config": {"workload": "SURVEY 8f rank 2: StandardPipeline::render's four dispatches, 1920x1080, cache 32 Mi entries, 345 600 surfels",
                      "scene": S["label"], "rays_counted": "closest-hit rays: primary + ambient-occlusion + final-gather + surfel (sun shadow rays traced, not counted)"},
           "rays_per_step": closest, "warm_frame_stats": st, "gpu_launches": int(launches),
           "cache": {"entries_used_after_frame_0": None, "table_bytes": (CAP + 2) * 12, "pool_bytes": POOL * 16}}
    if cpu:
        from oracle import pyoracle as po
        osc = oracle_scene(S["data"])
        ost = po.GiState(CAP, POOL)
        t0 = time.perf_counter()
        o = osc.gi_frame(cam, sky, W, H, ost, seed=SEED, frame_index=0)
        dt = time.perf_counter() - t0
        os_ = o["stats"]
        o_hit = int(np.isfinite(o["gbuffer"]["depth"]).sum())
        o_closest = W * H + o_hit + os_["far_hits"] + os_["far_misses"] + os_["surfel_rays"]
        same_img = bool(np.array_equal(first["illuminance"].view(np.uint16), o["illuminance"].view(np.uint16)))
        same_table = bool(np.array_equal(t0_table.view(np.uint8), ost.table.view(np.uint8)))
        same_pool = bool(np.array_equal(t0_pool.view(np.uint8), ost.pool.view(np.uint8)))
        gs = dict(first["stats"]); gs.pop("insert_segments")
        os2 = dict(os_); os2.pop("insert_segments")
        rec["cache"]["entries_used_after_frame_0"] = int((t0_table["fingerprint"] != 0).sum())
        rec["cpu_baseline"] = {"value": o_closest / dt / 1e6, "unit": "Mrays/s", "cores": os.cpu_count() or 1, "kind": "port",
                               "sample": f"frame 0 of the same sequence on a fresh cache, {dt:.1f} s"}
        rec["parity"] = {"frame": 0, "illuminance_bit_exact": same_img, "cache_table_bit_exact": same_table, "surfel_pool_bit_exact": same_pool,
                         "stats_equal": gs == os2, "ok": bool(same_img and same_table and same_pool and gs == os2),
                         "against": "oracle/oracle_gi.c (all 2 073 600 pixels, all 33 554 434 cache entries, all 345 600 surfels compared as bytes)"}
    cache.close()
    ctx.use_torch_stream()
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 4, 5], help="index into BASELINE.json configs (5 = alias of 4)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="castle configs at N > 1: what `value` measures (the other one is reported beside it)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="configs[2] only: skip the extra configs[3] / configs[4] records")
    ap.add_argument("--rays", type=int, default=0, help="--config 4: rays of the whole batch (default 1e8, BASELINE.json)")
    args = ap.parse_args()
    cfg_id = 4 if args.config == 5 else args.config
    cfg = CONFIGS[cfg_id]
    if args.impl == "reference":
        (reference_arm_rays if cfg_id == 4 else reference_arm)(args, cfg, cfg_id)
        return
    env = Env(args)
    cpu = not args.no_cpu_baseline
    base = {"metric": METRIC, "higher_is_better": True, "vs_baseline": None, "dtype": "f32", "data": "synthetic"}
    if cfg_id == 4:
        rec = rays_record(env, int(args.rays or cfg["rays"]), args.steps, args.warmup, cpu=cpu)
        if env.rank == 0:
            print(json.dumps({**base, **rec}), flush=True)
        env.close()
        return
    S = load_castle(env)
    rec = castle_record(env, S, cfg, cfg_id, args.scaling if env.world > 1 else "weak", args.steps, args.warmup, cpu=cpu)
    other = None
    if env.world > 1:
        # the other scaling mode of the same config beside it (shorter: no per-kernel pass, no clocks of its own)
        o_mode = "strong" if args.scaling == "weak" else "weak"
        other = castle_record(env, S, cfg, cfg_id, o_mode, args.steps, args.warmup, kernels=False, cpu=False, clocks=False)
    extra = {}
    if cfg_id == 2 and not args.no_extras:
        k = max(2, min(args.steps, 5))
        jobs = [("config3", lambda: castle_record(env, S, CONFIGS[3], 3, "strong", k, 3, kernels=False, cpu=cpu, clocks=True, max_cpu_s=8.0)),
                ("config4", lambda: rays_record(env, int(args.rays or CONFIGS[4]["rays"]), k, 3, cpu=cpu))]
        if env.world == 1:
            jobs.append(("gi", lambda: gi_record(env, S, k, 3, cpu=cpu)))
        for name, fn in jobs:
            try:
                r = fn()
            except Exception as e:                                  # an extra must not take the headline down with it
                r = {"error": f"{type(e).__name__}: {e}"}
                if env.world > 1:
                    raise
            extra[name] = r
    if env.rank == 0:
        line = {**base, **rec,
                "scene_stats": {"voxels": S["loaded"].num_voxels, "leaves": S["loaded"].num_leaves, "instances": len(S["loaded"].instances),
                                "device_tree_bytes": S["loaded"].device_bytes, "upload_s": S["upload_s"], "device_build_s": S["device_build_s"]}}
        if other is not None:
            line[other["scaling"]] = {k: other[k] for k in ("value", "unit", "ms_per_step", "rays_per_step", "scaling", "sharded_equals_whole", "step_parts_ms", "e2e")
                                      if k in other}
        if extra:
            line["extra"] = extra
        print(json.dumps(line), flush=True)
    env.close()


if __name__ == "__main__":
    main()
