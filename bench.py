#!/usr/bin/env python3
"""bench.py -- throughput of the batched draw path on B200, with roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = one pass of the hot path (column-binning kernel + tile draw kernel) over one batch of viewpoints whose
draw lists were produced beforehand by the host front-end (doom_rust_renderer_b200/csrc/host/drr_scene.cpp).

  value  : whole-job Mpixels/s (screen pixels W*H*frames / time) with draw lists resident in HBM, CUDA events on the
           launching stream, max over ranks.
  e2e    : same metric through the C ABI with HOST draw lists: every step copies the lists from pinned host memory
           (H2D), draws, and reads the per-frame checksums back (D2H); frames stay in HBM (north_star: "at most a
           host-side gather of per-frame CRCs").
  N > 1  : viewpoint batches shard across GPUs (one process per GPU, no collective on the draw path, per-GPU work
           fixed => "weak"); torch.distributed/NCCL is used only for the barrier, the max-over-ranks and the checksum gather.

Default workload = BASELINE.json configs[1]: E1M1-class synthetic map, 4096 walk viewpoints, 320x200, walls+flats+sky.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (wad kind, W, H, views per GPU, phases, description)
    "walk320": ("e1m1", 320, 200, 4096, 3, "BASELINE configs[1]: E1M1-class walk, 4096 viewpoints, 320x200, walls+flats+sky"),
    "walk1280": ("e1m1", 1280, 800, 512, 7, "E1M1-class walk, 512 viewpoints, 1280x800, all phases (north_star target resolution)"),
    "walls1280": ("e1m1", 1280, 800, 256, 1, "BASELINE configs[2]: walls only, 1280x800"),
    "flats1280": ("e1m1", 1280, 800, 256, 2, "BASELINE configs[2]: flats+sky only, 1280x800"),
    "empty1280": ("e1m1", 1280, 800, 256, 0, "no ops at all, 1280x800: clears + write-out only (fixed cost of a tile)"),
    "empty320": ("e1m1", 320, 200, 4096, 0, "no ops at all, 320x200: clears + write-out only (fixed cost of a tile)"),
    "things640": ("e1m1", 640, 400, 1024, 7, "BASELINE configs[3]: things, masked mids, lighting, 640x400 (bounded viewpoint count)"),
    "stress1920": ("stress", 1920, 1200, 128, 7, "BASELINE configs[4] map at 1920x1200 (bounded viewpoint count)"),
}


def measured_peak():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json, copy bandwidth)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU.  Started before the warm-up (NVML start-up takes longer than a
    short timed region); result() reports the samples that fall inside [mark_start, mark_end] and, when the region was
    too short to catch three of them, the samples of the surrounding loaded period (warm-up included), saying which."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag, self.max_mhz = index, [], False, None
        self.t0 = self.t1 = None
        self.error = None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            while not self.stop_flag:
                t = time.perf_counter()
                mhz = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.samples.append((t, mhz, r))
                time.sleep(0.002)
        except Exception as e:  # pragma: no cover
            self.error = "%s: %s" % (type(e).__name__, e)

    def mark_start(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def result(self):
        self.stop_flag = True
        self.join(timeout=2)
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap", 0x80: "hw_power_brake"}
        inside = [s for s in self.samples if self.t0 is not None and self.t0 <= s[0] <= self.t1]
        window = "timed region"
        if len(inside) < 3:
            inside = [s for s in self.samples if self.t0 is not None and self.t0 - 0.25 <= s[0] <= self.t1 + 0.01]
            window = "timed region + preceding warm-up (region shorter than 3 sampling periods)"
        mhz = sorted(s[1] for s in inside)
        reasons = sorted({n for s in inside for bit, n in names.items() if s[2] & bit})
        out = {"sm_mhz": mhz[len(mhz) // 2] if mhz else None, "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(inside), "window": window}
        if self.error:
            out["sampler_error"] = self.error
        return out


def make_wad(kind: str):
    from doom_rust_renderer_b200 import synth_wad
    data, gm, stats = synth_wad.build_wad(kind)
    d = tempfile.mkdtemp(prefix="drr_bench_")
    path = os.path.join(d, "synth_%s.wad" % kind)
    with open(path, "wb") as f:
        f.write(data)
    return path, gm


def viewpoints(gm, kind: str, n_total: int) -> np.ndarray:
    from doom_rust_renderer_b200 import synth_wad
    return synth_wad.walk_viewpoints(gm, n_total) if kind == "e1m1" else synth_wad.scatter_viewpoints(gm, n_total)


def record_batch(drr, ctx, scene, views, phases):
    """Run the host front-end for every viewpoint (worker threads, one recorder each).  A viewpoint on which the reference
    would panic (a seg passing exactly through the eye) is nudged by 1/8 map unit until it renders; returns the viewpoints
    actually used."""
    used = np.array(views, np.float32)
    for k in scene.emit_views(ctx, used, 0.0, phases):
        for attempt in range(16):
            used[k, 0] += np.float32(0.125)
            try:
                scene.emit_view(ctx, k, float(used[k, 0]), float(used[k, 1]), float(used[k, 2]), 0.0, phases)
                break
            except drr.DrrError as e:
                if e.code != -7:
                    raise
        else:
            raise RuntimeError("viewpoint %d cannot be rendered" % k)
    return used


def run_workload(name, args, rank, world, local_rank, dist, torch):
    from doom_rust_renderer_b200 import lib as drr
    kind, W, H, n_views, phases, desc = WORKLOADS[name]
    if args.views:
        n_views = args.views
    path, gm = make_wad(kind)
    from doom_rust_renderer_b200 import shard
    all_views = viewpoints(gm, kind, n_views * world)
    lo, hi = shard.shard_range(len(all_views), rank, world)  # contiguous viewpoint range of this GPU
    mine = all_views[lo:hi]

    ctx = drr.Context(W, H, local_rank, n_views)
    scene = drr.Scene(path, "E1M1", W, H)
    scene.upload_assets(ctx)
    t0 = time.perf_counter()
    used = record_batch(drr, ctx, scene, mine, phases)
    host_build_s = time.perf_counter() - t0
    st0 = ctx.stats()

    stream = torch.cuda.Stream(device=local_rank)
    ctx.set_stream(stream.cuda_stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident lists: `value` ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    ctx.upload_lists()
    for _ in range(args.warmup):
        ctx.draw()
    ctx.sync()
    t_wait = time.perf_counter()
    while not sampler.samples and sampler.error is None and time.perf_counter() - t_wait < 3.0:
        ctx.draw()  # keep the GPU loaded until NVML delivers its first sample (untimed)
        ctx.sync()
    barrier()
    launches0 = ctx.stats()["kernel_launches"]
    ctx.profile_begin(args.steps)
    sampler.mark_start()
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(args.steps):
            ctx.draw()
        e1.record(stream)
    e1.synchronize()
    sampler.mark_end()
    barrier()
    clocks = sampler.result()
    prof_steps, setup_ms_tot, march_ms_tot = ctx.profile_end()
    launches = ctx.stats()["kernel_launches"] - launches0
    ms_step = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    crc_dev = ctx.read_checksums(0, n_views)

    # ---- host lists every step: `e2e` ----
    for _ in range(max(1, args.warmup // 2)):
        ctx.submit()
        ctx.read_checksums(0, n_views)
    barrier()
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(args.steps):
            ctx.submit()  # pinned host lists -> H2D (chunked, overlapped with the kernels) -> bin + draw
            crc_e2e = ctx.read_checksums(0, n_views)
        e1.record(stream)
    e1.synchronize()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    assert (crc_e2e == crc_dev).all(), "checksums changed between device-resident and end-to-end passes"

    # host-side gather of the per-frame checksums (8 B/frame), off the timed path: checksum of checksums
    coc = shard.checksum_of_checksums(shard.gather_checksums(crc_dev, device="cuda") if world > 1 else crc_dev)

    st = ctx.stats()
    kernel_name = ctx.kernel_name()
    frames_total = n_views * world
    px_total = W * H * frames_total
    alg_bytes_launch = 3 * W * H * n_views + st["drawlist_bytes_algorithmic"]  # per GPU per launch
    march_ms = march_ms_tot / max(prof_steps, 1)
    setup_ms = setup_ms_tot / max(prof_steps, 1)
    peak, peak_src = measured_peak()
    achieved = alg_bytes_launch / (march_ms * 1e-3) / 1e9
    traffic = None  # DRAM bytes of one tile-kernel launch: per-frame figure from the committed ncu capture x frames of this launch
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
        if name in tj and args.views == 0:
            traffic = tj[name]["dram_bytes_per_frame"] * n_views
    except Exception:
        pass
    res = {
        "workload": name, "desc": desc, "W": W, "H": H, "views_per_gpu": n_views, "phases": phases,
        "value": px_total / (ms_step * 1e-3) / 1e6, "frames_per_s": frames_total / (ms_step * 1e-3), "ms_per_step": ms_step,
        "e2e_value": px_total / (ms_e2e * 1e-3) / 1e6, "e2e_ms_per_step": ms_e2e, "e2e_frames_per_s": frames_total / (ms_e2e * 1e-3),
        "h2d_bytes_per_step": st["device_list_bytes"], "d2h_bytes_per_step": 8 * n_views,
        "gpu_launches": launches, "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "frac_of_nominal_8000": achieved / 8000.0, "peak_source": peak_src, "traffic": traffic,
                     "traffic_source": "profiles/r1_traffic.json (ncu dram__bytes_read.sum + dram__bytes_write.sum per frame x frames)" if traffic else None,
                     "algorithmic_bytes_per_launch": alg_bytes_launch, "framebuffer_bytes_per_launch": 3 * W * H * n_views,
                     "drawlist_bytes_per_launch": st["drawlist_bytes_algorithmic"], "kernel_ms": march_ms, "setup_ms": setup_ms,
                     "kernel_share_of_step": march_ms / (march_ms + setup_ms) if march_ms + setup_ms > 0 else None},
        "lists": {k: st[k] for k in ("seg_headers", "column_records", "visplanes", "visplane_columns", "spans", "device_list_bytes")},
        "host_build_s": host_build_s, "checksum_of_checksums": "%016x" % coc,
    }
    # ---- viewpoints in, checksums out: the front-end on the GPU too (every phase, map objects included) ----
    if True:
        ctx.reset()
        assert scene.emit_views_device(ctx, used, 0.0, phases) == []
        ctx.draw()
        crc_fe = ctx.read_checksums(0, n_views)
        fe_ok = bool((crc_fe == crc_dev).all())
        st_fe = ctx.stats()
        barrier()
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(args.steps):
                ctx.reset()
                scene.emit_views_device(ctx, used, 0.0, phases)  # 12 B per viewpoint up; count pass, offsets on the host, emit pass
                ctx.draw()
                crc_fe = ctx.read_checksums(0, n_views)
            e1.record(stream)
        e1.synchronize()
        count_ms, emit_ms = ctx.fe_last_times()
        barrier()
        ms_fe = max_over_ranks(e0.elapsed_time(e1) / args.steps)
        res["device_front_end"] = {
            "value": px_total / (ms_fe * 1e-3) / 1e6, "unit": "Mpixel/s", "frames_per_s": frames_total / (ms_fe * 1e-3), "ms_per_step": ms_fe,
            "mode": "single pass: per-view slabs, then compaction" if ctx.fe_last_mode() == 1 else "two passes: count, then emit",
            "front_end_kernel_ms": emit_ms, "compaction_or_count_ms": count_ms, "h2d_bytes_per_step": 28 * n_views, "d2h_bytes_per_step": (40 + 8) * n_views,
            "what": "viewpoints -> drr_frontend_kernel -> drr_fe_compact_kernel -> bin -> tile -> checksums; no draw list crosses PCIe",
            "checksums_equal_host_front_end": fe_ok and bool((crc_fe == crc_dev).all()),
            "lists_equal_host_front_end": all(st_fe[k] == st[k] for k in ("seg_headers", "column_records", "visplanes", "visplane_columns", "spans")),
            "host_front_end_s": host_build_s}
    ctx.close()
    scene.close()
    return res, (path, used, crc_dev, W, H, phases)


# ---- CPU baseline (the oracle; the Rust reference cannot be built in this image: no rustc/cargo/SDL2) -------------------
def _cpu_worker(job):
    path, W, H, phases, views = job
    from oracle import orc
    g = orc.Game(path, "E1M1", W, H)
    out = np.empty((H, W, 3), np.uint8)
    from doom_rust_renderer_b200.lib import checksum_numpy
    sums = []
    t0 = time.perf_counter()
    for v in views:
        g.render(float(v[0]), float(v[1]), float(v[2]), 0.0, phases, out=out)
        sums.append(checksum_numpy(out))
    return time.perf_counter() - t0, sums


def cpu_render(path, W, H, phases, views, procs):
    """Render `views` with `procs` processes (one oracle instance each, disjoint slices).  Returns (wall seconds, checksums)."""
    import multiprocessing as mp
    if procs == 1:
        t, sums = _cpu_worker((path, W, H, phases, views))
        return t, sums
    chunks = [views[i::procs] for i in range(procs)]
    with mp.get_context("fork").Pool(procs) as pool:
        pool.map(_cpu_worker, [(path, W, H, phases, c[:1]) for c in chunks])  # load WAD / warm up outside the timing
        t0 = time.perf_counter()
        outs = pool.map(_cpu_worker, [(path, W, H, phases, c) for c in chunks])
        wall = time.perf_counter() - t0
    sums = [None] * len(views)
    for i, (_, s) in enumerate(outs):
        sums[i::procs] = s
    return wall, sums


def cpu_baseline(path, used, crc_dev, W, H, phases):
    import __graft_entry__ as g
    g.build(quiet=True)
    cores = os.cpu_count() or 1
    per_frame_guess = 8e-3 * (W * H) / 64000.0
    n1 = int(max(8, min(len(used), 6.0 / per_frame_guess)))
    t1, sums1 = cpu_render(path, W, H, phases, used[:n1], 1)
    nm = int(max(cores, min(len(used), cores * 8.0 / per_frame_guess)))
    tm, sumsm = cpu_render(path, W, H, phases, used[:nm], cores)
    ok = all(int(crc_dev[i]) == s for i, s in enumerate(sumsm)) and all(int(crc_dev[i]) == s for i, s in enumerate(sums1))
    return {"value": W * H * nm / tm / 1e6, "unit": "Mpixel/s", "cores": cores, "kind": "port",
            "sample": "%d frames of the same batch, one oracle process per core (%d); oracle = literal C++ restatement, g++ -O2 -ffp-contract=off" % (nm, cores),
            "frames_per_s": nm / tm, "single_thread_value": W * H * n1 / t1 / 1e6, "single_thread_frames_per_s": n1 / t1,
            "single_thread_sample": "%d frames" % n1, "parity_checked_frames": nm, "parity_ok": bool(ok)}


def parity_sample(path, used, crc_dev, W, H, phases, n=16):
    """Oracle frames of a seeded sample of the batch against the device checksums (for the secondary workloads; the default
    workload is checked frame by frame in cpu_baseline)."""
    idx = np.sort(np.random.default_rng(0xD00D1993).choice(len(used), min(n, len(used)), replace=False))
    _, sums = cpu_render(path, W, H, phases, used[idx], min(os.cpu_count() or 1, len(idx)))
    return {"frames": [int(i) for i in idx], "ok": bool(all(int(crc_dev[i]) == s for i, s in zip(idx, sums)))}


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path (oracle port) on the box's host cores."""
    if rank != 0:
        return
    import __graft_entry__ as g
    g.build(quiet=True)
    kind, W, H, n_views, phases, desc = WORKLOADS[args.workload]
    path, gm = make_wad(kind)
    cores = os.cpu_count() or 1
    per_step = cores * 16
    views = viewpoints(gm, kind, n_views * max(1, args.gpus))
    from oracle import orc
    game = orc.Game(path, "E1M1", W, H)
    good = []
    for v in views:  # same rule as the GPU arm: skip what the reference would panic on
        try:
            game.render(float(v[0]), float(v[1]), float(v[2]), 0.0, phases)
            good.append(v)
        except orc.OracleError:
            pass
        if len(good) >= per_step:
            break
    good = np.array(good, np.float32)
    for _ in range(args.warmup):
        cpu_render(path, W, H, phases, good, cores)
    t = 0.0
    for _ in range(args.steps):
        dt, _ = cpu_render(path, W, H, phases, good, cores)
        t += dt
    ms = t / args.steps * 1e3
    val = W * H * len(good) / (ms * 1e-3) / 1e6
    sample = "%d frames per step, one process per core (%d cores)" % (len(good), cores)
    print(json.dumps({
        "impl": "reference", "metric": "textured Mpixels/s", "value": val, "unit": "Mpixel/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": args.workload, "desc": desc, "W": W, "H": H, "phases": phases, "frames_per_step": len(good)},
        "frames_per_s": len(good) / (ms * 1e-3),
        "cpu_baseline": {"value": val, "unit": "Mpixel/s", "cores": cores, "kind": "port", "sample": sample,
                         "note": "the Rust reference cannot be compiled here (no rustc/cargo/SDL2); this is the literal C++ restatement in oracle/"},
        "e2e": {"value": val, "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="walk320", choices=sorted(WORKLOADS))
    ap.add_argument("--views", type=int, default=0, help="override viewpoints per GPU")
    ap.add_argument("--secondary", default="walk1280,walls1280,flats1280,things640,stress1920", help="extra workloads reported under 'secondary' (N=1 only); '' = none")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank, world, local_rank = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the draw path has no CPU fallback (use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local_rank)
    saved_stdout = None
    if world > 1:
        # NCCL prints its version banner on the C-level stdout: keep stdout for the one JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if local_rank == 0:
        import __graft_entry__ as g
        g.build(quiet=True)
    if world > 1:
        dist.barrier()

    res, ctxinfo = run_workload(args.workload, args, rank, world, local_rank, dist, torch)
    out = {
        "metric": "textured Mpixels/s", "value": res["value"], "unit": "Mpixel/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "desc": res["desc"], "W": res["W"], "H": res["H"], "views_per_gpu": res["views_per_gpu"],
                   "phases": res["phases"], "sharding": "contiguous viewpoint ranges per GPU, no collective on the draw path",
                   "l2": "per-step working set (framebuffers %.0f MB + lists) exceeds the 126 MB L2; no explicit flush" %
                         (3 * res["W"] * res["H"] * res["views_per_gpu"] / 1e6)},
        "frames_per_s": res["frames_per_s"],
        "e2e": {"value": res["e2e_value"], "unit": "Mpixel/s", "h2d_bytes_per_step": res["h2d_bytes_per_step"],
                "d2h_bytes_per_step": res["d2h_bytes_per_step"], "ms_per_step": res["e2e_ms_per_step"], "frames_per_s": res["e2e_frames_per_s"],
                "result": "per-frame checksums (8 B/frame); framebuffers stay resident in HBM"},
        "gpu_launches": res["gpu_launches"], "clocks": res["clocks"], "roofline": res["roofline"], "lists": res["lists"],
        "host_build_s": res["host_build_s"], "checksum_of_checksums": res["checksum_of_checksums"],
        # SURVEY 8(d) "separately: end-to-end incl. host list build": one batch = the C++ front-end on all host cores
        # (drr_scene_emit_views) + one drr_submit + checksums back.  The front-end is the caller's side of the boundary (the
        # reference's own BSP walk), not the draw path; it is what a GPU-side front-end would remove.
        "with_front_end": {"value": res["W"] * res["H"] * res["views_per_gpu"] * world / (res["host_build_s"] + res["e2e_ms_per_step"] * 1e-3) / 1e6,
                           "unit": "Mpixel/s", "front_end_s": res["host_build_s"], "host_threads": os.cpu_count()},
    }
    if "device_front_end" in res:  # the same batch with the front-end on the GPU too: viewpoints in, checksums out
        out["with_front_end_device"] = res["device_front_end"]
    if world == 1 and rank == 0:
        if not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(*ctxinfo)
        sec = []
        for name in [s for s in args.secondary.split(",") if s]:
            a2 = argparse.Namespace(**vars(args))
            a2.views = 0
            a2.steps = max(3, args.steps // 2)
            r2, info2 = run_workload(name, a2, rank, world, local_rank, dist, torch)
            if not args.no_cpu_baseline and not name.startswith("empty"):
                r2["parity_sample"] = parity_sample(*info2)
            sec.append({k: r2[k] for k in ("workload", "desc", "W", "H", "views_per_gpu", "phases", "value", "frames_per_s", "ms_per_step",
                                            "e2e_value", "gpu_launches", "roofline", "lists", "clocks", "host_build_s", "device_front_end", "parity_sample") if k in r2})
        if sec:
            out["secondary"] = sec
    if saved_stdout is not None:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        os.dup2(2, 1)
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
