#!/usr/bin/env python3
"""bench.py -- throughput of the batched draw path on B200, with roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = `passes_per_step` passes of the hot path over one batch of viewpoints (the batch is small -- 4096 frames of
320x200 draw in under a millisecond -- so a step repeats it, the same number of times at every N, to make the timed window
long enough for an 8-process max-over-ranks; every figure is per pixel / per frame actually drawn).

  value  : whole-job Mpixels/s (screen pixels W*H*frames / time) with the draw lists resident in HBM: column-binning kernel +
           tile draw kernel per pass, CUDA events on the launching stream, max over ranks.
  e2e    : the same metric from VIEWPOINTS to per-frame checksums through the C ABI, i.e. what one Renderer::render() call of
           the reference covers (src/renderer/mod.rs:118-136: front-end + draw): every pass uploads the viewpoints (12 B each),
           runs the device front-end (drr_fe_emit_views), bins, draws and reads the checksums back (8 B per frame); the
           framebuffers stay in HBM (north_star: "at most a host-side gather of per-frame CRCs").  This is the pair of the
           CPU arm, which also walks the BSP.  Measured with TWO batches in flight (two contexts with their own stream and
           framebuffers, one host thread each: the second batch's kernels fill the GPU while the first one's counts are
           turned into offsets on the host and its checksums are read back); `e2e.one_batch_in_flight` is the strictly
           sequential figure.  (The stress map's 8192 frames of 1920x1200 are 57 GB: one batch in flight there.)
  e2e_host_lists : the draw path alone through its C-ABI boundary: the HOST front-end's recorded lists go up every pass
           (pinned H2D, chunked and overlapped), are drawn, and the checksums come back.
  N > 1  : viewpoints shard across GPUs by stride (rank r draws viewpoints r, r+N, ...: the same mix of the walk on every GPU),
           one process per GPU, no collective on the draw path, per-GPU work fixed => "weak"; torch.distributed/NCCL is used
           only for the barrier, the max-over-ranks and the gathers of checksums and per-rank times.

Default workload = BASELINE.json configs[1]; `secondary` carries 1280x800 (all phases = the north_star target, walls-only and
flats-only = configs[2]), configs[3] (640x400 with things, 4096 viewpoints per GPU) and configs[4] (stress map at 1920x1200,
8192 viewpoints per GPU = 65536 over 8 GPUs), at every N, each with a seeded oracle parity sample on rank 0.
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
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from doom_rust_renderer_b200.checksum import checksum_numpy  # noqa: E402  (pure numpy)
from doom_rust_renderer_b200.workloads import WORKLOADS, Content  # noqa: E402  (pure Python)

# what binds the tile kernel, from the committed ncu captures (profiles/r2_final_tile_*.txt): not HBM
LIMITER = {
    "what": "instruction issue (68-72 % of the slots busy) together with the L1/shared pipe (61-73 %), not HBM (DRAM 22 %): the per-pixel f32 "
            "arithmetic the reference prescribes; 38.8 issue slots per screen pixel at 1280x800, 53 at 320x200",
    "source": "profiles/r2b_final_tile_1280x800.txt, profiles/r2b_final_tile_320x200.txt (ncu --set full)",
}


def measured_peak():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json, copy bandwidth)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def load_json(name):
    try:
        return json.load(open(os.path.join(ROOT, "profiles", name)))
    except Exception:
        return {}


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU every 5 ms.  Started before the warm-up (NVML start-up takes longer
    than a short timed region); result() reports the samples that fall inside [mark_start, mark_end]."""

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
                time.sleep(0.005)
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


def make_config(name, n_views, passes, content_source):
    kind, W, H, _, phases, _, desc = WORKLOADS[name]
    return {"workload": name, "desc": desc, "W": W, "H": H, "views_per_gpu": n_views, "phases": phases, "passes_per_step": passes,
            "content": content_source, "sharding": "strided viewpoint assignment (rank r: viewpoints r, r+N, ...), no collective on the draw path",
            "l2": "per-pass working set (framebuffers %.0f MB + lists) exceeds the 126 MB L2; no explicit flush" % (3 * W * H * n_views / 1e6)}


# ---- the reference's CPU implementation of the path (oracle port; the Rust reference cannot be built here) -------------------
_WORKER = {}


def _pool_init(path, W, H):
    from oracle import orc
    _WORKER["game"] = orc.Game(path, "E1M1", W, H)  # WAD parse + asset decode once per worker, outside every timed region
    _WORKER["out"] = np.empty((H, W, 3), np.uint8)


def _pool_render(job):
    views, phases, want_sums = job
    g, out = _WORKER["game"], _WORKER["out"]
    t0 = time.perf_counter()
    for v in views:
        g.render(float(v[0]), float(v[1]), float(v[2]), 0.0, phases, out=out)
    dt = time.perf_counter() - t0
    sums = None
    if want_sums:  # second, untimed pass: the checksum is not part of what the reference does per frame
        sums = []
        for v in views:
            g.render(float(v[0]), float(v[1]), float(v[2]), 0.0, phases, out=out)
            sums.append(checksum_numpy(out))
    return dt, sums


def _pool_usable(job):
    """Which of `views` the reference renders without panicking (same rule as the GPU arm's nudging)."""
    from oracle import orc
    views, phases = job
    ok = []
    for v in views:
        try:
            _WORKER["game"].render(float(v[0]), float(v[1]), float(v[2]), 0.0, phases, out=_WORKER["out"])
            ok.append(True)
        except orc.OracleError:
            ok.append(False)
    return ok


class CpuPool:
    """`procs` worker processes, one oracle instance each, alive across steps."""

    def __init__(self, path, W, H, procs):
        import multiprocessing as mp
        self.procs = procs
        # "spawn": the parent may hold a CUDA context and NCCL threads, which a forked child must not inherit
        self.pool = mp.get_context("spawn").Pool(procs, initializer=_pool_init, initargs=(path, W, H))

    def render(self, views, phases, want_sums=False):
        """Render `views` (split evenly over the workers).  Returns (wall seconds, checksums or None)."""
        chunks = [views[i::self.procs] for i in range(self.procs)]
        t0 = time.perf_counter()
        outs = self.pool.map(_pool_render, [(c, phases, want_sums) for c in chunks], chunksize=1)
        wall = time.perf_counter() - t0
        sums = None
        if want_sums:
            sums = [None] * len(views)
            for i, (_, s) in enumerate(outs):
                sums[i::self.procs] = s
        return wall, sums

    def usable(self, views, phases):
        chunks = [views[i::self.procs] for i in range(self.procs)]
        outs = self.pool.map(_pool_usable, [(c, phases) for c in chunks], chunksize=1)
        ok = np.zeros(len(views), bool)
        for i, o in enumerate(outs):
            ok[i::self.procs] = o
        return ok

    def close(self):
        self.pool.close()
        self.pool.join()


FRAMES_PER_WORKER = 64  # CPU sample: frames per worker process per step


def build_oracle():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)


def cpu_baseline(name, path, used, crc_dev, W, H, phases):
    """The oracle on the GPU box's host cores, on a bounded sample of the same batch: one process per core and
    single-threaded.  The timing IS the reference arm (`bench.py --impl reference`, run here as a child process so that it
    meets the same conditions as when the driver runs it: no CUDA context, no other threads of this process in its way);
    every sampled frame's checksum is compared with the device's in a separate, untimed pass."""
    build_oracle()
    cores = os.cpu_count() or 1

    def arm(procs, steps):
        cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", name, "--steps", str(steps), "--warmup", "2", "--ref-procs", str(procs)]
        return json.loads(subprocess.run(cmd, capture_output=True, text=True, check=True).stdout.strip().splitlines()[-1])

    multi, single = arm(cores, 6), arm(1, 1)
    nm = int(max(cores, min(len(used), multi["frames_per_step"])))
    pool = CpuPool(path, W, H, cores)
    _, sums = pool.render(used[:nm], phases, want_sums=True)
    pool.close()
    ok = all(int(crc_dev[i]) == s for i, s in enumerate(sums))
    return {"value": multi["value"], "unit": "Mpixel/s", "cores": cores, "kind": "port",
            "sample": multi["cpu_baseline"]["sample"] + "; 6 steps after 2 warm-ups; oracle = literal C++ restatement, g++ -O2 -ffp-contract=off",
            "frames_per_s": multi["frames_per_s"], "single_thread_value": single["value"], "single_thread_frames_per_s": single["frames_per_s"],
            "single_thread_sample": "%d frames" % single["frames_per_step"], "parity_checked_frames": nm, "parity_ok": bool(ok)}


def parity_sample(path, used, crc_dev, W, H, phases, n=16):
    """Oracle frames of a seeded sample of the batch against the device checksums."""
    build_oracle()
    idx = np.sort(np.random.default_rng(0xD00D1993).choice(len(used), min(n, len(used)), replace=False))
    pool = CpuPool(path, W, H, min(os.cpu_count() or 1, len(idx)))
    _, sums = pool.render(used[idx], phases, want_sums=True)
    pool.close()
    return {"frames": [int(i) for i in idx], "ok": bool(all(int(crc_dev[i]) == s for i, s in zip(idx, sums)))}


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path (oracle port) on all the box's host cores.  Nothing of
    the product is loaded here: the oracle library is the only native code of this process tree."""
    if rank != 0:
        return
    build_oracle()
    kind, W, H, n_views, phases, passes, desc = WORKLOADS[args.workload]
    if args.views:
        n_views = args.views
    content = Content(kind)
    cores = args.ref_procs or os.cpu_count() or 1
    scale = 64000.0 / (W * H)
    per_step = cores * max(4, int(FRAMES_PER_WORKER * scale))
    pool = CpuPool(content.path, W, H, cores)
    views = content.viewpoints(n_views * max(1, args.gpus))
    views = np.array(views[:min(len(views), per_step * 2)], np.float32)
    good = views[pool.usable(views, phases)][:per_step]  # same rule as the GPU arm: skip what the reference would panic on
    for _ in range(args.warmup):
        pool.render(good, phases)
    t = 0.0
    for _ in range(args.steps):
        t += pool.render(good, phases)[0]
    pool.close()
    ms = t / args.steps * 1e3
    val = W * H * len(good) / (ms * 1e-3) / 1e6
    sample = "%d frames of the workload per step (%d per worker), one oracle process per core (%d cores), workers and their WAD/asset state alive across steps" % (
        len(good), len(good) // cores, cores)
    print(json.dumps({
        "impl": "reference", "metric": "textured Mpixels/s", "value": val, "unit": "Mpixel/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": make_config(args.workload, n_views, passes, content.source),
        "frames_per_s": len(good) / (ms * 1e-3), "frames_per_step": len(good),
        "cpu_baseline": {"value": val, "unit": "Mpixel/s", "cores": cores, "kind": "port", "sample": sample,
                         "note": "the Rust reference cannot be compiled here (no rustc/cargo/SDL2); this is the literal C++ restatement in oracle/"},
        "e2e": {"value": val, "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ---- the B200 arm ------------------------------------------------------------------------------------------------------------
def settle_views_device(scene, ctx, views, phases):
    """Emit the batch with the device front-end; a viewpoint on which the reference would panic (a seg passing exactly
    through the eye) is nudged by 1/8 map unit until the whole batch renders.  Returns the viewpoints actually used."""
    used = np.array(views, np.float32)
    for _ in range(16):
        ctx.reset()
        bad = scene.emit_views_device(ctx, used, 0.0, phases)
        if not bad:
            return used
        used[bad, 0] += np.float32(0.125)
    raise RuntimeError("some viewpoints cannot be rendered")


def settle_views_host(drr, scene, ctx, views, phases):
    """The same with the host front-end (worker threads, one recorder each)."""
    used = np.array(views, np.float32)
    for k in scene.emit_views(ctx, used, 0.0, phases):
        for _ in range(16):
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


def run_workload(name, args, rank, world, local_rank, dist, torch, headline):
    from doom_rust_renderer_b200 import lib as drr, shard
    kind, W, H, n_views, phases, passes, desc = WORKLOADS[name]
    if args.views and headline:
        n_views = args.views
    content = Content(kind)
    all_views = content.viewpoints(n_views * world)
    mine = all_views[shard.shard_indices(len(all_views), rank, world)]  # this GPU's viewpoints (strided)
    ctx = drr.Context(W, H, local_rank, n_views)
    scene = drr.Scene(content.path, "E1M1", W, H)
    scene.upload_assets(ctx)
    stream = torch.cuda.Stream(device=local_rank)
    ctx.set_stream(stream.cuda_stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = args.steps if headline else max(3, args.steps // 2)

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

    def all_ranks(v):
        if world == 1:
            return [float(v)]
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        parts = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        return [float(p.item()) for p in parts]

    def timed(body):
        """K steps of `passes` passes each, bracketed as the contract says; returns ms per PASS (max over ranks, own)."""
        barrier()
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(steps * passes):
                body()
            e1.record(stream)
        e1.synchronize()
        barrier()
        own = e0.elapsed_time(e1) / (steps * passes)
        return max_over_ranks(own), own

    # ---- viewpoints -> checksums, front-end on the GPU: `e2e`; also how the secondary workloads get their lists ----
    used = settle_views_device(scene, ctx, mine, phases)
    ctx.draw()
    crc_fe = ctx.read_checksums(0, n_views)
    st = ctx.stats()

    def pass_fe():
        ctx.reset()
        scene.emit_views_device(ctx, used, 0.0, phases)  # 12 B per viewpoint up; front-end kernel, counts to the host (offsets of the record slots), the draw kernels read the per-view slabs
        ctx.draw()
        ctx.read_checksums(0, n_views)

    # ---- device-resident lists: `value` ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(max(3, args.warmup)):
        ctx.draw()
    ctx.sync()
    t_wait = time.perf_counter()
    while not sampler.samples and sampler.error is None and time.perf_counter() - t_wait < 3.0:
        ctx.draw()  # keep the GPU loaded until NVML delivers its first sample (untimed)
        ctx.sync()
    launches0 = ctx.stats()["kernel_launches"]
    ctx.profile_begin(steps * passes)
    sampler.mark_start()
    ms_pass, ms_pass_own = timed(ctx.draw)
    sampler.mark_end()
    clocks = sampler.result()
    prof_n, bin_ms_tot, tile_ms_tot = ctx.profile_end()
    launches = ctx.stats()["kernel_launches"] - launches0
    crc_dev = ctx.read_checksums(0, n_views)
    assert (crc_dev == crc_fe).all(), "checksums changed between passes"
    tile_ms, bin_ms = tile_ms_tot / max(prof_n, 1), bin_ms_tot / max(prof_n, 1)

    for _ in range(2):
        pass_fe()
    ms_fe, _ = timed(pass_fe)
    fe_front_ms, fe_kernel_ms = ctx.fe_last_times()
    fe_mode = ctx.fe_last_mode()
    assert (ctx.read_checksums(0, n_views) == crc_dev).all(), "end-to-end pass drew different frames"

    # ---- the same with TWO batches in flight: a second context (own stream, own framebuffers) driven by a second host thread.
    # One batch alone leaves the GPU idle while the host turns the front-end's counts into offsets, launches and reads back,
    # and the front-end kernel's last wave runs at a fraction of the machine; a second batch fills both.  Every pass still
    # uploads its viewpoints, runs the front-end, bins, draws and reads its checksums back.  (Not when two sets of framebuffers
    # would take more than 60 GB: the stress map's 8192 x 1920x1200 frames -- there front-end and draw are both throughput-bound, and two
    # HALF batches in flight measured slower than one whole batch at a time: 117.2 against 108.0 ms.)
    ms_fe2 = None
    if 2 * 3 * W * H * n_views <= 60e9:
        import threading
        ctx2 = drr.Context(W, H, local_rank, n_views)
        scene2 = drr.Scene(content.path, "E1M1", W, H)
        scene2.upload_assets(ctx2)
        stream2 = torch.cuda.Stream(device=local_rank)
        ctx2.set_stream(stream2.cuda_stream)

        def pass_fe2():
            ctx2.reset()
            scene2.emit_views_device(ctx2, used, 0.0, phases)
            ctx2.draw()
            return ctx2.read_checksums(0, n_views)

        assert (pass_fe2() == crc_dev).all(), "second context drew different frames"
        half = max(1, steps * passes // 2)

        def pair():
            errors = []

            def loop(body):
                try:
                    for _ in range(half):
                        body()
                except BaseException as ex:  # noqa: BLE001 -- re-raised below, on the main thread
                    errors.append(ex)

            ta, tb = threading.Thread(target=loop, args=(pass_fe,)), threading.Thread(target=loop, args=(pass_fe2,))
            ta.start(); tb.start(); ta.join(); tb.join()
            if errors:
                raise errors[0]

        pair()  # warm-up
        barrier()
        e0.record(stream)
        pair()
        done2 = torch.cuda.Event()
        done2.record(stream2)
        stream.wait_event(done2)
        e1.record(stream)
        e1.synchronize()
        barrier()
        ms_fe2 = max_over_ranks(e0.elapsed_time(e1) / (2 * half))
        assert (ctx.read_checksums(0, n_views) == crc_dev).all() and (ctx2.read_checksums(0, n_views) == crc_dev).all(), "pipelined passes drew different frames"
        ctx2.close()
        scene2.close()

    frames_total, px_total = n_views * world, W * H * n_views * world
    peak, peak_src = measured_peak()
    alg_bytes = 3 * W * H * n_views + st["drawlist_bytes_algorithmic"]  # per GPU per launch
    achieved = alg_bytes / (tile_ms * 1e-3) / 1e9
    tj = load_json("r2_traffic.json").get(name)
    traffic = tj["dram_bytes_per_launch"] if tj and tj.get("views") == n_views else None
    res = {
        "config": make_config(name, n_views, passes, content.source), "steps": steps,
        "value": px_total / (ms_pass * 1e-3) / 1e6, "frames_per_s": frames_total / (ms_pass * 1e-3), "ms_per_pass": ms_pass,
        "ms_per_step": ms_pass * passes, "gpu_launches": launches, "clocks": clocks,
        "e2e": {"value": px_total / ((ms_fe2 or ms_fe) * 1e-3) / 1e6, "unit": "Mpixel/s", "h2d_bytes_per_step": 28 * n_views * passes,
                "d2h_bytes_per_step": (40 + 8) * n_views * passes, "ms_per_step": (ms_fe2 or ms_fe) * passes, "ms_per_pass": ms_fe2 or ms_fe,
                "frames_per_s": frames_total / ((ms_fe2 or ms_fe) * 1e-3),
                "batches_in_flight": 2 if ms_fe2 else 1,
                "path": "viewpoints -> drr_fe_emit_views (front-end kernel into per-view slabs, counts to the host) -> bin (reads the slabs) -> tile -> per-frame checksums; no draw list crosses PCIe, framebuffers stay resident in HBM"
                        + ("; two batches in flight (two contexts, two host threads, one stream each)" if ms_fe2 else ""),
                "one_batch_in_flight": {"value": px_total / (ms_fe * 1e-3) / 1e6, "unit": "Mpixel/s", "ms_per_pass": ms_fe, "frames_per_s": frames_total / (ms_fe * 1e-3)},
                "front_end_kernel_ms": fe_kernel_ms, "compaction_or_count_ms": fe_front_ms,
                "mode": "single pass: per-view slabs, read in place by the draw kernels" if fe_mode == 1 else "two passes: count, then emit"},
        "roofline": {"bound": "issue", "roof": "hbm", "kernel": ctx.kernel_name(), "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "frac_of_nominal_8000": achieved / 8000.0, "peak_source": peak_src, "traffic": traffic,
                     "traffic_source": "profiles/r2_traffic.json: ncu dram__bytes_read.sum + dram__bytes_write.sum of one tile-kernel launch over this batch" if traffic else None,
                     "algorithmic_bytes_per_launch": alg_bytes, "framebuffer_bytes_per_launch": 3 * W * H * n_views,
                     "drawlist_bytes_per_launch": st["drawlist_bytes_algorithmic"], "kernel_ms": tile_ms, "bin_kernel_ms": bin_ms,
                     "kernel_share_of_pass": tile_ms / (tile_ms + bin_ms) if tile_ms + bin_ms > 0 else None, "limiter": LIMITER},
        "lists": {k: st[k] for k in ("seg_headers", "column_records", "visplanes", "visplane_columns", "spans", "device_list_bytes")},
        "per_rank": [{"rank": r, "ms_per_pass": a, "tile_kernel_ms": b, "bin_kernel_ms": c, "spans": int(d)} for r, (a, b, c, d) in
                     enumerate(zip(all_ranks(ms_pass_own), all_ranks(tile_ms), all_ranks(bin_ms), all_ranks(st["spans"])))],
    }

    # ---- the draw path alone through its boundary, host lists every pass: `e2e_host_lists` (headline workload only) ----
    if headline:
        ctx.reset()
        t0 = time.perf_counter()
        used_host = settle_views_host(drr, scene, ctx, mine, phases)
        host_build_s = time.perf_counter() - t0
        assert (used_host == used).all(), "host and device front-ends disagree on which viewpoints render"
        st_h = ctx.stats()

        def pass_host():
            ctx.submit()  # pinned host lists -> H2D (chunked, overlapped with the kernels) -> bin + draw
            ctx.read_checksums(0, n_views)

        for _ in range(2):
            pass_host()
        ms_h, _ = timed(pass_host)
        crc_h = ctx.read_checksums(0, n_views)
        res["e2e_host_lists"] = {
            "value": px_total / (ms_h * 1e-3) / 1e6, "unit": "Mpixel/s", "ms_per_pass": ms_h, "frames_per_s": frames_total / (ms_h * 1e-3),
            "h2d_bytes_per_pass": st_h["device_list_bytes"], "d2h_bytes_per_pass": 8 * n_views, "host_front_end_s": host_build_s, "host_threads": os.cpu_count(),
            "with_host_front_end": {"value": px_total / (host_build_s + ms_h * 1e-3) / 1e6, "unit": "Mpixel/s"},
            "path": "host front-end's recorded lists -> drr_submit (pinned H2D in chunks, overlapped) -> bin -> tile -> per-frame checksums",
            "checksums_equal_device_front_end": bool((crc_h == crc_dev).all()),
            "lists_equal_device_front_end": all(st_h[k] == st[k] for k in ("seg_headers", "column_records", "visplanes", "visplane_columns", "spans"))}

    # host-side gather of the per-frame checksums (8 B/frame), off the timed path: checksum of checksums
    allsums = shard.gather_checksums(crc_dev, device="cuda") if world > 1 else crc_dev
    res["checksum_of_checksums"] = "%016x" % shard.checksum_of_checksums(allsums)
    ctx.close()
    scene.close()

    return res, (content.path, used, crc_dev, W, H, phases)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="walk320", choices=sorted(WORKLOADS))
    ap.add_argument("--views", type=int, default=0, help="override viewpoints per GPU of the headline workload")
    ap.add_argument("--secondary", default="walk1280,walls1280,flats1280,things640,stress1920", help="extra workloads reported under 'secondary'; '' = none")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-procs", type=int, default=0, help="--impl reference: worker processes (0 = one per host core)")
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

    res, ctxinfo = run_workload(args.workload, args, rank, world, local_rank, dist, torch, headline=True)
    out = {
        "metric": "textured Mpixels/s", "value": res["value"], "unit": "Mpixel/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": res["config"], "frames_per_s": res["frames_per_s"], "ms_per_pass": res["ms_per_pass"], "e2e": res["e2e"],
        "e2e_host_lists": res.get("e2e_host_lists"), "gpu_launches": res["gpu_launches"], "clocks": res["clocks"], "roofline": res["roofline"],
        "lists": res["lists"], "per_rank": res["per_rank"], "checksum_of_checksums": res["checksum_of_checksums"],
    }
    if rank == 0 and not args.no_cpu_baseline:
        if world == 1:
            out["cpu_baseline"] = cpu_baseline(args.workload, *ctxinfo)
        else:
            out["parity_sample"] = parity_sample(*ctxinfo)
    sec = []
    for name in [s for s in args.secondary.split(",") if s and s != args.workload]:
        r2, info2 = run_workload(name, args, rank, world, local_rank, dist, torch, headline=False)
        if rank == 0 and not args.no_cpu_baseline and not name.startswith("empty"):
            r2["parity_sample"] = parity_sample(*info2)
        r2["workload"] = name
        sec.append(r2)
        if name == "walk1280":  # the north_star target resolution, where the driver keeps it
            rf = r2["roofline"]
            out["roofline"]["at_1280x800"] = {"workload": "walk1280", "achieved": rf["achieved"], "frac": rf["frac"], "kernel_ms": rf["kernel_ms"],
                                              "traffic": rf["traffic"], "algorithmic_bytes_per_launch": rf["algorithmic_bytes_per_launch"],
                                              "value": r2["value"], "e2e_value": r2["e2e"]["value"], "n_gpus": world}
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
