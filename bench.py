#!/usr/bin/env python
"""bench.py -- Mrays/s and ms/frame of the Hybrid pipeline hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg1|cfg3|cfg4|cfg5]

One "step" = one frame of the hot path (Renderer::rasterize_hybrid: G-buffer -> shadow rays -> denoise -> resolve +
reflection rays) on the default scene at 1920x1080 with a static camera, frame index advancing (temporal accumulation),
i.e. BASELINE.json configs[1].  Rays counted are the traceRayEXT calls the reference would issue (shadow + reflection +
secondary shadow); our own primary-visibility rays are reported separately and NOT counted.

N > 1 (torchrun, one rank per GPU): the frame is tile-sharded over a replicated BVH and the finished sRGB8 pixels are stored
straight into rank 0's image over NVLink peer memory -- the composite is the only exchange; "scaling": "strong".

Prints ONE JSON line (rank 0) with the contract keys plus "roofline", "cpu_baseline", "parity", "e2e", "clocks", "gpu_launches",
"frame_ms", "build", "cfg3".  `config` is the static description of the workload (identical in both arms); everything measured
lives outside it.  The CPU oracle (oracle/) is used only for the parity / cpu_baseline leg and for --impl reference; it is never
on the GPU path.  A parity result below north_star's bar makes the run fail (exit code 1) after the line is printed.
"""
import argparse
import glob
import json
import math
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (scene index, width, height, mode, spp, shadow samples, faithful_rng, description)
    "cfg2": (0, 1920, 1080, 2, 1, 1, 1, "default scene, Hybrid 1920x1080, 1 shadow ray/light/pixel + reflection rays, static camera"),
    "cfg1": (0, 1280, 720, 2, 1, 1, 1, "default scene, Hybrid 1280x720 (oracle-sized case)"),
    "cfg3": (0, 1920, 1080, 1, 64, 64, 0, "default scene, Raytraced 1920x1080, 64 spp x 64 shadow samples, advancing RNG"),
    "cfg4": (2, 3840, 2160, 2, 1, 1, 1, "instanced stress scene (~10M instanced triangles), Hybrid 3840x2160"),
    "cfg5": (3, 3840, 2160, 2, 1, 16, 0, "default scene, 16 ring lights x 16 shadow samples, Hybrid 3840x2160"),
}
METRIC = "Mrays/s (Hybrid 1080p hot path: shadow + reflection + secondary-shadow rays per second)"


def workload_config(name, scene):
    """The static description of the workload: the same dict in the GPU arm and in the reference arm."""
    scene_idx, W, H, mode, spp, samples, faithful, desc = WORKLOADS[name]
    return {"workload": f"{name}: {desc}", "scene": scene_idx, "width": W, "height": H, "mode": {0: "deferred", 1: "raytraced", 2: "hybrid"}[mode],
            "spp": spp, "shadow_samples": samples, "max_bounces": 8, "rng": "faithful (seed by value)" if faithful else "advancing",
            "triangles": scene.n_triangles, "instances": int(len(scene.instances)), "lights": int(len(scene.lights)), "pixels": W * H}


def b_ray(T):
    """bytes a ray would request on a root-to-leaf walk of a BVH over T triangles (32-B nodes, both children
    inspected per level, 4 leaf triangles, one instance record).  A *request* model: almost all of it is served by L1 / L2."""
    return 32 + 16 + 32 * (1 + 2 * math.ceil(math.log2(max(T, 4) / 4))) + 4 * 36


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy figure)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        # the first sample or two may precede the load (nvidia-smi reports the idle clock then): the median is over the window
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm),
                "sm_mhz_min": min(sm) if sm else None}


def stats_ms(xs):
    xs = np.asarray(xs, dtype=np.float64)
    if xs.size == 0:
        return None
    return {"min": float(xs.min()), "median": float(np.median(xs)), "max": float(xs.max()), "mean": float(xs.mean()), "n": int(xs.size),
            "argmax": int(xs.argmax())}


# ------------------------------------------------------------------------------------------------ the CPU oracle, timed ----
def oracle_steps(scene, name, steps, warmup, budget_s):
    """--impl reference: the oracle on all host cores, K timed steps after W warm-up steps, each step a bounded sample of the
    workload (every stride-th row of the frame, stride chosen so that the whole run fits budget_s)."""
    import vulkanhybridraytracer_b200 as vk
    from oracle.oracle_py import Oracle
    scene_idx, W, H, mode, spp, samples, faithful, desc = WORKLOADS[name]
    cores = host_threads()
    o = Oracle(scene, threads=cores)      # orc_set_num_threads: torchrun exports OMP_NUM_THREADS=1, the oracle must not inherit that
    cores = o.threads()
    cam = scene.camera(W, H)

    def one(frame, stride):
        o.set_row_sampling(stride, stride // 2)
        o.reset_counters()
        p = vk.make_params(W, H, mode=mode, spp=spp, shadow_samples=samples, faithful_rng=faithful, frame=frame)
        t0 = time.perf_counter()
        if mode == 2:
            o.hybrid_frame(cam, p)
        else:
            o.raytrace(cam, p)
        dt = time.perf_counter() - t0
        c = list(o.counters.rays)
        return dt, c[1] + c[2] + c[3], c[0]

    stride = 16 if mode == 2 else 270
    one(1, stride)                                  # first touch (page faults, OpenMP team start-up)
    dt, _, _ = one(1, stride)
    per_row = dt / max(1, len(range(stride // 2, H, stride)))
    rows = max(4, min(H, int(budget_s / max(1, steps + warmup) / max(per_row, 1e-9))))
    stride = max(1, H // rows)
    for w in range(warmup):
        one(1 + w, stride)
    times, tot_rays = [], 0
    for k in range(steps):
        dt, rays, _ = one(1 + warmup + k, stride)
        times.append(dt); tot_rays += rays
    o.set_row_sampling(1, 0)
    o.close()
    rows = len(range(stride // 2, H, stride))
    tot_t = float(sum(times))
    return {"mrays_s": tot_rays / tot_t / 1e6, "ms_per_frame_est": tot_t / steps * 1e3 * (H / rows), "cores": cores, "step_ms": stats_ms(np.array(times) * 1e3),
            "sample": f"{steps} steps x every {stride}th row ({rows} of {H} rows) of the {W}x{H} frame, {tot_t:.1f} s wall on {cores} threads", "seconds": tot_t}


def parity_and_cpu_baseline(vk, ctx, scene, name, chain, budget_s):
    """N = 1: the oracle renders the frames the GPU renders -- whole frames of the benchmark's own size where the oracle manages
    them, a row sample otherwise -- and its wall time for exactly those rows is the cpu_baseline (oracle/parity.py)."""
    from oracle.oracle_py import Oracle
    from oracle import parity
    scene_idx, W, H, mode, spp, samples, faithful, desc = WORKLOADS[name]
    cores = host_threads()
    o = Oracle(scene, threads=cores)
    cores = o.threads()
    cam = scene.camera(W, H)
    # calibrate on a sparse row sample
    stride0 = 32 if mode == 2 else 270
    o.set_row_sampling(stride0, stride0 // 2)
    p = vk.make_params(W, H, mode=mode, spp=spp, shadow_samples=samples, faithful_rng=faithful, frame=1)
    (o.hybrid_frame if mode == 2 else o.raytrace)(cam, p)
    t0 = time.perf_counter()
    (o.hybrid_frame if mode == 2 else o.raytrace)(cam, p)
    per_row = (time.perf_counter() - t0) / len(range(stride0 // 2, H, stride0))
    o.set_row_sampling(1, 0)
    full = per_row * H
    if mode == 2 and 2.5 * full <= budget_s:
        frames = int(max(3, min(12, budget_s / full)))
        par, cpu = parity.run_pair(vk, ctx, o, scene, W, H, mode, spp, samples, faithful, frames=frames, stride=1, chain=chain, time_budget_s=budget_s)
        sample = f"{cpu['frames']} whole frames (indices 0..{cpu['frames'] - 1}, all {H} rows) of the {W}x{H} frame"
    else:
        rows = max(4, min(H, int(budget_s / max(per_row, 1e-9))))
        stride = max(2, H // rows)
        par, cpu = parity.run_pair(vk, ctx, o, scene, W, H, mode, spp, samples, faithful, stride=stride, chain=chain)
        sample = f"1 frame x every {stride}th row ({cpu['rows']} of {H} rows) of the {W}x{H} frame"
    o.close()
    sec = max(cpu["seconds"], 1e-9)
    rows_total = cpu["rows"] * cpu["frames"]
    base = {"value": cpu["rays"] / sec / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
            "sample": f"{sample}, {sec:.1f} s wall on {cores} threads ({sec * cores:.0f} core-seconds); the same rows are the parity check",
            "ms_per_frame_est": sec / max(rows_total, 1) * H * 1e3}
    return par, base


def nccl_debug_to_stderr():
    """The driver checks the communicator size from NCCL's own log ("... nranks N ...").  NCCL writes that log to stdout by default,
    where the contract wants exactly one JSON line: NCCL_DEBUG_FILE=/dev/stderr sends every rank's INIT lines to stderr instead
    (torchrun forwards them).  A caller's INFO / TRACE setting, subsystem list and log file are kept."""
    if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):      # unset, or VERSION / WARN (this image presets VERSION): no "nranks" lines
        os.environ["NCCL_DEBUG"] = "INFO"
        os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT,ENV")
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    if os.environ.get("RANK", "0") == "0":
        print("bench.py: NCCL_DEBUG=%s NCCL_DEBUG_SUBSYS=%s NCCL_DEBUG_FILE=%s" % tuple(os.environ.get(k) for k in ("NCCL_DEBUG", "NCCL_DEBUG_SUBSYS", "NCCL_DEBUG_FILE")),
              file=sys.stderr, flush=True)


def load_ncu_measured():
    """Per-kernel ncu figures of the committed capture (profiles/ncu_kernels.json, written by scripts/ncu_summary.py)."""
    p = os.path.join(ROOT, "profiles", "ncu_kernels.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the oracle leg (no parity block, no cpu_baseline)")
    ap.add_argument("--no-flush", action="store_true", help="do not flush L2 between timed frames")
    ap.add_argument("--nccl-barrier", action="store_true", help="N > 1, peer composite: 4-byte NCCL all-reduce per frame instead of the device-side flag barrier")
    ap.add_argument("--no-peer", action="store_true", help="N > 1: composite with pack + NCCL gather + unpack instead of NVLink peer stores")
    ap.add_argument("--no-verify", action="store_true", help="N > 1: skip the bit-identity check of the composite against a 1-GPU render")
    ap.add_argument("--no-cfg3", action="store_true", help="skip the 2-step Raytraced 64 spp x 64 samples block (BASELINE configs[2])")
    ap.add_argument("--chain", type=int, default=7, choices=[0, 1, 2, 3, 4, 5, 6, 7, 8],
                    help="reflection-chain schedule: 0 inline in k_resolve, 1 wavefront queues, 2 k_chain overlapped with the shadow pass, 3 warp per path, 4 staged first bounce, 5 = 2 + the resolve does not wait for the chain, 6 persistent wavefront, 7 = 5 with the paths split by predicted length (default)")
    ap.add_argument("--fuse-denoise", action="store_true", help="run blur.comp inside the resolve kernel instead of as its own pass (A/B)")
    ap.add_argument("--cpu-budget", type=float, default=20.0, help="wall seconds the parity / cpu_baseline oracle leg may take")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and args.impl != "reference":
        nccl_debug_to_stderr()       # before torch (and with it libnccl) is loaded: NCCL caches its debug settings on first use
    scene_idx, W, H, mode, spp, samples, faithful, desc = WORKLOADS[args.workload]
    peak, peak_src = measured_peak()

    import vulkanhybridraytracer_b200 as vk

    # ---------------------------------------------------------------------------------------- reference arm ----
    if args.impl == "reference":
        if rank != 0:
            return 0
        scene = vk.HostScene(scene_idx)          # libvkp_scene.so only: this arm never maps the CUDA library
        leg = oracle_steps(scene, args.workload, args.steps, args.warmup, budget_s=90.0)
        line = {
            "impl": "reference", "metric": METRIC, "value": leg["mrays_s"],
            "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": leg["ms_per_frame_est"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.workload, scene),
            "note": "the reference (Vulkan RT, MSVC) cannot be built or run here; this arm times the CPU oracle = scalar port of its shaders + a "
                    "binned-SAH BVH, OpenMP over rows, all host threads; ms_per_step is the full-frame estimate from the row sample",
            "frame_ms": leg["step_ms"],
            "cpu_baseline": {"value": leg["mrays_s"], "unit": "Mrays/s", "cores": leg["cores"], "kind": "port", "sample": leg["sample"]},
            "e2e": {"value": leg["mrays_s"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------------------------------------ ours ----
    import torch
    import torch.distributed as dist
    from vulkanhybridraytracer_b200 import PLANE
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = local if world > 1 else 0
    scene = vk.HostScene(scene_idx)
    T = scene.n_triangles
    ctx = vk.Context(dev)
    ctx.upload_scene(scene)
    # acceleration-structure build as a measured component (buildBlas renderer.cpp:1696-1753, buildTlas :1757-1865, per-frame refit
    # vk_engine.cpp:172-187): CUDA-event times from the library, against the algorithmic bytes of SURVEY 8d (132 B read + 80 B written per triangle)
    for _ in range(3):
        ctx.build_tlas(True)
    ctx.sync()
    build = ctx.build_stats()
    build["blas_ms_first"] = build["blas_ms"]          # the process's first build: includes loading every build kernel (lazy module loading)
    if world == 1 or rank == 0:
        warm = vk.Context(dev); warm.upload_scene(scene)       # the same builds again in a fresh context: what a scene change costs
        w2 = warm.build_stats(); warm.close()
        build["blas_ms"] = w2["blas_ms"]; build["tlas_ms"] = w2["tlas_ms"]
        build["blas_launches"] = w2["blas_launches"]; build["blas_host_syncs"] = w2["blas_host_syncs"]
    if build["blas_ms"] > 0:
        build["blas_algorithmic_bytes"] = int(build["blas_triangles"]) * (132 + 80)
        build["blas_GBps"] = build["blas_algorithmic_bytes"] / (build["blas_ms"] * 1e-3) / 1e9
        build["blas_mtris_s"] = build["blas_triangles"] / (build["blas_ms"] * 1e-3) / 1e6
        build["builder"] = os.environ.get("VKP_BLAS_BUILDER", "ploc")
        build["how"] = ("CUDA events around every BLAS build of the scene + the compaction into one arena, queued without a host wait "
                        "(device-driven PLOC rounds), one synchronisation at the end; blas_ms = second context (kernels loaded), blas_ms_first = first")
    cam = scene.camera(W, H)
    ctx.set_camera(cam)
    TW, TH = (32, 16)
    ctx.set_tiling(rank, world, TW, TH)
    L = len(scene.lights)

    def params(frame, wl=None):
        _, w_, h_, mode_, spp_, samples_, faithful_, _ = WORKLOADS[wl or args.workload]
        return vk.make_params(w_, h_, mode=mode_, spp=spp_, shadow_samples=samples_, faithful_rng=faithful_, frame=frame, wavefront=args.chain,
                              fuse_denoise=1 if args.fuse_denoise else 0)

    from vulkanhybridraytracer_b200.multigpu import TiledRenderer
    tr = TiledRenderer(ctx, rank, world, L, tile=(TW, TH), backend="nccl", device=torch.device("cuda", dev), stream_ordered=world > 1)

    peer = False
    if world > 1 and not args.no_peer:
        ctx.set_params(params(0))
        peer = tr.enable_peer_output(device_barrier=not args.nccl_barrier)

    def composite():
        return tr.composite(PLANE.FINAL_SRGB8)

    def barrier():
        ctx.sync()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ctx.set_params(params(0)); ctx.reset_history()
    frame = 0
    sampler = ClockSampler(dev); sampler.start()      # nvidia-smi starts before the warm-up: the timed region itself is a few tens of ms
    for _ in range(args.warmup):                      # warm-up: frame 0 (reset frame, with its halo exchange for N > 1) .. W-1
        tr.render(params(frame))
        if world > 1:
            composite()
        frame += 1
    ctx.sync()

    # -- timed region 1: device time, inputs resident, K frames, L2 flushed between frames ---------------------------------------
    ctx.set_profiling(True)
    stage_ms = np.zeros(8)
    rays = np.zeros(4, dtype=np.int64)
    trav = np.zeros(3, dtype=np.int64)
    launches = 0
    frame_times = []
    import gc
    gc.collect(); gc.disable()          # no collector pause while frames are being queued (at N = 8 the host is only ~1.5x ahead of the device)
    barrier()
    total_ms = 0.0
    if world == 1:
        for k in range(args.steps):
            if not args.no_flush:
                ctx.flush_l2()
            p_k = params(frame)
            ctx.timer_begin()
            tr.render(p_k)
            ms = ctx.timer_end()
            total_ms += ms; frame_times.append(ms)
            c = ctx.counters()
            stage_ms += np.array(list(c.ms_stage))
            rays += np.array([c.rays_primary, c.rays_shadow, c.rays_reflect, c.rays_secondary_shadow], dtype=np.int64)
            launches += c.kernel_launches
            frame += 1
    else:
        # N ranks: the K frames are queued back to back on the rank's stream -- flush, [event] render + composite [event] -- with no
        # host round trip in between (a per-frame host sync would put the enqueue latency of the barrier into every frame);
        # the frame barrier keeps the ranks in lockstep on the device.  Ray counters accumulate.
        ctx.set_profiling(False)
        ctx.set_counter_accumulation(True)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for k in range(args.steps):
            if not args.no_flush:
                ctx.flush_l2()
            ev[k][0].record(tr.stream)
            tr.render(params(frame))
            launches += composite()
            ev[k][1].record(tr.stream)
            frame += 1
        barrier()
        frame_times = [a.elapsed_time(b) for a, b in ev]
        total_ms = float(sum(frame_times))
        c = ctx.counters()
        rays += np.array([c.rays_primary, c.rays_shadow, c.rays_reflect, c.rays_secondary_shadow], dtype=np.int64)
        launches += c.kernel_launches * args.steps
        ctx.set_counter_accumulation(False)
        # per-stage times: a few more frames, profiled one by one (not part of `value`)
        ctx.set_profiling(True)
        n_prof = min(8, args.steps)
        for k in range(n_prof):
            tr.render(params(frame)); composite()
            c = ctx.counters()
            stage_ms += np.array(list(c.ms_stage)) * (args.steps / n_prof)
            frame += 1
    barrier()
    gc.enable()
    # one extra, untimed frame with traversal accounting on: what the rays actually requested from memory
    ctx.set_profiling(2)
    tr.render(params(frame))
    c = ctx.counters()
    trav = np.array([c.node_fetches, c.triangle_tests, c.instance_entries], dtype=np.int64) * args.steps
    if world > 1:
        composite()
    frame += 1
    ctx.set_profiling(False)
    t = torch.tensor([total_ms], dtype=torch.float64, device=f"cuda:{dev}")
    r = torch.tensor(rays, dtype=torch.int64, device=f"cuda:{dev}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)            # max over ranks
        dist.all_reduce(r, op=dist.ReduceOp.SUM)            # whole-job rays
    total_ms = float(t.item()); rays_all = r.cpu().numpy()
    ms_per_step = total_ms / args.steps
    counted = int(rays_all[1] + rays_all[2] + rays_all[3])
    value = counted / args.steps / (ms_per_step * 1e-3) / 1e6

    # -- clocks: K timed frames are over in 10-30 ms, less than one nvidia-smi sampling period.  The same frames go on, untimed, until
    #    the GPU has been under this load for ~0.7 s since the warm-up began (the same count on every rank), and the sampler -- started
    #    before the warm-up -- reports the median over that whole window: warm-up + timed frames + hold frames of one workload.
    hold = int(min(4000, max(0.0, 700.0 - args.steps * ms_per_step) / max(ms_per_step, 1e-3)))
    for _ in range(hold):
        tr.render(params(frame))
        if world > 1:
            composite()
        frame += 1
    barrier()
    clocks = sampler.stop()
    clocks["window"] = f"warm-up ({args.warmup}) + timed ({args.steps}) + {hold} untimed frames of the same workload, nvidia-smi -lms 100"

    # -- N > 1: the composited image must be bit-identical to the same frame sequence rendered on one GPU ---------------------------
    composite_ok = None
    if world > 1 and not args.no_verify:
        if rank == 0:
            got = ctx.readback(PLANE.FINAL_SRGB8).copy()
            solo = vk.Context(dev); solo.upload_scene(scene); solo.set_camera(cam)
            for f in range(frame):
                solo.set_params(params(f)); solo.render()
            composite_ok = bool(np.array_equal(got, solo.readback(PLANE.FINAL_SRGB8)))
            solo.close()
        dist.barrier()

    # -- BASELINE configs[2] (Raytraced, 64 spp x 64 shadow samples), 2 timed frames after 1: the workload north_star's 80 % target is on ---
    cfg3 = None
    if not args.no_cfg3 and args.workload == "cfg2":
        ctx.set_profiling(False)
        ctx.set_counter_accumulation(True)
        ctx.set_params(params(0, "cfg3")); ctx.reset_history()
        tr.render(params(0, "cfg3"))
        if world > 1:
            composite()
        barrier()
        ctx.counters()                                     # clears the accumulated counters
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stream3 = tr.stream if world > 1 else None
        if world == 1:
            ctx.timer_begin()
        else:
            e0.record(stream3)
        for k in range(2):
            tr.render(params(1 + k, "cfg3"))
            if world > 1:
                composite()
        if world == 1:
            ms3 = ctx.timer_end()
        else:
            e1.record(stream3)
            barrier()
            ms3 = e0.elapsed_time(e1)
        c3 = ctx.counters()
        ctx.set_counter_accumulation(False)
        t3 = torch.tensor([ms3], dtype=torch.float64, device=f"cuda:{dev}")
        r3 = torch.tensor([c3.rays_reflect + c3.rays_secondary_shadow], dtype=torch.int64, device=f"cuda:{dev}")
        if world > 1:
            dist.all_reduce(t3, op=dist.ReduceOp.MAX); dist.all_reduce(r3, op=dist.ReduceOp.SUM)
        cfg3 = {"workload": "cfg3: " + WORKLOADS["cfg3"][7], "steps": 2, "warmup": 1, "ms_per_step": float(t3.item()) / 2,
                "mrays_s": float(r3.item()) / 2 / (float(t3.item()) / 2 * 1e-3) / 1e6, "n_gpus": world, "scaling": "strong",
                "timing": "CUDA events around 2 frames incl. the composite, max over ranks, no L2 flush (a frame is 1.6e10 rays)"}
        ctx.set_params(params(frame)); ctx.reset_history()

    # -- timed region 2: end to end through the reference-facing host API (VulkanEngine::update + rasterize_hybrid + read-back) ---
    e2e = None
    parity = None
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        # parity at the benchmark's own size + the cpu_baseline, from the same oracle frames (before the context goes away)
        ctx.set_tiling(0, 1, TW, TH)
        parity, cpu = parity_and_cpu_baseline(vk, ctx, scene, args.workload, args.chain, args.cpu_budget)
    if world == 1 and mode == 2 and scene_idx == 0:
        ctx.close()
        eng = vk.HostEngine(scene_idx, W, H, dev)
        eng.set("chain", args.chain)
        eng.set("samples", samples)
        pinned = [torch.empty((H, W, 4), dtype=torch.uint8).pin_memory() for _ in range(2)]
        nbytes = pinned[0].numel()
        # (1) fully synchronous: every frame is read back before the next one is submitted (vkQueueWaitIdle per frame)
        for _ in range(args.warmup):
            eng.frames(1)
            eng._H.vkph_engine_read(eng._h, 0, pinned[0].data_ptr(), nbytes)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(args.steps):
            eng.frames(1)          # update(): lights + instance matrices H2D, TLAS refit; then the frame kernels
            eng._H.vkph_engine_read(eng._h, 0, pinned[0].data_ptr(), nbytes)      # D2H of the presented image into pinned memory
        dt_sync = time.perf_counter() - t0
        # (2) the reference's FRAME_OVERLAP = 2 (renderer.h:34): frame k's image travels to the host while frame k+1 renders
        #     into the other output buffer; every step still uploads its inputs and brings one finished image into host memory
        for k in range(args.warmup):
            eng.frames(1); eng.present(pinned[k & 1].data_ptr(), nbytes); eng.present_wait(1)
        eng.present_wait(0)
        torch.cuda.synchronize()
        eng.counters()
        iter_t = []
        t0 = time.perf_counter()
        for k in range(args.steps):
            ta = time.perf_counter()
            eng.frames(1)
            eng.present(pinned[k & 1].data_ptr(), nbytes)
            eng.present_wait(1)        # frame k-1 is complete in pinned[(k-1) & 1]
            iter_t.append((time.perf_counter() - ta) * 1e3)
        eng.present_wait(0)
        dt = time.perf_counter() - t0
        c = eng.counters()
        e_rays = (c.rays_shadow + c.rays_reflect + c.rays_secondary_shadow)
        present_ok = None
        if not args.no_verify:
            last = pinned[(args.steps - 1) & 1].numpy().copy()
            present_ok = bool(np.array_equal(last, eng.read_srgb8()))
        e2e = {"value": e_rays / (dt / args.steps) / 1e6, "unit": "Mrays/s", "ms_per_step": dt / args.steps * 1e3,
               "h2d_bytes_per_step": 48 * L + 144 * len(scene.instances) + 400 + 64, "d2h_bytes_per_step": W * H * 4,
               "frames_in_flight": 2, "presented_image_equals_sync_readback": present_ok, "iteration_ms": stats_ms(iter_t),
               "sync": {"value": e_rays / (dt_sync / args.steps) / 1e6, "ms_per_step": dt_sync / args.steps * 1e3, "frames_in_flight": 1},
               "api": "VulkanEngine::update + Renderer::rasterize_hybrid + present (D2H of the sRGB8 image into pinned memory) through the host C++ "
                      "API over the C-ABI, wall clock; FRAME_OVERLAP = 2 as renderer.h:34",
               "vs_value": "e2e keeps two frames in flight and does not flush L2 between frames; `value` times each frame alone after an L2 flush, so "
                           "e2e may be the faster of the two"}
        eng.close()
    else:
        # tile-sharded / other workloads: same loop through the C-ABI with per-frame host inputs and a present on rank 0.
        # Two frames in flight (FRAME_OVERLAP = 2, renderer.h:34): frame k's image travels to pinned memory while frame k+1 renders
        # into the other output buffer.  With peer output the other ranks store into rank 0's buffers, so rank 0 lets the copy of
        # frame k-1 land before it joins the barrier that releases frame k+1 (which reuses that buffer).
        pinned = [torch.empty((H, W, 4), dtype=torch.uint8).pin_memory() for _ in range(2)]
        nbytes = pinned[0].numel()
        ctx.set_tiling(rank, world, TW, TH)

        def e2e_frame(k):
            ctx.set_lights(scene.lights); ctx.set_camera(cam)
            if world == 1:
                ctx.set_output_buffer(k & 1)
            tr.render(params(frame + k))
            if world > 1:
                if rank == 0 and not getattr(tr, "device_barrier", False):
                    ctx.readback_wait(0)
                tr.composite(PLANE.FINAL_SRGB8, release=False)
            if rank == 0:
                ctx.readback_async(PLANE.FINAL_SRGB8, pinned[k & 1].data_ptr(), nbytes)
                if world == 1:
                    ctx.readback_wait(1)
                else:
                    tr.release(after_present=True)      # the peers may reuse this buffer once the copy has landed
                    ctx.readback_wait(1)
        for k in range(args.warmup):
            e2e_frame(k)
        if rank == 0:
            ctx.readback_wait(0)
        ctx.sync()
        barrier()
        t0 = time.perf_counter()
        for k in range(args.steps):
            e2e_frame(args.warmup + k)
        if rank == 0:
            ctx.readback_wait(0)
        ctx.sync()
        barrier()
        dt = time.perf_counter() - t0
        if world == 1:
            ctx.set_output_buffer(0)
        tt = torch.tensor([dt], dtype=torch.float64, device=f"cuda:{dev}")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        e2e = {"value": counted / args.steps / (dt / args.steps) / 1e6, "unit": "Mrays/s", "ms_per_step": dt / args.steps * 1e3,
               "h2d_bytes_per_step": 48 * L + 400 + 64, "d2h_bytes_per_step": W * H * 4, "frames_in_flight": 2,
               "api": "vkp_set_lights/camera/params + vkp_render (+ NVLink peer composite) + vkp_readback_async(FINAL_SRGB8) into pinned memory on rank 0, wall clock",
               "vs_value": "e2e keeps two frames in flight and does not flush L2 between frames; `value` times each frame alone after an L2 flush, so "
                           "e2e may be the faster of the two"}

    comm_nranks = None
    if world > 1:
        dist.barrier()
        comm_nranks = dist.get_world_size()          # NCCL's own "nranks" lines are on stderr (nccl_debug_to_stderr)
    if rank != 0:
        if world > 1:
            dist.barrier(); dist.destroy_process_group()
        return 0

    # -- roofline of the dominant kernel -----------------------------------------------------------------------------------------------
    # achieved = algorithmic bytes / CUDA-event duration.  Algorithmic = what the stage has to move to and from HBM once:
    # its planes, plus the part of the acceleration structure its rays touch -- at most all of it, because a node fetched by a second
    # ray is an L2 / L1 hit, not HBM traffic.  SURVEY 8d's per-ray figure R * B_ray(T) is kept beside it as `requested_model`: a request
    # volume the caches serve, which is why it may exceed the HBM peak.
    P = W * H
    names = ["k_gbuffer", "k_shadow", "k_denoise", "k_resolve", "k_post", "k_raytrace", "k_deferred"]
    per_frame = rays_all / args.steps / world          # this rank's share (tiles are balanced)
    Pr = P / world
    BR = b_ray(T)
    bvh_bytes = 32.0 * (T / 2.0) + 48.0 * T             # ~T/2 nodes of 32 B (<= 4-triangle leaves) + 48 B per leaf-ordered triangle
    plane_bytes = {"k_gbuffer": Pr * 40, "k_shadow": Pr * (16 + L), "k_denoise": Pr * (2 * L + 4 + L), "k_resolve": Pr * (32 + L + 16 + 4),
                   "k_raytrace": Pr * (16 + 16 + 4)}
    # accumulation frames (every timed frame here): k_shadow's threads read the history byte and write the new one themselves and the filter
    # pass is not launched (library default; VKP_FUSE_ACCUMULATE=0 or --fuse-denoise keep / move the pass) -- the bytes move with the work
    if mode == 2 and samples < 4 and not args.fuse_denoise and os.environ.get("VKP_FUSE_ACCUMULATE", "1") != "0":
        plane_bytes["k_shadow"] += Pr * 2 * L
        plane_bytes["k_denoise"] = 0.0
    ray_count = {"k_gbuffer": per_frame[0], "k_shadow": per_frame[1], "k_denoise": 0.0, "k_resolve": per_frame[2] + per_frame[3],
                 "k_raytrace": per_frame[2] + per_frame[3]}
    kbytes = {k: plane_bytes[k] + min(ray_count[k] * BR, bvh_bytes if ray_count[k] > 0 else 0.0) for k in plane_bytes}
    kreq = {k: plane_bytes[k] + ray_count[k] * BR for k in plane_bytes}
    ms = stage_ms / args.steps
    chain_stream_ms = float(ms[7])                       # the reflection chain on its own stream (start to end), beside the main stream's stages
    dom = int(np.argmax(ms[:7]))
    dname = names[dom]
    achieved = kbytes.get(dname, 0.0) / (ms[dom] * 1e-3) / 1e9 if ms[dom] > 0 else 0.0
    frame_kernels = ("k_gbuffer", "k_shadow", "k_denoise", "k_resolve") if mode == 2 else ("k_raytrace",)
    frame_bytes = sum(kbytes[n] for n in frame_kernels)
    ncu = load_ncu_measured() or {}
    ncu_k = ncu.get("kernels", {}).get(dname)
    bound = "hbm"
    if ncu_k and ncu_k.get("dram_frac") is not None and ncu_k["dram_frac"] < 0.5:
        bound = ncu_k.get("bound", "issue+latency")
    roofline = {"bound": bound, "kernel": dname, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": kbytes.get(dname, 0.0),
                "algorithmic_model": "planes of the stage + min(rays * B_ray(T), bytes of the BVH): the compulsory HBM traffic (DESIGN.md section 5)",
                "stage_ms": float(ms[dom]), "stage_note": "CUDA-event time of the stage inside the frame (k_shadow shares the SMs with the concurrent k_chain); "
                                                          "the kernel alone is measured.kernel_us",
                "kernel_share_of_step": float(ms[dom] / max(ms.sum(), 1e-9)),
                "stage_ms_all": {names[i]: float(ms[i]) for i in range(7) if ms[i] > 0},
                "chain_stream_ms": chain_stream_ms,
                "frame": {"algorithmic_bytes": frame_bytes, "achieved": frame_bytes * world / (ms_per_step * 1e-3) / 1e9,
                          "frac": frame_bytes * world / (ms_per_step * 1e-3) / 1e9 / (peak * world)},
                "requested_model": {"note": "SURVEY 8d: planes + rays * B_ray(T), B_ray = %d B at T = %d -- a request volume served by L1 / L2, not an HBM figure" % (BR, T),
                                    "bytes_per_launch": kreq.get(dname, 0.0), "GBps": kreq.get(dname, 0.0) / (ms[dom] * 1e-3) / 1e9 if ms[dom] > 0 else 0.0},
                "requested_counted": {"note": "bytes the traversal actually asked the memory system for on this rank (device counters): 64 B per child-pair fetch, "
                                              "48 B per triangle, 80 B per instance entry, + the per-pixel plane traffic",
                                      "node_fetches_per_frame": int(trav[0] // args.steps), "triangle_tests_per_frame": int(trav[1] // args.steps),
                                      "instance_entries_per_frame": int(trav[2] // args.steps),
                                      "bytes_per_frame": float((64 * trav[0] + 48 * trav[1] + 80 * trav[2]) / args.steps + Pr * (100 + 5 * L)),
                                      "GBps": float(((64 * trav[0] + 48 * trav[1] + 80 * trav[2]) / args.steps + Pr * (100 + 5 * L)) / (ms_per_step * 1e-3) / 1e9)},
                "measured": ncu_k, "measured_source": ncu.get("_source")}
    if ncu_k and ncu_k.get("dram_bytes") is not None:
        roofline["traffic"] = ncu_k["dram_bytes"]
    else:
        prof = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(prof):
            try:
                tj = json.load(open(prof))
                roofline["traffic"] = tj.get(dname, tj.get(dname + "<2>"))
                roofline["traffic_source"] = tj.get("_source")
            except Exception:
                pass

    line = {
        "metric": METRIC, "value": value, "unit": "Mrays/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, scene),
        "run": {"rays_per_frame": {"shadow": int(rays_all[1] // args.steps), "reflect": int(rays_all[2] // args.steps),
                                   "secondary_shadow": int(rays_all[3] // args.steps), "primary_not_counted": int(rays_all[0] // args.steps)},
                "mrays_s_incl_primary": float(rays_all.sum()) / args.steps / (ms_per_step * 1e-3) / 1e6,
                "nominal_2P_mrays_s": 2 * P / (ms_per_step * 1e-3) / 1e6,
                "chain_schedule": {0: "inline in k_resolve", 1: "wavefront queues (k_bounce per bounce)", 2: "k_chain (one thread per queued path) on a second stream overlapping k_shadow + k_denoise", 3: "k_chain_warp (one warp per queued path, cooperative traversal) on a second stream", 4: "staged chain on a second stream overlapping k_shadow + k_denoise: first bounce as a wavefront (k_chain_first, k_chain_shadow0, k_chain_shade0), continuing paths compacted (k_chain_rest)", 5: "k_chain (one thread per queued path) on a second stream overlapping k_shadow + k_denoise + k_resolve; k_resolve_paths merges the chain into the mirror / glass pixels", 6: "k_chain_q: persistent queue-fed grid, one bounce per work item, scattered paths appended to the queue (ballot / prefix compaction), on a second stream overlapping k_shadow + k_denoise + k_resolve; k_resolve_paths merges it", 7: "paths split by last frame's length: the long ones first in k_chain (persistent, queue-fed), the others through the staged kernels (k_chain_first / _shadow0 / _shade0 / _rest) on a third stream, both overlapping k_shadow + k_denoise + k_resolve; k_resolve_paths merges them", 8: "as 7 with the first bounce folded into the frame's own kernels: k_gbuffer queues the path pixels (no scan pass), the first reflection rays are the leading blocks of k_shadow's grid, k_chain_shadow0 / _shade0 / _rest follow the shadow pass on a third stream beside k_denoise + k_resolve; k_resolve_paths merges"}[args.chain],
                "l2": "flushed between timed frames (256 MiB write)" if not args.no_flush else "not flushed",
                "parallelism": (f"tiles {TW}x{TH} dealt over {world} GPU(s), replicated BVH, " +
                                ("resolve kernels store sRGB8 pixels straight into rank 0's image over NVLink peer memory (CUDA IPC) + a device-side flag barrier"
                                 if peer else "pack + NCCL gather of sRGB8 tiles + unpack")) if world > 1 else "1 GPU",
                "composite_bit_identical_to_1gpu": composite_ok,
                "comm_world_size": comm_nranks,
                "timing": "CUDA events on the context stream around each frame, summed over K frames, max over ranks"},
        "frame_ms": stats_ms(frame_times),
        "roofline": roofline, "cpu_baseline": cpu, "parity": parity, "build": build, "cfg3": cfg3, "e2e": e2e, "clocks": clocks, "gpu_launches": int(launches),
    }
    print(json.dumps(line))
    sys.stdout.flush()
    if world > 1:
        dist.barrier(); dist.destroy_process_group()
    if parity is not None and not parity.get("pass", False):
        print("bench.py: PARITY BELOW THE BAR: %s" % json.dumps({k: parity[k] for k in ("hit_id_agreement", "mae", "max")}), file=sys.stderr)
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
