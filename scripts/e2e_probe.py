"""Host-side frame-loop probes used for DESIGN.md section 4 (one script, one probe per sub-command).

  host       Where does a frame's wall clock go on the host side?  Times VulkanEngine::update+render, present and present_wait separately.
  overlap    Does the present copy overlap the next frame?  Direct C-ABI loop with / without per-frame uniform uploads.
  iterations Per-iteration wall clock of the frame loop in several modes: where do the occasional long iterations come from?
  outliers   Outlier frames: is the GPU slow (per-stage CUDA events) or the host late?  Direct C-ABI, synchronous frames.
  tail       Which thread sets the tail of k_shadow in the slow frames?

Usage: python scripts/e2e_probe.py {host|overlap|iterations|outliers|tail}
"""
import sys
import time

import numpy as np
import torch

import vulkanhybridraytracer_b200 as vk
from vulkanhybridraytracer_b200 import PLANE, STAGE

W, H = 1920, 1080


def probe_host():
    eng = vk.HostEngine(0, W, H, 0)
    eng.set("chain", 2)
    pinned = [torch.empty((H, W, 4), dtype=torch.uint8).pin_memory() for _ in range(3)]
    n = pinned[0].numel()
    for k in range(5):
        eng.frames(1); eng.read_srgb8()
    for mode in ("sync", "pipe", "pipe2", "pipe_nowait", "render_only", "pipe"):
        torch.cuda.synchronize()
        t = np.zeros(3)
        t00 = time.perf_counter()
        for k in range(50):
            t0 = time.perf_counter(); eng.frames(1); t1 = time.perf_counter()
            if mode == "sync":
                eng._H.vkph_engine_read(eng._h, 0, pinned[0].data_ptr(), n); t2 = t3 = time.perf_counter()
            elif mode == "render_only":
                t2 = t3 = t1
            else:
                eng.present(pinned[k % 3].data_ptr(), n); t2 = time.perf_counter()
                if mode == "pipe":
                    eng.present_wait(1)
                elif mode == "pipe2":
                    eng.present_wait(2)
                t3 = time.perf_counter()
            t += (t1 - t0, t2 - t1, t3 - t2)
        eng.present_wait(0); torch.cuda.synchronize()
        tot = time.perf_counter() - t00
        print(mode, "total ms/frame %.3f" % (tot / 50 * 1e3), "frames() %.3f  read/present %.3f  wait %.3f" % tuple(t / 50 * 1e3))
    eng.close()


def probe_overlap():
    sc = vk.HostScene(0)
    ctx = vk.Context(0); ctx.upload_scene(sc)
    cam = sc.camera(W, H); ctx.set_camera(cam)
    pinned = [torch.empty((H, W, 4), dtype=torch.uint8).pin_memory() for _ in range(2)]
    n = pinned[0].numel()
    fno = [0]
    def frame(upload, refit):
        if upload:
            ctx.set_lights(sc.lights)
        if refit:
            ctx.build_tlas(True)
        ctx.set_params(vk.make_params(W, H, frame=fno[0], wavefront=2)); fno[0] += 1
        ctx.render()
    for k in range(5):
        frame(True, True); ctx.readback(PLANE.FINAL_SRGB8)
    for upload, refit, present in ((0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 0, 2)):
        ctx.sync(); ctx.readback_wait(0)
        t0 = time.perf_counter()
        for k in range(60):
            ctx.set_output_buffer(k & 1)
            frame(upload, refit)
            if present == 1:
                ctx.readback_async(PLANE.FINAL_SRGB8, pinned[k & 1].data_ptr(), n)
                ctx.readback_wait(1)
            elif present == 2:
                ctx.L.vkp_readback(ctx.h, PLANE.FINAL_SRGB8, pinned[0].data_ptr(), n)
            elif k % 4 == 3:
                ctx.sync()
        ctx.readback_wait(0); ctx.sync()
        dt = (time.perf_counter() - t0) / 60 * 1e3
        print(f"upload={upload} refit={refit} present={present}: {dt:.3f} ms/frame")


def probe_iterations():
    eng = vk.HostEngine(0, W, H, 0)
    pinned = [torch.empty((H, W, 4), dtype=torch.uint8).pin_memory() for _ in range(2)]
    n = pinned[0].numel()
    def run(mode, chain, N=200):
        eng.set("chain", chain)
        for k in range(10):
            eng.frames(1); eng._H.vkph_engine_read(eng._h, 0, pinned[0].data_ptr(), n)
        ts = [time.perf_counter()]; parts = []
        for k in range(N):
            a = time.perf_counter(); eng.frames(1); b = time.perf_counter()
            if mode == "pipe":
                eng.present(pinned[k & 1].data_ptr(), n); c = time.perf_counter(); eng.present_wait(1)
            elif mode == "sync":
                c = b; eng._H.vkph_engine_read(eng._h, 0, pinned[0].data_ptr(), n)
            elif mode == "render_sync":
                c = b; eng.renderer_sync()
            d = time.perf_counter()
            parts.append((b - a, c - b, d - c)); ts.append(d)
        eng.present_wait(0)
        d = np.diff(np.array(ts)) * 1e3; parts = np.array(parts) * 1e3
        out = np.where(d > 1.3 * np.median(d))[0]
        print(f"{mode:12s} chain={chain} median {np.median(d):.3f} mean {d.mean():.3f} p90 {np.percentile(d, 90):.3f} outliers {len(out)}/{N} at {out[:14].tolist()}")
        for i in out[:4]:
            print("     iter", i, "total %.3f  frames %.3f present %.3f wait %.3f" % (d[i], *parts[i]))
    eng.renderer_sync = lambda: eng.counters()
    for mode, chain in (("pipe", 2), ("pipe", 0), ("sync", 2), ("sync", 0), ("render_sync", 2), ("render_sync", 0)):
        run(mode, chain)
    eng.close()


def probe_outliers():
    sc = vk.HostScene(0)
    ctx = vk.Context(0); ctx.upload_scene(sc)
    ctx.set_camera(sc.camera(W, H))
    ctx.set_profiling(1)
    for refit in (0, 1):
        for chain in (2, 0):
            rows = []
            for k in range(240):
                t0 = time.perf_counter()
                if refit:
                    ctx.set_lights(sc.lights); ctx.build_tlas(True)
                ctx.set_params(vk.make_params(W, H, frame=k, wavefront=chain))
                ctx.render(); t1 = time.perf_counter()
                ctx.sync(); t2 = time.perf_counter()
                c = ctx.counters()
                rows.append(((t2 - t0) * 1e3, (t1 - t0) * 1e3, c.ms_frame, *[c.ms_stage[i] for i in range(8)]))
            r = np.array(rows[40:])
            out = np.where(r[:, 0] > 1.3 * np.median(r[:, 0]))[0]
            print(f"refit={refit} chain={chain}: wall median {np.median(r[:,0]):.3f} mean {r[:,0].mean():.3f} | gpu ms_frame median {np.median(r[:,2]):.3f} mean {r[:,2].mean():.3f} | outliers {len(out)}/200")
            for i in out[:5]:
                print("    wall %.3f submit %.3f gpu_frame %.3f stages" % tuple(r[i, :3]), np.round(r[i, 3:], 3))
            i = int(np.argsort(r[:, 0])[len(r) // 2])
            print("    typical: wall %.3f submit %.3f gpu_frame %.3f stages" % tuple(r[i, :3]), np.round(r[i, 3:], 3))


def probe_tail():
    sc = vk.HostScene(0)
    ctx = vk.Context(0); ctx.upload_scene(sc)
    ctx.set_camera(sc.camera(W, H))
    rows = []
    for k in range(120):
        ctx.set_params(vk.make_params(W, H, frame=k, wavefront=2))
        ctx.set_profiling(1); ctx.render(); ctx.sync(); c = ctx.counters()
        ms_shadow = c.ms_stage[STAGE.SHADOW]
        ctx.set_profiling(2); ctx.run_stage(STAGE.SHADOW); ctx.sync(); c2 = ctx.counters()
        rows.append((k, ms_shadow, c2.worst_thread_nodes, c2.worst_thread_tag, c2.node_fetches))
    r = np.array(rows, dtype=np.float64)
    print("corr(ms_shadow, worst_thread_nodes) = %.3f" % np.corrcoef(r[:, 1], r[:, 2])[0, 1])
    for k, ms, wn, tag, nf in rows:
        tag = int(tag); pix = tag & ((1 << 23) - 1)
        if ms > 0.6 or k < 12:
            print(f"frame {k:3d} k_shadow {ms:.3f} ms  worst thread: {int(wn):6d} node fetches, light {(tag >> 23) & 31}, pixel ({pix % W},{pix // W}); total fetches {int(nf)}")
    ids = ctx.readback(PLANE.HIT_IDS)
    pos = ctx.readback(PLANE.POSITION).astype(np.float32)
    for k, ms, wn, tag, nf in rows:
        if ms > 0.6:
            pix = int(tag) & ((1 << 23) - 1); x, y = pix % W, pix // W
            print("   pixel", (x, y), "hit ids", ids[y, x], "pos", pos[y, x])


if __name__ == "__main__":
    probes = {"host": probe_host, "overlap": probe_overlap, "iterations": probe_iterations, "outliers": probe_outliers, "tail": probe_tail}
    if len(sys.argv) != 2 or sys.argv[1] not in probes:
        sys.exit(__doc__)
    probes[sys.argv[1]]()
