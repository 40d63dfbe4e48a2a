"""Per-iteration wall clock of the frame loop in several modes: where do the occasional long iterations come from?"""
import sys, time
import numpy as np
import torch
import vulkanhybridraytracer_b200 as vk

W, H = 1920, 1080
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
