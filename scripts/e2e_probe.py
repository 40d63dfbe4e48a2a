"""Where does a frame's wall clock go on the host side?  Times VulkanEngine::update+render, present and present_wait separately."""
import sys, time
import numpy as np
import torch
import vulkanhybridraytracer_b200 as vk

W, H = 1920, 1080
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
