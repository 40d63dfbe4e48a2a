"""How long are the reflection / refraction paths of the 1080p default frame, and what does a bounce cost?
For max_bounces = 1 .. 8: paths still alive at that bounce (difference of the reflection-ray counters) and the time of the chain
kernels alone (the RESOLVE stage run on its own launches queue clear + k_chain_scan + k_chain serially, then k_resolve<2>)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vulkanhybridraytracer_b200 as vk
from vulkanhybridraytracer_b200 import STAGE

W, H = 1920, 1080
sc = vk.HostScene(0)
ctx = vk.Context(0); ctx.upload_scene(sc)
ctx.set_camera(sc.camera(W, H))
ctx.set_profiling(1)
prev = 0
base = None
for chain in [int(c) for c in (sys.argv[1:] or ["2", "3"])]:
    print("chain mode", chain)
    for mb in range(0, 9):
        ts = []
        for k in range(6):
            ctx.set_params(vk.make_params(W, H, frame=3 + k, wavefront=chain, max_bounces=mb))
            ctx.run_stage(STAGE.GBUFFER); ctx.run_stage(STAGE.SHADOW); ctx.run_stage(STAGE.DENOISE)
            ctx.run_stage(STAGE.RESOLVE); ctx.sync()
            c = ctx.counters()
            ts.append(c.ms_stage[STAGE.RESOLVE])
        t = float(np.median(ts[1:]))
        if mb == 0:
            base = t
        print(f"  max_bounces {mb}: reflection rays {c.rays_reflect:7d} (+{c.rays_reflect - prev:6d} paths alive at this bounce), secondary shadow {c.rays_secondary_shadow:7d}, "
              f"chain + resolve {t:.3f} ms (chain ~ {t - base:.3f})")
        prev = c.rays_reflect
    prev = 0
