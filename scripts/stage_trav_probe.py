"""Node-pair fetches and triangle tests per ray, per stage (G-buffer / shadow / reflection chain), for the BLAS builders."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vulkanhybridraytracer_b200 as vk
from vulkanhybridraytracer_b200 import STAGE

W, H = 1920, 1080
sc = vk.HostScene(0)
for builder in ("ploc", "lbvh"):
    os.environ["VKP_BLAS_BUILDER"] = builder
    ctx = vk.Context(0); ctx.upload_scene(sc)
    ctx.set_camera(sc.camera(W, H))
    ctx.set_profiling(2)
    ctx.set_params(vk.make_params(W, H, frame=3, wavefront=2))
    print("builder", builder, ctx.build_stats()["blas_height_max"])
    for st, name in ((STAGE.GBUFFER, "gbuffer"), (STAGE.SHADOW, "shadow"), (STAGE.DENOISE, "denoise"), (STAGE.RESOLVE, "chain+resolve")):
        ctx.run_stage(st); ctx.sync()
        c = ctx.counters()
        rays = {"gbuffer": c.rays_primary, "shadow": c.rays_shadow, "chain+resolve": c.rays_reflect + c.rays_secondary_shadow}.get(name, 0)
        if rays:
            print(f"  {name:14s} rays {rays:8d}  node pairs/ray {c.node_fetches / rays:6.1f}  triangles/ray {c.triangle_tests / rays:5.2f}  instance entries/ray {c.instance_entries / rays:5.2f}  worst thread {c.worst_thread_nodes} node pairs  ({c.rays_reflect} reflect + {c.rays_secondary_shadow} secondary)" if name.startswith("chain") else
                  f"  {name:14s} rays {rays:8d}  node pairs/ray {c.node_fetches / rays:6.1f}  triangles/ray {c.triangle_tests / rays:5.2f}  instance entries/ray {c.instance_entries / rays:5.2f}  worst thread {c.worst_thread_nodes} node pairs")
    ctx.close()
