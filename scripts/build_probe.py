"""BLAS / TLAS build figures of a scene (vkp_get_build_stats): first build of the process and a second one with the kernels loaded."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vulkanhybridraytracer_b200 as vk

for scene_idx in [int(a) for a in (sys.argv[1:] or ["0"])]:
    sc = vk.HostScene(scene_idx)
    print(f"scene {scene_idx}: {sc.n_triangles} instanced triangles, {len(sc.instances)} instances", flush=True)
    for k in range(2):
        ctx = vk.Context(0); ctx.upload_scene(sc)
        for _ in range(3):
            ctx.build_tlas(True)
        ctx.sync()
        b = ctx.build_stats()
        print("  build", k, {k2: (round(v, 4) if isinstance(v, float) else v) for k2, v in b.items()}, flush=True)
        ctx.close()
    sc.close()
