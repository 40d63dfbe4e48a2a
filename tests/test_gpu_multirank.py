"""GPU test of the tile-sharded multi-rank driver: 2 and 3 ranks sharing ONE GPU (gloo, host-staged collectives) must
composite to the bit-identical single-context image, including reset frames and a camera move."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, H = 250, 140
DETAIL = (4, 3, 4, 6)


def frame_sequence(vk, sc):
    """(camera, params) for: 3 static frames, then the camera moves (reset) and 2 more frames."""
    cam0 = sc.camera(W, H)
    seq = [(cam0, vk.make_params(W, H, frame=f)) for f in range(3)]
    sc.move_camera(3, 6.0); sc.rotate_camera(25.0, -10.0)
    cam1 = sc.camera(W, H, prev=cam0)
    seq.append((cam1, vk.make_params(W, H, frame=0)))
    cam2 = sc.camera(W, H)                       # static again
    seq.append((cam2, vk.make_params(W, H, frame=1)))
    return seq


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import vulkanhybridraytracer_b200 as vk
    from vulkanhybridraytracer_b200 import PLANE
    from vulkanhybridraytracer_b200.multigpu import TiledRenderer
    torch.cuda.set_device(0)
    sc = vk.HostScene(0, detail=DETAIL)
    ctx = vk.Context(0); ctx.upload_scene(sc)
    tr = TiledRenderer(ctx, rank, world, len(sc.lights), tile=(16, 8), backend="gloo", device=torch.device("cuda", 0))
    outs = []
    ctx.set_params(vk.make_params(W, H))
    assert tr.enable_peer_output()          # sRGB8 pixels go straight into rank 0's plane through CUDA IPC peer memory
    for cam, p in frame_sequence(vk, sc):
        ctx.set_camera(cam)
        tr.render(p)
        tr.composite(PLANE.FINAL_F32)       # pack + gather + unpack
        tr.composite(PLANE.DENOISED0 + 1)
        tr.composite(PLANE.FINAL_SRGB8)     # barrier only
        if rank == 0:
            outs.append((ctx.readback(PLANE.FINAL_F32).copy(), ctx.readback(PLANE.DENOISED0 + 1).copy(), ctx.readback(PLANE.FINAL_SRGB8).copy()))
    if rank == 0:
        q.put(outs)
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_multirank_frames_are_bit_identical_to_one_gpu(world):
    import torch.multiprocessing as mp
    import vulkanhybridraytracer_b200 as vk
    from vulkanhybridraytracer_b200 import PLANE
    sc = vk.HostScene(0, detail=DETAIL)
    ctx = vk.Context(0); ctx.upload_scene(sc)
    ref = []
    for cam, p in frame_sequence(vk, sc):
        ctx.set_camera(cam); ctx.set_params(p); ctx.render()
        ref.append((ctx.readback(PLANE.FINAL_F32).copy(), ctx.readback(PLANE.DENOISED0 + 1).copy(), ctx.readback(PLANE.FINAL_SRGB8).copy()))
    ctx.close(); sc.close()
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    port = 29600 + (os.getpid() % 1500) + world
    procs = [mpc.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert len(outs) == len(ref)
    for k, ((f, d, g), (rf, rd, rg)) in enumerate(zip(outs, ref)):
        assert np.array_equal(d, rd), f"denoised plane differs on frame {k}"
        assert np.array_equal(f.view(np.uint32), rf.view(np.uint32)), f"final image differs on frame {k}"
        assert np.array_equal(g, rg), f"peer-written sRGB8 image differs on frame {k}"
