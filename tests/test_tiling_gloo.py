"""CPU tests of the multi-GPU partition: tile index math and a world_size-2 gloo composite with the oracle as the
per-rank "device" (the N-GPU image must be bit-identical to the 1-GPU image)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

import vulkanhybridraytracer_b200 as vk
from vulkanhybridraytracer_b200 import tiling


@pytest.mark.parametrize("w,h,world,tw,th", [(64, 36, 2, 16, 8), (100, 60, 3, 16, 8), (1920, 1080, 8, 32, 16), (17, 9, 4, 16, 8)])
def test_pack_unpack_roundtrip(w, h, world, tw, th):
    rng = np.random.default_rng(0)
    img = rng.integers(0, 255, size=(h, w, 4), dtype=np.uint8)
    out = np.zeros_like(img)
    own = tiling.owner_map(w, h, world, tw, th)
    total = 0
    for r in range(world):
        p = tiling.pack_tiles(img, r, world, tw, th)
        assert p.shape[0] == tiling.own_tile_count(w, h, r, world, tw, th)
        assert p.nbytes == tiling.packed_bytes(w, h, r, world, 4, tw, th)
        # a rank's pack only contains its own pixels
        mine = np.zeros_like(img); tiling.unpack_tiles(mine, p, r, world, tw, th)
        assert np.array_equal(mine[own == r], img[own == r]) and not mine[own != r].any()
        tiling.unpack_tiles(out, p, r, world, tw, th)
        total += p.shape[0]
    tx, ty = tiling.tile_grid(w, h, tw, th)
    assert total == tx * ty and np.array_equal(out, img)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import vulkanhybridraytracer_b200 as vk
    from vulkanhybridraytracer_b200 import tiling
    from oracle.oracle_py import Oracle
    W, H = 80, 40
    sc = vk.HostScene(0, detail=(3, 2, 3, 4))
    o = Oracle(sc, threads=2)
    cam = sc.camera(W, H); p = vk.make_params(W, H, frame=1)
    full = o.hybrid_frame(cam, p)["final"]
    own = tiling.owner_map(W, H, world)
    mine = np.where((own == rank)[..., None], full, np.float32(0))          # a rank only "has" its own tiles
    packed = tiling.pack_tiles(mine, rank, world)
    sizes = [tiling.packed_bytes(W, H, r, world, 16) for r in range(world)]
    t = torch.from_numpy(packed.reshape(-1).view(np.uint8).copy())
    got = tiling.gather_packed(t, rank, world, sizes)
    if rank == 0:
        out = np.zeros_like(full)
        for r in range(world):
            n = tiling.own_tile_count(W, H, r, world)
            tiling.unpack_tiles(out, got[r].numpy().view(np.float32).reshape(n, 8, 16, 4), r, world)
        q.put(bool(np.array_equal(out.view(np.uint32), full.view(np.uint32))))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_composite_is_bit_identical():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok


def _worker_reset_frame(rank, world, port, q, reproject_fix):
    """The one real exchange step of the path: on a reset frame blur.comp reads the 3x3 neighbourhood of the
    shadow plane and the re-projected history, so the ranks all-gather their SHADOW tiles and their previous DENOISED tiles
    before the filter.  Emulated with the oracle as each rank's device: a rank only ever holds its own tiles of either plane."""
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import vulkanhybridraytracer_b200 as vk
    from vulkanhybridraytracer_b200 import tiling
    from oracle.oracle_py import Oracle
    W, H = 80, 40
    sc = vk.HostScene(0, detail=(3, 2, 3, 4))
    o = Oracle(sc, threads=2)
    L = len(sc.lights)
    cam0 = sc.camera(W, H)
    hist = None
    for f in range(3):
        hist = o.hybrid_frame(cam0, vk.make_params(W, H, frame=f), history=hist)["denoised"]
    sc.move_camera(3, 6.0); sc.rotate_camera(12.0, 4.0)
    cam1 = sc.camera(W, H, prev=cam0)
    p = vk.make_params(W, H, frame=0, reproject_fix=reproject_fix)
    g = o.gbuffer(cam1, p)
    shadow = o.shadow(cam1, p, g)
    want = o.denoise(p, g, shadow, hist)                                   # the single-device answer
    own = tiling.owner_map(W, H, world) == rank

    def exchange(planes):                                                  # planes: (L, H, W) u8 of which this rank holds its own tiles
        mine = np.where(own[None], planes, 0).astype(np.uint8)
        packed = tiling.pack_tiles(np.ascontiguousarray(mine.transpose(1, 2, 0)), rank, world)          # [tiles, 8, 16, L]
        sizes = [tiling.packed_bytes(W, H, r, world, L) for r in range(world)]
        pad = max(sizes)
        send = torch.zeros(pad, dtype=torch.uint8); send[:packed.size] = torch.from_numpy(packed.reshape(-1).copy())
        bufs = [torch.empty(pad, dtype=torch.uint8) for _ in range(world)]
        dist.all_gather(bufs, send)
        full = np.zeros((H, W, L), np.uint8)
        for r in range(world):
            n = tiling.own_tile_count(W, H, r, world)
            tiling.unpack_tiles(full, bufs[r][:sizes[r]].numpy().reshape(n, 8, 16, L), r, world)
        return np.ascontiguousarray(full.transpose(2, 0, 1))
    shadow_all, hist_all = exchange(shadow), exchange(hist)
    ok = bool(np.array_equal(shadow_all, shadow) and np.array_equal(hist_all, hist))
    den = o.denoise(p, g, shadow_all, hist_all)
    ok = ok and bool(np.array_equal(den[:, own], want[:, own]))            # this rank's tiles of the filtered planes
    # without the exchange the tile borders (3x3 reads) and the re-projected fetches would differ: the test has teeth
    lonely = o.denoise(p, g, np.where(own[None], shadow, 0).astype(np.uint8), np.where(own[None], hist, 0).astype(np.uint8))
    ok = ok and not np.array_equal(lonely[:, own], want[:, own])
    flags = [None] * world
    dist.all_gather_object(flags, ok)
    if rank == 0:
        q.put(all(flags))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("reproject_fix", [0, 1])
def test_two_rank_reset_frame_exchange_matches_one_device(reproject_fix):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + reproject_fix
    procs = [ctx.Process(target=_worker_reset_frame, args=(r, 2, port, q, reproject_fix)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok
