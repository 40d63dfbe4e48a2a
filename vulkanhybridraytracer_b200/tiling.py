"""Host-side logic of the multi-GPU frame partition.

A frame is cut into tile_w x tile_h screen tiles numbered row-major; tile t belongs to rank ``t % world``.  Every
rank renders its own tiles over a replicated BVH (vkp_set_tiling), packs them densely on the device
(vkp_tiles_pack), the packed buffers are gathered to rank 0 over NCCL -- the only collective of the path -- and rank
0 scatters them into the full frame (vkp_tiles_unpack).  The numpy functions here restate the device index math so
the partition can be tested without a GPU (gloo, world_size 2) and are what the device kernels are checked against.
"""
import numpy as np


def tile_grid(width, height, tile_w=16, tile_h=8):
    return (width + tile_w - 1) // tile_w, (height + tile_h - 1) // tile_h


def own_tile_count(width, height, rank, world, tile_w=16, tile_h=8):
    tx, ty = tile_grid(width, height, tile_w, tile_h)
    total = tx * ty
    return (total - rank + world - 1) // world if total > rank else 0


def owner_map(width, height, world, tile_w=16, tile_h=8):
    """(H, W) int array: which rank renders each pixel."""
    tx, _ = tile_grid(width, height, tile_w, tile_h)
    ys, xs = np.mgrid[0:height, 0:width]
    return ((ys // tile_h) * tx + (xs // tile_w)) % world


def packed_bytes(width, height, rank, world, elem_bytes, tile_w=16, tile_h=8):
    return own_tile_count(width, height, rank, world, tile_w, tile_h) * tile_w * tile_h * elem_bytes


def pack_tiles(plane, rank, world, tile_w=16, tile_h=8):
    """Dense [own_tiles, tile_h, tile_w, ...] copy of the tiles of `rank` (partial edge tiles are zero padded)."""
    h, w = plane.shape[:2]
    tx, _ = tile_grid(w, h, tile_w, tile_h)
    n = own_tile_count(w, h, rank, world, tile_w, tile_h)
    out = np.zeros((n, tile_h, tile_w) + plane.shape[2:], plane.dtype)
    for k in range(n):
        t = rank + k * world
        y0, x0 = (t // tx) * tile_h, (t % tx) * tile_w
        blk = plane[y0:y0 + tile_h, x0:x0 + tile_w]
        out[k, :blk.shape[0], :blk.shape[1]] = blk
    return out


def unpack_tiles(plane, packed, rank, world, tile_w=16, tile_h=8):
    """Inverse of pack_tiles: scatter the tiles of `rank` into `plane` (in place)."""
    h, w = plane.shape[:2]
    tx, _ = tile_grid(w, h, tile_w, tile_h)
    for k in range(packed.shape[0]):
        t = rank + k * world
        y0, x0 = (t // tx) * tile_h, (t % tx) * tile_w
        hh, ww = min(tile_h, h - y0), min(tile_w, w - x0)
        plane[y0:y0 + hh, x0:x0 + ww] = packed[k, :hh, :ww]
    return plane


def gather_packed(packed_tensor, rank, world, sizes, group=None):
    """The one collective: gather every rank's packed tile buffer (a torch uint8 tensor on the rank's device) to rank 0.
    `sizes[r]` is the byte size of rank r's buffer.  Works on NCCL (device tensors, NVLink) and gloo (CPU tensors)."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return [packed_tensor]
    pad = max(sizes)
    send = packed_tensor
    if send.numel() < pad:
        send = torch.zeros(pad, dtype=torch.uint8, device=packed_tensor.device)
        send[:packed_tensor.numel()] = packed_tensor
    if rank == 0:
        bufs = [torch.empty(pad, dtype=torch.uint8, device=packed_tensor.device) for _ in range(world)]
        dist.gather(send, bufs, dst=0, group=group)
        return [b[:sizes[r]] for r, b in enumerate(bufs)]
    dist.gather(send, None, dst=0, group=group)
    return None
