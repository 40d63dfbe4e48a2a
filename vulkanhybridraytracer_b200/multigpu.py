"""Tile-sharded multi-GPU frame driver: one process per GPU, replicated scene + BVH.

Rank r renders the screen tiles t with ``t % world == r`` (vkp_set_tiling).  Pixels are independent except inside the
shadow denoiser on *reset* frames (``denoise_frame <= 0``, blur.comp:49-80), which reads the 3x3 neighbourhood of the
current shadow planes and the previous history at a reprojected position.  To stay bit-identical to the single-GPU
image those frames exchange data -- the path's only real exchange step besides the final composite:

    G-buffer, shadow (own tiles) -> all-gather SHADOW planes and previous DENOISED planes -> denoise, resolve (own tiles)

Static frames (``denoise_frame > 0``) need no exchange.  The composite gathers the finished tiles of a plane to rank 0.
With ``backend="nccl"`` buffers stay on the device (NVLink); ``backend="gloo"`` stages through the host and exists so the
logic can be tested with several ranks on one GPU.

All compute is in libvkp (C-ABI); this module only sequences stages and collectives.
"""
import numpy as np

from . import tiling
from .capi import PLANE, STAGE, MODE_HYBRID


class TiledRenderer:
    def __init__(self, ctx, rank, world, n_lights, tile=(32, 16), backend="nccl", device=None, group=None, stream_ordered=False):
        import torch
        self.torch = torch
        self.ctx, self.rank, self.world, self.n_lights = ctx, rank, world, n_lights
        self.tw, self.th = tile
        self.backend, self.group = backend, group
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        ctx.set_tiling(rank, world, self.tw, self.th)
        # stream_ordered: the context runs on torch's current stream (vkp_set_stream), so NCCL collectives are ordered
        # after the pack kernel and before the unpack kernels by the stream itself -- no host synchronisation per frame
        self.stream_ordered = stream_ordered and backend == "nccl"
        self.stream = None
        if self.stream_ordered:
            # a dedicated non-default torch stream: the legacy default stream's handle is NULL, which vkp_set_stream reads as "own stream"
            self.stream = torch.cuda.Stream(device=self.device)
            ctx.set_stream(self.stream.cuda_stream)
        self._bufs = {}
        self.collectives = 0
        self.peer_output = False
        self.device_barrier = False

    def enable_peer_output(self, device_barrier=True):
        """Composite of FINAL_SRGB8 without a gather: every rank's resolve kernel stores its pixels straight into rank 0's
        image over NVLink peer memory (CUDA IPC); a frame then only needs a barrier.  With device_barrier that barrier is a
        pair of flags in rank 0's memory written / polled by one-warp kernels (vkp_sync_*): no collective call and no host
        round trip per frame; otherwise a 4-byte NCCL all-reduce.  NCCL backend (set-up only), one GPU per rank."""
        import torch.distributed as dist
        if self.world == 1:
            return False
        # two buffers: frame k+1 goes to the other one, so rank 0 may still be consuming frame k when the peers start writing
        box = [None, None]
        if self.rank == 0:
            for which in (0, 1):
                self.ctx.set_output_buffer(which)
                box[which] = self.ctx.ipc_export_plane(PLANE.FINAL_SRGB8)
            self.ctx.set_output_buffer(0)
        dist.broadcast_object_list(box, src=0, group=self.group)
        if self.rank != 0:
            self.ctx.set_output_redirect(self.ctx.ipc_open(box[0]), self.ctx.ipc_open(box[1]))
        self.device_barrier = bool(device_barrier) and self.world <= 64
        if self.device_barrier:
            hs = [self.ctx.sync_export() if self.rank == 0 else None]
            dist.broadcast_object_list(hs, src=0, group=self.group)
            if self.rank != 0:
                self.ctx.sync_attach(hs[0])
        self._frame_no = 0
        self._released = 0
        self._flag = self.torch.zeros(1, dtype=self.torch.int32, device=self.device if self.backend == "nccl" else "cpu")
        self.peer_output = True
        return True

    # ---- helpers -------------------------------------------------------------------------------------------
    def _sizes(self, plane):
        return [self.ctx.tiles_packed_bytes(plane, r, self.world) for r in range(self.world)]

    def _buf(self, key, nbytes):
        b = self._bufs.get(key)
        if b is None or b.numel() < nbytes:
            b = self.torch.empty(max(nbytes, 1), dtype=self.torch.uint8, device=self.device)
            self._bufs[key] = b
        return b

    def _all_gather_plane(self, plane):
        """Every rank ends up with every rank's tiles of `plane` (used for the reset-frame exchange)."""
        import torch.distributed as dist
        if self.world == 1:
            return
        sizes = self._sizes(plane)
        pad = max(sizes)
        send = self._buf(("ag_send", plane), pad)
        self.ctx.tiles_pack(plane, send.data_ptr(), sizes[self.rank])
        if not self.stream_ordered:
            self.ctx.sync()
        recv = [self._buf(("ag_recv", plane, r), pad) for r in range(self.world)]
        if self.backend == "gloo":
            hs = send.cpu(); hr = [self.torch.empty(pad, dtype=self.torch.uint8) for _ in range(self.world)]
            dist.all_gather(hr, hs, group=self.group)
            for r in range(self.world):
                recv[r].copy_(hr[r])
        elif self.stream_ordered:
            with self.torch.cuda.stream(self.stream):
                dist.all_gather(recv, send, group=self.group)
        else:
            dist.all_gather(recv, send, group=self.group)
        if not self.stream_ordered:
            self.torch.cuda.synchronize(self.device)
        self.collectives += 1
        for r in range(self.world):
            if r != self.rank:
                self.ctx.tiles_unpack(plane, r, self.world, recv[r].data_ptr(), sizes[r])

    # ---- frame -----------------------------------------------------------------------------------------------
    def render(self, params):
        """One frame of params.mode on this rank's tiles; exact w.r.t. the single-GPU image."""
        ctx = self.ctx
        ctx.set_params(params)
        if self.peer_output:
            self._frame_no += 1
            ctx.set_output_buffer(self._frame_no & 1)
            if self.device_barrier and self.rank != 0 and self._frame_no > 2:
                ctx.sync_wait_release(self._frame_no - 2)      # rank 0 is done with the buffer this frame's pixels go to
        if self.world == 1 or params.mode != MODE_HYBRID or params.denoise_frame > 0:
            ctx.render()
            return
        # reset frame: the denoiser looks across tile borders and at reprojected history
        ctx.run_stage(STAGE.GBUFFER)       # (the motion plane is only read at own pixels)
        ctx.run_stage(STAGE.SHADOW)
        for l in range(self.n_lights):
            self._all_gather_plane(PLANE.SHADOW0 + l)
            self._all_gather_plane(PLANE.DENOISED0 + l)
        ctx.run_stage(STAGE.DENOISE)
        ctx.run_stage(STAGE.RESOLVE)

    def release(self, after_present=False):
        """Rank 0, device barrier: the frame composited last has been consumed (after_present: once its queued present copy
        has landed); the peers may overwrite its buffer two frames later."""
        if self.peer_output and self.device_barrier and self.rank == 0 and self._released != self._frame_no:
            self.ctx.sync_release(self._frame_no, after_present)
            self._released = self._frame_no
            return 1
        return 0

    def composite(self, plane=PLANE.FINAL_SRGB8, release=True):
        """Gather the finished tiles of `plane` to rank 0 (the one collective of a static frame).  Returns the number of
        kernels launched on this rank.  release=False (device barrier): the caller presents the frame first and then calls
        release(after_present=True)."""
        import torch.distributed as dist
        if self.world == 1:
            return 0
        if self.peer_output and plane == PLANE.FINAL_SRGB8 and self.device_barrier:
            if self.rank == 0:
                self.ctx.sync_wait_arrivals(self._frame_no)
                return 1 + (self.release() if release else 0)
            self.ctx.sync_arrive(self._frame_no)
            return 1
        if self.peer_output and plane == PLANE.FINAL_SRGB8:
            # the pixels are already in rank 0's image; a 4-byte all-reduce is the completion barrier (stream-ordered)
            if self.stream_ordered:
                with self.torch.cuda.stream(self.stream):
                    dist.all_reduce(self._flag, group=self.group)
            else:
                self.ctx.sync()
                dist.all_reduce(self._flag, group=self.group)
                self.torch.cuda.synchronize(self.device)
            self.collectives += 1
            return 0
        sizes = self._sizes(plane)
        pad = max(sizes)
        send = self._buf(("c_send", plane), pad)
        self.ctx.tiles_pack(plane, send.data_ptr(), sizes[self.rank])
        if not self.stream_ordered:
            self.ctx.sync()
        launches = 1
        if self.backend == "gloo":
            hs = send.cpu()
            hr = [self.torch.empty(pad, dtype=self.torch.uint8) for _ in range(self.world)] if self.rank == 0 else None
            dist.gather(hs, hr, dst=0, group=self.group)
            if self.rank == 0:
                recv = [self._buf(("c_recv", plane, r), pad) for r in range(self.world)]
                for r in range(1, self.world):
                    recv[r].copy_(hr[r])
        else:
            recv = [self._buf(("c_recv", plane, r), pad) for r in range(self.world)] if self.rank == 0 else None
            if self.stream_ordered:
                with self.torch.cuda.stream(self.stream):
                    dist.gather(send, recv, dst=0, group=self.group)
            else:
                dist.gather(send, recv, dst=0, group=self.group)
        self.collectives += 1
        if self.rank == 0:
            if not self.stream_ordered:
                self.torch.cuda.synchronize(self.device)
            for r in range(1, self.world):
                self.ctx.tiles_unpack(plane, r, self.world, recv[r].data_ptr(), sizes[r])
                launches += 1
        return launches
