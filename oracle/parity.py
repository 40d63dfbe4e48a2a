"""TEST INFRASTRUCTURE: the CUDA path (through the C-ABI) and the CPU oracle side by side on the SAME frames.

Used by tests/ (full-size parity tests) and by bench.py (the `parity` block of the JSON line and the `cpu_baseline` timing, which
come from the very same oracle frames).  The product never imports this module.

Two ways of running a workload:

* whole frames (stride 1): a frame sequence 0 .. F-1 starting from a cleared history -- frame 0 is the reset frame of the shadow
  filter (3x3 clamp + reprojection, blur.comp:49-80), frames >= 1 accumulate (blur.comp:41-48); every plane of every frame is compared;
* a row sample (stride > 1) for the workloads whose full frame would take the oracle minutes: ONE frame whose stages are all
  per-pixel -- Hybrid: frame index 1 over a cleared history (accumulation branch of the filter); Raytraced: frame 0 -- so that the
  rows y % stride == offset of the oracle are exactly the pixels the full GPU frame holds there.

The bar is north_star's: closest-hit ids agree on >= 99.9 % of the rays, linear radiance MAE <= 1e-3 and max <= 1e-2.  Bit
equality of every plane is reported beside it (it is what the numerics contract is built to give).
"""
import time

import numpy as np

BAR = {"hit_id_agreement_min": 0.999, "mae_max": 1e-3, "max_abs_max": 1e-2}


def _bits(a):
    a = np.ascontiguousarray(a)
    if a.dtype == np.float16:
        return a.view(np.uint16)
    if a.dtype == np.float32:
        return a.view(np.uint32)
    return a


def _radiance_error(got, want):
    """|got - want| over RGB with identical bit patterns (incl. NaN / inf the reference produces) counted as 0."""
    g = np.ascontiguousarray(got[..., :3], dtype=np.float32); w = np.ascontiguousarray(want[..., :3], dtype=np.float32)
    same = g.view(np.uint32) == w.view(np.uint32)
    with np.errstate(invalid="ignore", over="ignore"):
        d = np.abs(g.astype(np.float64) - w.astype(np.float64))
    d = np.where(same, 0.0, d)
    d = np.where(np.isfinite(d), d, np.inf)
    return d


class Tally:
    """Accumulates the comparison of several frames / row samples."""

    def __init__(self):
        self.pixels = 0; self.rows = 0; self.frames = 0
        self.ids_equal = 0; self.ids_pixels = 0; self.abs_sum = 0.0; self.abs_max = 0.0
        self.planes_equal = {}
        self.final_bits_equal = 0

    def add_plane(self, name, got, want):
        eq = bool(np.array_equal(_bits(got), _bits(want)))
        self.planes_equal[name] = self.planes_equal.get(name, True) and eq
        return eq

    def add_ids(self, got, want):
        self.ids_equal += int((got == want).all(-1).sum()); self.ids_pixels += int(got.shape[0] * got.shape[1])

    def add_final(self, got, want):
        d = _radiance_error(got, want)
        self.abs_sum += float(d.sum()); self.abs_max = max(self.abs_max, float(d.max()) if d.size else 0.0)
        self.final_bits_equal += int((_bits(got[..., :3]) == _bits(want[..., :3])).all(-1).sum())

    def result(self, extra=None):
        n = max(self.pixels, 1)
        agree = self.ids_equal / self.ids_pixels if self.ids_pixels else None      # Raytraced mode stores no id plane
        mae = self.abs_sum / (3 * n)
        out = {"pixels_compared": self.pixels, "rows_compared": self.rows, "frames_compared": self.frames,
               "hit_id_agreement": agree, "mae": mae, "max": self.abs_max,
               "final_bits_equal_fraction": self.final_bits_equal / n,
               "planes_byte_equal": dict(self.planes_equal),
               "shadow_bytes_equal": all(v for k, v in self.planes_equal.items() if k.startswith("shadow")) if any(k.startswith("shadow") for k in self.planes_equal) else None,
               "bar": BAR}
        ok = mae <= BAR["mae_max"] and self.abs_max <= BAR["max_abs_max"] and (agree is None or agree >= BAR["hit_id_agreement_min"])
        out["pass"] = bool(ok)
        if extra:
            out.update(extra)
        return out


def run_pair(vk, ctx, oracle, scene, W, H, mode, spp=1, samples=1, faithful=1, frames=3, stride=1, chain=2, tally=None, time_budget_s=None,
             max_bounces=8):
    """Renders the same frames with the oracle and with the context `ctx` (scene already uploaded, 1 GPU, no tiling) and compares them.

    Returns (parity dict, cpu dict).  cpu = {"seconds", "rays" (shadow + reflection + secondary shadow), "primary", "frames", "rows"}:
    the wall time the oracle took for exactly the rows that were compared -- bench.py's cpu_baseline.
    time_budget_s: whole-frame mode only -- stop adding frames once the oracle has used that much wall time (at least 2 frames)."""
    from vulkanhybridraytracer_b200 import PLANE, MODE_HYBRID
    cam = scene.camera(W, H)
    L = len(scene.lights)
    t = tally or Tally()
    offset = stride // 2
    rows = np.arange(offset, H, stride) if stride > 1 else np.arange(H)
    oracle.set_row_sampling(stride, offset)
    ctx.set_camera(cam)
    cpu = {"seconds": 0.0, "rays": 0, "primary": 0, "frames": 0, "rows": int(len(rows))}
    try:
        if mode == MODE_HYBRID:
            seq = list(range(frames)) if stride == 1 else [1]
            ctx.set_params(vk.make_params(W, H, mode=mode, spp=spp, shadow_samples=samples, faithful_rng=faithful, frame=seq[0], wavefront=chain, max_bounces=max_bounces))
            ctx.reset_history()
            hist = None
            for f in seq:
                p = vk.make_params(W, H, mode=mode, spp=spp, shadow_samples=samples, faithful_rng=faithful, frame=f, wavefront=chain, max_bounces=max_bounces)
                oracle.reset_counters()
                t0 = time.perf_counter()
                ref = oracle.hybrid_frame(cam, p, history=hist)
                cpu["seconds"] += time.perf_counter() - t0
                c = list(oracle.counters.rays)
                cpu["rays"] += c[1] + c[2] + c[3]; cpu["primary"] += c[0]; cpu["frames"] += 1
                hist = ref["denoised"]
                ctx.set_params(p); ctx.render()
                g = ref["g"]
                for name, plane, want in (("position", PLANE.POSITION, g.position), ("normal", PLANE.NORMAL, g.normal), ("albedo", PLANE.ALBEDO, g.albedo),
                                          ("motion", PLANE.MOTION, g.motion), ("material", PLANE.MATERIAL, g.material), ("emissive", PLANE.EMISSIVE, g.emissive)):
                    t.add_plane(name, ctx.readback(plane)[rows], want[rows])
                t.add_ids(ctx.readback(PLANE.HIT_IDS)[rows], g.hit_ids[rows])
                for l in range(L):
                    t.add_plane(f"shadow{l}", ctx.readback(PLANE.SHADOW0 + l)[rows], ref["shadow"][l][rows])
                    t.add_plane(f"denoised{l}", ctx.readback(PLANE.DENOISED0 + l)[rows], ref["denoised"][l][rows])
                final = ctx.readback(PLANE.FINAL_F32)
                t.add_final(final[rows], ref["final"][rows])
                t.add_plane("final_srgb8", ctx.readback(PLANE.FINAL_SRGB8)[rows], oracle.post(p, ref["final"])[rows])
                t.pixels += int(len(rows)) * W; t.rows += int(len(rows)); t.frames += 1
                if time_budget_s is not None and stride == 1 and cpu["frames"] >= 2 and cpu["seconds"] >= time_budget_s:
                    break
        else:
            p = vk.make_params(W, H, mode=mode, spp=spp, shadow_samples=samples, faithful_rng=faithful, frame=0, max_bounces=max_bounces)
            ctx.set_params(p); ctx.reset_history()
            oracle.reset_counters()
            t0 = time.perf_counter()
            ref = oracle.raytrace(cam, p)
            cpu["seconds"] += time.perf_counter() - t0
            c = list(oracle.counters.rays)
            cpu["rays"] += c[1] + c[2] + c[3]; cpu["primary"] += c[0]; cpu["frames"] += 1
            ctx.render()
            final = ctx.readback(PLANE.FINAL_F32)
            t.add_final(final[rows], ref[rows])
            t.add_plane("final_srgb8", ctx.readback(PLANE.FINAL_SRGB8)[rows], oracle.post(p, ref)[rows])
            t.pixels += int(len(rows)) * W; t.rows += int(len(rows)); t.frames += 1
    finally:
        oracle.set_row_sampling(1, 0)
    return t.result({"row_stride": stride, "size": [W, H]}), cpu
