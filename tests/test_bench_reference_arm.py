"""bench.py --impl reference under torchrun (the way the driver launches it at N > 1): rank 0 alone runs the CPU oracle on ALL host threads
(torchrun exports OMP_NUM_THREADS=1; round 1's arm inherited it and reported 1 core), the other ranks exit 0 without work, and the line
carries the same `config` as the GPU arm.  No GPU, no CUDA library mapped."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_under_torchrun_uses_every_host_thread():
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29611", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]                      # exactly one JSON line, from rank 0
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["higher_is_better"] is True and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["cpu_baseline"]["cores"] == (os.cpu_count() or 1)     # not the 1 thread torchrun's OMP_NUM_THREADS would give
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    sys.path.insert(0, ROOT)
    import bench
    import vulkanhybridraytracer_b200 as vk
    sc = vk.HostScene(0)
    assert d["config"] == bench.workload_config("cfg2", sc)         # the GPU arm prints exactly this dict
    sc.close()
    assert "libvkp.so" not in r.stderr                             # (and the arm never needs the CUDA library)
