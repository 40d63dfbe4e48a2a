import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SMALL_DETAIL = (4, 3, 4, 6)      # lucy / helmet / stress-blob subdivisions, stress instances: ~6.5k triangles


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu under gpurun)")


def _have_libs():
    import vulkanhybridraytracer_b200 as vk
    return os.path.exists(os.path.join(vk.lib_dir(), "libvkp.so")) and os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so"))


@pytest.fixture(scope="session", autouse=True)
def _built():
    if not _have_libs():
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="module")
def small_scene(_built):
    import vulkanhybridraytracer_b200 as vk
    sc = vk.HostScene(0, detail=SMALL_DETAIL)
    yield sc
    sc.close()


@pytest.fixture(scope="module")
def small_oracle(small_scene):
    from oracle.oracle_py import Oracle
    o = Oracle(small_scene)
    yield o
    o.close()


@pytest.fixture(scope="module")
def gpu_ctx(small_scene):
    import vulkanhybridraytracer_b200 as vk
    ctx = vk.Context(0)
    ctx.upload_scene(small_scene)
    yield ctx
    ctx.close()
