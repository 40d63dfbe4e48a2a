"""vulkanhybridraytracer_b200 -- B200-native hot path of VkPinut's Hybrid pipeline.

The product is two native libraries built in-tree by ``make`` (see ``__graft_entry__.build``):

* ``lib/libvkp.so``       hand-written sm_100a CUDA kernels behind the C-ABI of ``include/vkp.h``
* ``lib/libvkp_scene.so`` C++ host scene model of the reference (Scene / Prefab / Mesh / Material / Camera, asset readers); no CUDA
* ``lib/libvkp_host.so``  C++ host mirror of the reference's Renderer / VulkanEngine over libvkp

This Python package is only the ctypes harness the tests and ``bench.py`` use to call that C-ABI;
there is no Python or CPU compute path.  Loading fails loudly when the libraries are missing.
"""
from .capi import (  # noqa: F401
    Context,
    Group,
    HostEngine,
    HostScene,
    VkpError,
    load_host,
    load_scene,
    load_vkp,
    lib_dir,
    MODE_DEFERRED,
    MODE_HYBRID,
    MODE_RAYTRACING,
    PLANE,
    STAGE,
    make_params,
)
from . import types  # noqa: F401

__all__ = [
    "Context", "Group", "HostEngine", "HostScene", "VkpError", "load_host", "load_scene", "load_vkp", "lib_dir",
    "MODE_DEFERRED", "MODE_HYBRID", "MODE_RAYTRACING", "PLANE", "STAGE", "make_params", "types",
]
