# Builds everything in-tree (the .so files travel to the GPU box with the gpurun snapshot):
#   vulkanhybridraytracer_b200/lib/libvkp.so       CUDA kernels + C-ABI (include/vkp.h), sm_100a only
#   vulkanhybridraytracer_b200/lib/libvkp_scene.so host scene model (Scene / Prefab / Mesh / Material / Camera, OBJ / glTF / image readers): no CUDA
#   vulkanhybridraytracer_b200/lib/libvkp_host.so  host Renderer / VulkanEngine over libvkp + C shim
#   vulkanhybridraytracer_b200/bin/vkpinut         headless CLI (replaces src/main.cpp + window.cpp)
#   oracle/liboracle*.so                           CPU oracle (test infrastructure)
ROOT   := $(dir $(abspath $(lastword $(MAKEFILE_LIST))))
PKG    := $(ROOT)vulkanhybridraytracer_b200
CSRC   := $(PKG)/csrc
HOST   := $(PKG)/host
LIBDIR := $(PKG)/lib
BINDIR := $(PKG)/bin
OBJDIR := $(ROOT)build

NVCC   ?= /usr/local/cuda/bin/nvcc
CXX    := $(shell if [ -x /usr/bin/g++ ]; then echo /usr/bin/g++; else echo g++; fi)
ARCH   := -gencode arch=compute_100a,code=sm_100a
# -fmad=false: the numerics contract (bit-identical to the scalar fp32 evaluation); explicit __fmaf_rn where safe
NVFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC,-fvisibility=hidden -ccbin $(CXX) --cudart static \
           -Xptxas -v -diag-suppress 20012
# -fwrapv: the integer IDCT of the JPEG reader may overflow on corrupt streams (as in stb_image); wrapping keeps that defined
CXXFLAGS := -O2 -std=c++17 -fPIC -Wall -Wno-unused-variable -fwrapv

CU_SRCS := $(CSRC)/vkp_api.cu $(CSRC)/bvh_build.cu $(CSRC)/kernels_frame.cu $(CSRC)/vkp_group.cu
CU_OBJS := $(patsubst $(CSRC)/%.cu,$(OBJDIR)/%.o,$(CU_SRCS))
CU_HDRS := $(wildcard $(CSRC)/*.cuh) $(wildcard $(CSRC)/*.h) $(wildcard $(ROOT)include/*.h)
SCENE_SRCS := $(HOST)/scene_model.cpp $(HOST)/procedural.cpp $(HOST)/output.cpp $(HOST)/image_decode.cpp $(HOST)/gltf.cpp $(HOST)/scene_capi.cpp
HOST_SRCS := $(HOST)/renderer.cpp $(HOST)/host_capi.cpp
HOST_HDRS := $(wildcard $(HOST)/*.h) $(wildcard $(ROOT)include/*.h)

all: libvkp host oracle stress
libvkp: $(LIBDIR)/libvkp.so
# test artefact: the same library with a 2-level shared-memory traversal stack and a 48-item warp queue, so that the local-memory
# overflow of the stack and the serial fall-back of the warp-cooperative traversal are exercised by tests/ on small scenes
stress: $(LIBDIR)/libvkp_stress.so
host: $(LIBDIR)/libvkp_scene.so $(LIBDIR)/libvkp_host.so $(BINDIR)/vkpinut
oracle:
	$(MAKE) -C $(ROOT)oracle

$(OBJDIR)/%.o: $(CSRC)/%.cu $(CU_HDRS)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(OBJDIR)/$*.ptxas.log || (cat $(OBJDIR)/$*.ptxas.log; exit 1)

STRESS_OBJS := $(patsubst $(CSRC)/%.cu,$(OBJDIR)/stress_%.o,$(CU_SRCS))
$(OBJDIR)/stress_%.o: $(CSRC)/%.cu $(CU_HDRS)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -DVKP_SSTACK=2 -DVKP_WQ_CAP=48 -c $< -o $@ 2> $(OBJDIR)/stress_$*.ptxas.log || (cat $(OBJDIR)/stress_$*.ptxas.log; exit 1)
$(LIBDIR)/libvkp_stress.so: $(STRESS_OBJS)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(ARCH) -shared --cudart static -ccbin $(CXX) -o $@ $(STRESS_OBJS) -lpthread -ldl -lrt

$(LIBDIR)/libvkp.so: $(CU_OBJS)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(ARCH) -shared --cudart static -ccbin $(CXX) -o $@ $(CU_OBJS) -lpthread -ldl -lrt

$(LIBDIR)/libvkp_scene.so: $(SCENE_SRCS) $(HOST_HDRS)
	@mkdir -p $(LIBDIR)
	$(CXX) $(CXXFLAGS) -shared -o $@ $(SCENE_SRCS) -lz

$(LIBDIR)/libvkp_host.so: $(HOST_SRCS) $(HOST_HDRS) $(LIBDIR)/libvkp.so $(LIBDIR)/libvkp_scene.so
	$(CXX) $(CXXFLAGS) -shared -o $@ $(HOST_SRCS) -L$(LIBDIR) -lvkp_scene -lvkp -Wl,-rpath,'$$ORIGIN'

$(BINDIR)/vkpinut: $(HOST)/main.cpp $(LIBDIR)/libvkp_host.so
	@mkdir -p $(BINDIR)
	$(CXX) $(CXXFLAGS) -o $@ $(HOST)/main.cpp -L$(LIBDIR) -lvkp_host -lvkp_scene -lvkp -Wl,-rpath,'$$ORIGIN/../lib'

sass: $(LIBDIR)/libvkp.so
	@mkdir -p $(ROOT)profiles
	/usr/local/cuda/bin/cuobjdump -sass $(LIBDIR)/libvkp.so > $(ROOT)profiles/libvkp.sass

clean:
	rm -rf $(OBJDIR) $(LIBDIR) $(BINDIR)
	$(MAKE) -C $(ROOT)oracle clean

.PHONY: all libvkp host oracle stress clean sass
