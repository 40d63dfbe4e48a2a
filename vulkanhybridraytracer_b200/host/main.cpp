// main.cpp -- headless replacement of src/main.cpp + window.cpp: VulkanEngine engine; init(); run(); cleanup();
//   vkpinut [--scene N] [--mode deferred|raytracing|hybrid] [--width W --height H] [--frames F] [--samples S]
//           [--spp N] [--device D] [--png out.png] [--exr out.exr] [--prefab model.obj|.gltf|.glb]
// Assets (sphere.obj, lucy.obj, DamagedHelmet.gltf + textures, the sky .jpg / .hdr) are picked up from the reference's
// search paths or $VKPINUT_ASSETS when present; otherwise the deterministic stand-ins are used.
#include "vkpinut.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <string>

int main(int argc, char** argv)
{
    VulkanEngine engine;
    int frames = 4;
    std::string png, exr;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() -> const char* { return i + 1 < argc ? argv[++i] : ""; };
        if (a == "--scene") engine._sceneIndex = std::atoi(next());
        else if (a == "--width") engine._width = std::atoi(next());
        else if (a == "--height") engine._height = std::atoi(next());
        else if (a == "--frames") frames = std::atoi(next());
        else if (a == "--samples") engine._samples = std::atoi(next());
        else if (a == "--spp") engine._spp = std::atoi(next());
        else if (a == "--device") engine._device = std::atoi(next());
        else if (a == "--prefab") engine._prefabFile = next();
        else if (a == "--png") png = next();
        else if (a == "--exr") exr = next();
        else if (a == "--mode") {
            std::string m = next();
            engine._mode = m == "deferred" ? DEFERRED : (m == "raytracing" ? RAYTRACING : HYBRID);
        } else { std::fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
    }
    try {
        engine.init();
        engine.run(frames);
        vkp_counters c = engine.renderer->counters();
        std::printf("frames %d  %dx%d  rays: primary %llu shadow %llu reflect %llu secondary %llu\n", frames, engine._width, engine._height,
                    (unsigned long long)c.rays_primary, (unsigned long long)c.rays_shadow, (unsigned long long)c.rays_reflect,
                    (unsigned long long)c.rays_secondary_shadow);
        if (!png.empty()) { auto img = engine.renderer->read_final_srgb8(); write_png_rgba8(png.c_str(), img.data(), engine._width, engine._height); }
        if (!exr.empty()) { auto img = engine.renderer->read_final_linear(); write_exr_rgba32f(exr.c_str(), img.data(), engine._width, engine._height); }
        engine.cleanup();
    } catch (const std::exception& e) {
        std::fprintf(stderr, "vkpinut: %s\n", e.what());
        return 1;
    }
    return 0;
}
