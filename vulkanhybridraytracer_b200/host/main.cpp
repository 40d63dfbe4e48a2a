// main.cpp -- headless replacement of src/main.cpp + window.cpp: VulkanEngine engine; init(); run(); cleanup();
//   vkpinut [--scene N] [--mode deferred|raytracing|hybrid] [--width W --height H] [--frames F] [--samples S]
//           [--spp N] [--device D] [--gpus N] [--png out.png] [--exr out.exr] [--prefab model.obj|.gltf|.glb]
//           [--walk DIR:DT[,DIR:DT...]] [--turn DX:DY] [--sequence prefix] [--reproject-fix] [--detail a,b,c,d]
// --walk / --turn replay key presses and mouse motion every frame (DIR = forward|backward|left|right|up|down, DT in ms as in
// Camera::processKeyboard; DX, DY in pixels as in Camera::rotate): a moving-camera sequence.  --sequence writes prefix_%04d.png per frame.
// Assets (sphere.obj, lucy.obj, DamagedHelmet.gltf + textures, the sky .jpg / .hdr) are picked up from the reference's
// search paths or $VKPINUT_ASSETS when present; otherwise the deterministic stand-ins are used.
#include "vkpinut.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <string>
#include <utility>
#include <vector>

int main(int argc, char** argv)
{
    VulkanEngine engine;
    int frames = 4;
    std::string png, exr, sequence;
    std::vector<std::pair<Camera_Movement, float>> walk;
    float turn_x = 0.f, turn_y = 0.f;
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
        else if (a == "--gpus") engine._gpus = std::atoi(next());          // GPUs device .. device + N - 1 share every frame (tile-sharded)
        else if (a == "--prefab") engine._prefabFile = next();
        else if (a == "--detail") {      // tessellation of the procedural stand-ins: lucy,helmet,stress,instances (default 7,5,7,31)
            MeshDetail& d = engine._detail;
            if (std::sscanf(next(), "%d,%d,%d,%d", &d.lucy_subdiv, &d.helmet_subdiv, &d.stress_subdiv, &d.stress_instances) != 4) { std::fprintf(stderr, "--detail wants a,b,c,d\n"); return 2; }
        }
        else if (a == "--sequence") sequence = next();
        else if (a == "--reproject-fix") engine._reprojectFix = true;
        else if (a == "--turn") { if (std::sscanf(next(), "%f:%f", &turn_x, &turn_y) != 2) { std::fprintf(stderr, "--turn wants DX:DY\n"); return 2; } }
        else if (a == "--walk") {
            std::string spec = next();
            size_t pos = 0;
            while (pos < spec.size()) {
                size_t end = spec.find(',', pos);
                std::string item = spec.substr(pos, end == std::string::npos ? std::string::npos : end - pos);
                size_t colon = item.find(':');
                std::string dir = item.substr(0, colon);
                float dt = colon == std::string::npos ? 16.6f : (float)std::atof(item.c_str() + colon + 1);
                static const char* names[] = {"forward", "backward", "left", "right", "up", "down"};
                int k = 0;
                for (; k < 6; k++) if (dir == names[k]) break;
                if (k == 6) { std::fprintf(stderr, "--walk: unknown direction %s\n", dir.c_str()); return 2; }
                walk.push_back({(Camera_Movement)k, dt});
                if (end == std::string::npos) break;
                pos = end + 1;
            }
        }
        else if (a == "--png") png = next();
        else if (a == "--exr") exr = next();
        else if (a == "--mode") {
            std::string m = next();
            engine._mode = m == "deferred" ? DEFERRED : (m == "raytracing" ? RAYTRACING : HYBRID);
        } else { std::fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
    }
    try {
        engine.init();
        if (!walk.empty() || turn_x != 0.f || turn_y != 0.f)
            engine._onFrame = [&](int frame) {
                if (frame == 0) return;                       // the first frame shows the scene's own camera
                for (auto& w : walk) engine._scene->_camera->processKeyboard(w.first, w.second);
                if (turn_x != 0.f || turn_y != 0.f) engine._scene->_camera->rotate(turn_x, turn_y);
            };
        if (sequence.empty()) engine.run(frames);
        else for (int f = 0; f < frames; f++) {
            engine.run(1);
            char name[1024];
            std::snprintf(name, sizeof(name), "%s_%04d.png", sequence.c_str(), f);
            auto img = engine.renderer->read_final_srgb8();
            write_png_rgba8(name, img.data(), engine._width, engine._height);
        }
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
