// sequential_main.cpp -- GPU-backed counterpart of /root/reference/worker/sequential/main.go:
//   sequential_gpu <scene.json> <width> <height> [out.ppm] [frames] [exact|fast]
// Loads the environment (:43), renders the screen with draw() (:15-32, one GPU region call instead of
// the per-pixel loop) and, instead of presenting an SDL surface, writes it as a binary PPM.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "gpu_worker.hpp"

int main(int argc, char **argv) {
    if (argc < 4) {
        std::fprintf(stderr, "Improper parameters.  This program requires the parameters:\n"
                             "\t(1) scene file path\n\t(2) screen width\n\t(3) screen height\n\t[4] output .ppm\n\t[5] frames\n\t[6] exact|fast\n");
        return 1;
    }
    long width = std::atol(argv[2]), height = std::atol(argv[3]);
    if (width <= 0 || height <= 0) { std::fprintf(stderr, "Invalid screen width/height.\n"); return 1; }
    const char *out = argc > 4 ? argv[4] : nullptr;
    int frames = argc > 5 ? std::atoi(argv[5]) : 1;
    uint32_t precision = (argc > 6 && !std::strcmp(argv[6], "exact")) ? RT_FP64_EXACT : RT_FP32_FAST;
    try {
        state::Environment env = state::EnvironmentFromFile(argv[1]);
        worker::Tracer tracer(env, uint32_t(width), uint32_t(height), 0, precision);
        if (!tracer.ok()) { std::fprintf(stderr, "Could not create the GPU scene: %s\n", tracer.error().c_str()); return 2; }
        std::vector<uint8_t> surface;
        std::string err;
        state::EnvMutables &scene = env.Mutable();
        for (int f = 0; f < frames; f++) {
            auto t0 = std::chrono::steady_clock::now();
            if (!worker::draw(tracer, scene, surface, err)) { std::fprintf(stderr, "draw failed: %s\n", err.c_str()); return 3; }
            double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            std::printf("frame %d: %.3f ms\n", f, ms);
        }
        if (out) {
            FILE *fp = std::fopen(out, "wb");
            if (!fp) { std::fprintf(stderr, "cannot write %s\n", out); return 4; }
            std::fprintf(fp, "P6\n%ld %ld\n255\n", width, height);
            std::fwrite(surface.data(), 1, surface.size(), fp);
            std::fclose(fp);
        }
    } catch (const std::exception &e) {
        std::fprintf(stderr, "Could not load environment: %s\n", e.what());
        return 1;
    }
    return 0;
}
