// mcskin_bench — headless caller of the drop-in API: synthetic skin -> Scene ->
// TileRenderer::render(scene, config) -> Image, timed, optional raw / PPM dump.
//
//   mcskin_bench [--width W] [--height H] [--spp N] [--bounces B] [--seed S] [--pose 0..6]
//                [--legacy] [--frames K] [--ppm out.ppm] [--raw out.f32]
//
// Prints one JSON line with ms/frame (wall clock around render(): flatten + upload + kernels
// + download) for the K frames after one warm-up frame.
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "mcskin/detail/unflatten.hpp"
#include "mcskin_cuda.h"
#include "raytracer/tile_renderer.h"
#include "scene/scene.h"

MCSKIN_DEFINE_UNFLATTEN()

namespace {

// deterministic synthetic atlas of SURVEY.md §8d (same as minecraftskin_raytracer_b200.scene.synth_skin)
uint32_t lowbias32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
std::vector<uint8_t> synthSkin(uint32_t seed, bool legacy) {
    const int h = legacy ? 32 : 64;
    std::vector<uint8_t> img(static_cast<size_t>(h) * 64 * 4);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < 64; ++x) {
            const uint32_t v = lowbias32(seed * 4096u + static_cast<uint32_t>(y) * 64u + x + 1u);
            uint8_t* p = &img[(static_cast<size_t>(y) * 64 + x) * 4];
            p[0] = (v >> 8) & 0xff; p[1] = (v >> 16) & 0xff; p[2] = (v >> 24) & 0xff; p[3] = 255;
            bool outer = x >= 32 && y < 16;
            if (!legacy) outer = outer || (y >= 32 && y < 48) || (y >= 48 && (x < 16 || x >= 48));
            if (outer && (v & 1u)) p[3] = 0;
        }
    return img;
}

const float kPoses[7][12] = {
    {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},          {0, 0, 0, 0, 30, 0, -30, 0, -25, 0, 25, 0},
    {-5, 0, 5, 0, 50, 0, -50, 0, -45, 0, 45, 0},   {5, 0, 0, 0, -140, -20, 0, 0, 0, 0, 0, 0},
    {0, 0, 0, 0, -10, 0, -10, 0, -90, 0, -90, 0},  {-10, 0, 5, 0, -90, 10, 20, -10, -15, 0, 20, 0},
    {30, 15, 0, 5, -45, 30, 150, -10, 0, 0, 0, 0}};

}  // namespace

int main(int argc, char** argv) {
    RayTracer::Config config;
    config.width = 1920; config.height = 1080; config.samplesPerPixel = 16; config.maxBounces = 4;
    uint32_t seed = 0;
    int pose = 0, frames = 5;
    bool legacy = false;
    std::string ppm, raw;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto next = [&]() { return i + 1 < argc ? argv[++i] : "0"; };
        if (a == "--width") config.width = std::atoi(next());
        else if (a == "--height") config.height = std::atoi(next());
        else if (a == "--spp") config.samplesPerPixel = std::atoi(next());
        else if (a == "--bounces") config.maxBounces = std::atoi(next());
        else if (a == "--seed") seed = static_cast<uint32_t>(std::atoi(next()));
        else if (a == "--pose") pose = std::atoi(next());
        else if (a == "--frames") frames = std::atoi(next());
        else if (a == "--legacy") legacy = true;
        else if (a == "--hard-shadows") config.softShadows = false;
        else if (a == "--ppm") ppm = next();
        else if (a == "--raw") raw = next();
        else { std::fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
    }
    if (pose < 0 || pose > 6) pose = 0;

    const std::vector<uint8_t> atlas = synthSkin(seed, legacy);
    std::vector<McBox> boxes(MCSKIN_MAX_SKIN_BOXES);
    std::vector<float> texels(static_cast<size_t>(MCSKIN_MAX_SKIN_TEXELS) * 4);
    McScene flat;
    if (mcskin_build_skin_scene(atlas.data(), 64, legacy ? 32 : 64, kPoses[pose], boxes.data(), texels.data(), &flat) != MC_OK) {
        std::fprintf(stderr, "scene build failed: %s\n", mcskin_cuda_last_error());
        return 1;
    }
    const Scene scene = unflattenScene(flat);

    int callbacks = 0;
    Image image = TileRenderer::render(scene, config, [&](int, int) { ++callbacks; });  // warm-up
    if (!TileRenderer::lastErrors().empty()) {
        std::fprintf(stderr, "render failed: %s\n", TileRenderer::lastErrors()[0].message.c_str());
        return 1;
    }
    const auto t0 = std::chrono::steady_clock::now();
    for (int f = 0; f < frames; ++f) image = TileRenderer::render(scene, config);
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / (frames > 0 ? frames : 1);

    if (!raw.empty()) {
        if (FILE* fp = std::fopen(raw.c_str(), "wb")) {
            std::fwrite(image.pixels.data(), sizeof(Color), image.pixels.size(), fp);
            std::fclose(fp);
        }
    }
    if (!ppm.empty()) {
        const std::vector<uint8_t> rgba = image.toRGBA8();
        if (FILE* fp = std::fopen(ppm.c_str(), "wb")) {
            std::fprintf(fp, "P6\n%d %d\n255\n", image.width, image.height);
            for (size_t i = 0; i < image.pixels.size(); ++i) std::fwrite(&rgba[4 * i], 1, 3, fp);
            std::fclose(fp);
        }
    }
    const int tiles = static_cast<int>(TileRenderer::generateTiles(config.width, config.height, config.tileSize).size());
    std::printf("{\"width\": %d, \"height\": %d, \"spp\": %d, \"bounces\": %d, \"meshes\": %zu, \"tiles\": %d, "
                "\"progress_callbacks\": %d, \"frames\": %d, \"ms_per_frame_e2e\": %.3f}\n",
                config.width, config.height, config.samplesPerPixel, config.maxBounces, scene.meshes.size(), tiles,
                callbacks, frames, ms);
    return 0;
}
