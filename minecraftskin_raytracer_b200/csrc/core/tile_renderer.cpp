// tile_renderer.cpp — TileRenderer over the CUDA C ABI.
//
// Drop-in for the reference's src/raytracer/tile_renderer.cpp: same class, same signatures,
// same observable contract (tile_renderer.h:16-47, tile_renderer.cpp:18-39,129-193):
//   * generateTiles: row-major tiles, edges clipped, {} for non-positive arguments;
//   * render: blocks, never throws, returns Image(width, height); the progress callback is
//     called exactly totalTiles times with (1..totalTiles, totalTiles) — also when the frame
//     failed, as the reference reports every tile whether or not it threw
//     (tile_renderer.cpp:150-171), so a caller waiting for done == total always gets there;
//     failures are recorded in lastErrors() instead of aborting the caller;
//   * renderTile: one tile of the same frame (same per-tile RNG stream); the tile must be one of
//     generateTiles(width, height, tileSize) — the kernels address tiles on that grid;
//   * which GPUs: MCSKIN_DEVICE=<index> picks the device of a single-GPU render (default 0);
//     MCSKIN_DEVICES=<n|all> spreads the frame's tiles over devices 0..n-1 of this process
//     (mcskin_cuda_render_multi: cost-balanced tile lists, every device copies its own part of the
//     image to the host over its own PCIe link).  Same pixels either way.
// Compiles against either include/mcskin/ or the reference's own src/ headers (the structs
// are source compatible); the scene is flattened by mcskin/detail/flatten.hpp.
#include "raytracer/tile_renderer.h"

#include <cstdlib>
#include <cstring>
#include <mutex>

#include "mcskin/detail/flatten.hpp"
#include "mcskin_cuda.h"

std::vector<TileRenderer::TileError> TileRenderer::errors_;

namespace {
std::mutex g_renderMutex;  // errors_ is process-wide, as in the reference (tile_renderer.cpp:16)

// MCSKIN_DEVICE: device index of single-GPU renders (default 0; out-of-range values fail in the C ABI).
int singleDevice() {
    const char* v = std::getenv("MCSKIN_DEVICE");
    return v && *v ? std::atoi(v) : 0;
}
// MCSKIN_DEVICES: number of devices a frame is spread over ("all" = every visible device); <= 1: one device.
int deviceSpread() {
    const char* v = std::getenv("MCSKIN_DEVICES");
    if (!v || !*v) return 1;
    if (std::strcmp(v, "all") == 0) return mcskin_cuda_device_count();
    return std::atoi(v);
}

}  // namespace

std::vector<Tile> TileRenderer::generateTiles(int imageWidth, int imageHeight, int tileSize) {
    const int32_t n = mcskin_generate_tiles(imageWidth, imageHeight, tileSize, nullptr, 0);
    std::vector<McTile> flat(static_cast<size_t>(n));
    if (n > 0) mcskin_generate_tiles(imageWidth, imageHeight, tileSize, flat.data(), n);
    std::vector<Tile> tiles;
    tiles.reserve(flat.size());
    for (const McTile& t : flat) tiles.push_back(Tile{t.x, t.y, t.width, t.height});
    return tiles;
}

Image TileRenderer::render(const Scene& scene, const RayTracer::Config& config,
                           std::function<void(int, int)> progressCallback) {
    std::lock_guard<std::mutex> lock(g_renderMutex);
    errors_.clear();
    Image output(config.width, config.height);
    const int totalTiles = mcskin_generate_tiles(config.width, config.height, config.tileSize, nullptr, 0);
    if (totalTiles == 0) return output;

    mcskin::FlatScene flat;
    mcskin::flattenScene(scene, flat);
    McConfig cfg = mcskin::flattenConfig(config);
    // MCSKIN_RNG=counter: counter-based random streams instead of std::mt19937 (not in RayTracer::Config; frames then
    // differ from the reference's in their noise only)
    if (const char* v = std::getenv("MCSKIN_RNG")) cfg.rng_mode = std::strcmp(v, "counter") == 0 ? MC_RNG_COUNTER : MC_RNG_MT19937;
    static_assert(sizeof(Color) == 4 * sizeof(float), "Image::pixels is handed to the C ABI as float RGBA");
    const int spread = deviceSpread();
    int rc;
    if (spread > 1) {
        rc = mcskin_cuda_render_multi(&flat.scene, &cfg, spread, &output.pixels[0].r, nullptr, nullptr);
    } else {
        rc = mcskin_cuda_render(&flat.scene, &cfg, singleDevice(), &output.pixels[0].r, nullptr, nullptr, nullptr, nullptr);
    }
    if (rc != MC_OK) {
        errors_.push_back(TileError{-1, mcskin_cuda_last_error()});
        output = Image(config.width, config.height);
    }
    if (progressCallback)
        for (int i = 1; i <= totalTiles; ++i) progressCallback(i, totalTiles);
    return output;
}

void TileRenderer::renderTile(const Tile& tile, const Scene& scene, const RayTracer::Config& config, Image& output) {
    if (output.width != config.width || output.height != config.height || output.pixels.empty()) return;
    mcskin::FlatScene flat;
    mcskin::flattenScene(scene, flat);
    const McConfig cfg = mcskin::flattenConfig(config);
    const McTile t{tile.x, tile.y, tile.width, tile.height};
    const int rc = mcskin_cuda_render_tile(&flat.scene, &cfg, singleDevice(), &t, &output.pixels[0].r, nullptr);
    if (rc != MC_OK) {
        std::lock_guard<std::mutex> lock(g_renderMutex);
        errors_.push_back(TileError{-1, mcskin_cuda_last_error()});
    }
}

const std::vector<TileRenderer::TileError>& TileRenderer::lastErrors() { return errors_; }
