// tile_renderer.cpp — TileRenderer over the CUDA C ABI.
//
// Drop-in for the reference's src/raytracer/tile_renderer.cpp: same class, same signatures,
// same observable contract (tile_renderer.h:16-47, tile_renderer.cpp:18-39,129-193):
//   * generateTiles: row-major tiles, edges clipped, {} for non-positive arguments;
//   * render: blocks, never throws, returns Image(width, height); the progress callback is
//     called exactly totalTiles times with (1..totalTiles, totalTiles); failures are recorded
//     in lastErrors() instead of aborting the caller;
//   * renderTile: one tile of the same frame (same per-tile RNG stream).
// Compiles against either include/mcskin/ or the reference's own src/ headers (the structs
// are source compatible); the scene is flattened by mcskin/detail/flatten.hpp.
#include "raytracer/tile_renderer.h"

#include <mutex>

#include "mcskin/detail/flatten.hpp"
#include "mcskin_cuda.h"

std::vector<TileRenderer::TileError> TileRenderer::errors_;

namespace {
std::mutex g_renderMutex;  // errors_ is process-wide, as in the reference (tile_renderer.cpp:16)

int deviceFromEnvironment() { return 0; }

void progressThunk(int32_t done, int32_t total, void* user) {
    auto* fn = static_cast<std::function<void(int, int)>*>(user);
    if (*fn) (*fn)(done, total);
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
    if (mcskin_generate_tiles(config.width, config.height, config.tileSize, nullptr, 0) == 0) return output;

    mcskin::FlatScene flat;
    mcskin::flattenScene(scene, flat);
    const McConfig cfg = mcskin::flattenConfig(config);
    static_assert(sizeof(Color) == 4 * sizeof(float), "Image::pixels is handed to the C ABI as float RGBA");
    const int rc = mcskin_cuda_render(&flat.scene, &cfg, deviceFromEnvironment(), &output.pixels[0].r, nullptr,
                                      progressCallback ? &progressThunk : nullptr, &progressCallback, nullptr);
    if (rc != MC_OK) {
        errors_.push_back(TileError{-1, mcskin_cuda_last_error()});
        output = Image(config.width, config.height);
    }
    return output;
}

void TileRenderer::renderTile(const Tile& tile, const Scene& scene, const RayTracer::Config& config, Image& output) {
    if (output.width != config.width || output.height != config.height || output.pixels.empty()) return;
    mcskin::FlatScene flat;
    mcskin::flattenScene(scene, flat);
    const McConfig cfg = mcskin::flattenConfig(config);
    const McTile t{tile.x, tile.y, tile.width, tile.height};
    const int rc = mcskin_cuda_render_tile(&flat.scene, &cfg, deviceFromEnvironment(), &t, &output.pixels[0].r, nullptr);
    if (rc != MC_OK) {
        std::lock_guard<std::mutex> lock(g_renderMutex);
        errors_.push_back(TileError{-1, mcskin_cuda_last_error()});
    }
}

const std::vector<TileRenderer::TileError>& TileRenderer::lastErrors() { return errors_; }
