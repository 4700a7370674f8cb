// TileRenderer — the drop-in boundary of the render hot path
// (reference: src/raytracer/tile_renderer.h:11-47).  Same signatures; render() flattens the
// scene and hands it to libmcskin_cuda (include/mcskin_cuda.h).  There is no CPU fallback:
// without a usable CUDA device the call records a TileError and returns a default image.
#pragma once

#include <functional>
#include <string>
#include <vector>

#include "raytracer/raytracer.h"
#include "scene/scene.h"
#include "skin/image.h"

struct Tile {
    int x, y;
    int width, height;
};

class TileRenderer {
public:
    static std::vector<Tile> generateTiles(int imageWidth, int imageHeight, int tileSize);

    // Blocks until the frame is done.  progressCallback(done, total) is called total times
    // (total = number of tiles) on the calling thread.
    static Image render(const Scene& scene, const RayTracer::Config& config,
                        std::function<void(int, int)> progressCallback = nullptr);

    static void renderTile(const Tile& tile, const Scene& scene, const RayTracer::Config& config, Image& output);

    struct TileError {
        int tileIndex;  // -1: the failure concerns the whole launch (CUDA error), not one tile
        std::string message;
    };
    static const std::vector<TileError>& lastErrors();

private:
    static std::vector<TileError> errors_;
};
