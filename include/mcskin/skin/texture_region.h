// One face texture: row-major float RGBA with nearest-texel lookup
// (reference: src/skin/texture_region.h:7-27).
#pragma once

#include <utility>
#include <vector>

#include "math/color.h"

struct TextureRegion {
    int width = 0;
    int height = 0;
    std::vector<Color> pixels;

    TextureRegion() = default;
    TextureRegion(int w, int h) : width(w), height(h), pixels(static_cast<size_t>(w) * h) {}
    TextureRegion(int w, int h, std::vector<Color> px) : width(w), height(h), pixels(std::move(px)) {}

    // u, v in [0,1]; truncating nearest neighbour, clamped to the region; an empty region is Color()
    Color sample(float u, float v) const {
        if (width <= 0 || height <= 0 || pixels.empty()) return Color();
        auto texel = [](float t, int n) {
            const int i = static_cast<int>(t * n);
            return i < 0 ? 0 : (i > n - 1 ? n - 1 : i);
        };
        return pixels[static_cast<size_t>(texel(v, height)) * width + texel(u, width)];
    }
};
