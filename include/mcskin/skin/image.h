// Float RGBA image, the return type of TileRenderer::render
// (reference: src/skin/image.h:9-36).  PNG load/save stays with the caller's stb build;
// this core carries the in-memory type, region extraction and the 8-bit conversion rule.
#pragma once

#include <cstdint>
#include <vector>

#include "math/color.h"
#include "skin/texture_region.h"

struct Image {
    int width = 0;
    int height = 0;
    std::vector<Color> pixels;  // row-major, every pixel starts as Color() = (0,0,0,1)

    Image() = default;
    Image(int w, int h) : width(w), height(h), pixels(static_cast<size_t>(w > 0 ? w : 0) * (h > 0 ? h : 0)) {}

    TextureRegion extractRegion(int x, int y, int w, int h) const {
        TextureRegion out(w, h);
        for (int row = 0; row < h; ++row)
            for (int col = 0; col < w; ++col) {
                const int sx = x + col, sy = y + row;
                if (sx >= 0 && sx < width && sy >= 0 && sy < height)
                    out.pixels[static_cast<size_t>(row) * w + col] = pixels[static_cast<size_t>(sy) * width + sx];
            }
        return out;
    }

    // uint8 = clamp(c) * 255 + 0.5, truncated (image_writer.cpp:18-22, image.cpp:30-35)
    std::vector<std::uint8_t> toRGBA8() const {
        std::vector<std::uint8_t> out(pixels.size() * 4);
        for (size_t i = 0; i < pixels.size(); ++i) {
            const Color c = pixels[i].clamp();
            out[4 * i + 0] = static_cast<std::uint8_t>(c.r * 255.0f + 0.5f);
            out[4 * i + 1] = static_cast<std::uint8_t>(c.g * 255.0f + 0.5f);
            out[4 * i + 2] = static_cast<std::uint8_t>(c.b * 255.0f + 0.5f);
            out[4 * i + 3] = static_cast<std::uint8_t>(c.a * 255.0f + 0.5f);
        }
        return out;
    }
};
