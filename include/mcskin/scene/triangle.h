// Triangle and HitResult (reference: src/scene/triangle.h:9-26).  The ray tracer treats a
// Mesh as one box; triangles only supply the box corners and the per-face texture pointer.
#pragma once

#include "math/color.h"
#include "math/vec3.h"

struct TextureRegion;

struct Triangle {
    Vec3 v0, v1, v2;
    Vec3 normal;
    float u0 = 0.0f, v0_uv = 0.0f;
    float u1 = 0.0f, v1_uv = 0.0f;
    float u2 = 0.0f, v2_uv = 0.0f;
    const TextureRegion* texture = nullptr;
};

struct HitResult {
    bool hit = false;
    float t = 0.0f;
    Vec3 point;
    Vec3 normal;
    Color textureColor;
    bool isOuterLayer = false;
};
