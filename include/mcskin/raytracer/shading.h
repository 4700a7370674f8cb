// ShadingParams and the shading entry points (reference: src/raytracer/shading.h:9-37).
// The functions run on the GPU through the C ABI (mcskin_cuda_in_shadow / _soft_shadow / _shade).
#pragma once

#include "math/color.h"
#include "math/vec3.h"
#include "scene/scene.h"
#include "scene/triangle.h"

struct ShadingParams {
    float kd = 0.75f;
    float ks = 0.15f;
    float ambient = 0.20f;
    float shininess = 16.0f;
};

bool isInShadow(const Vec3& point, const Vec3& normal, const Vec3& lightPos, const Scene& scene);
float computeSoftShadow(const Vec3& point, const Vec3& normal, const Light& light, const Scene& scene, int samples,
                        unsigned int seed);
Color shade(const HitResult& hit, const Vec3& viewDir, const Light& light, const Scene& scene,
            const ShadingParams& params = ShadingParams{}, float shadowFactor = -1.0f);
