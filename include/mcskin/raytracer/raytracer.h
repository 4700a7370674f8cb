// RayTracer::Config and the integrator entry points (reference: src/raytracer/raytracer.h:8-55).
#pragma once

#include "math/color.h"
#include "math/ray.h"
#include "raytracer/shading.h"
#include "scene/scene.h"

class RayTracer {
public:
    struct Config {
        int width = 256;
        int height = 256;
        int maxBounces = 3;
        int samplesPerPixel = 1;
        int tileSize = 32;
        int threadCount = 0;  // accepted for compatibility; the GPU path ignores it

        bool softShadows = true;
        int shadowSamples = 8;

        bool aoEnabled = false;
        int aoSamples = 8;
        float aoRadius = 3.0f;
        float aoIntensity = 0.5f;

        bool dofEnabled = false;
        float aperture = 0.5f;
        float focusDistance = 0.0f;  // 0 = distance to the camera target

        bool gradientBg = true;
        float gradientScale = 1.0f;
        Color bgCenter{0.91f, 0.89f, 0.86f, 1.0f};
        Color bgEdge{0.56f, 0.63f, 0.71f, 1.0f};
    };

    static Color traceRay(const Ray& ray, const Scene& scene, int depth, int maxBounces,
                          const ShadingParams& params = ShadingParams{}, const Config* config = nullptr);
    static Color backgroundColor(const Scene& scene, float u, float v, const Config* config);
    static float computeAO(const Vec3& point, const Vec3& normal, const Scene& scene, int samples, float radius,
                           unsigned int seed);
};
