// Light, Camera, Scene (reference: src/scene/scene.h:10-34).
#pragma once

#include <vector>

#include "math/color.h"
#include "math/ray.h"
#include "math/vec3.h"
#include "scene/mesh.h"

struct Light {
    Vec3 position;
    Color color;
    float intensity = 1.0f;  // carried for source compatibility; the ray tracer never reads it
    float radius = 3.0f;     // area-light radius of the soft shadows
};

struct Camera {
    Vec3 position;
    Vec3 target;
    Vec3 up;
    float fov = 60.0f;  // degrees

    // Pinhole ray through image coordinates (u, v) in [0,1]^2, v = 0 at the top
    // (reference: src/scene/camera.cpp:8-26).  Host-side evaluation.
    Ray generateRay(float u, float v, float aspectRatio) const;
};

struct Scene {
    std::vector<Mesh> meshes;
    Light light;
    Camera camera;
    Color backgroundColor;
};
