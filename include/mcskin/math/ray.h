// Ray (reference: src/math/ray.h:5-15).
#pragma once

#include "math/vec3.h"

struct Ray {
    Vec3 origin;
    Vec3 direction;

    Ray() = default;
    Ray(const Vec3& o, const Vec3& d) : origin(o), direction(d) {}

    Vec3 at(float t) const { return origin + direction * t; }
};
