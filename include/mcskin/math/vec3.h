// Vec3 of the render hot path.  Same fields, operators and rounding behaviour as the
// reference's src/math/vec3.h:6-51 (division multiplies by the rounded reciprocal,
// normalize() returns the zero vector below 1e-8), so caller code compiles unchanged.
#pragma once

#include <cmath>

struct Vec3 {
    float x = 0.0f, y = 0.0f, z = 0.0f;

    constexpr Vec3() = default;
    constexpr Vec3(float xx, float yy, float zz) : x(xx), y(yy), z(zz) {}

    constexpr Vec3 operator-() const { return Vec3(-x, -y, -z); }
    constexpr Vec3 operator+(const Vec3& o) const { return Vec3(x + o.x, y + o.y, z + o.z); }
    constexpr Vec3 operator-(const Vec3& o) const { return Vec3(x - o.x, y - o.y, z - o.z); }
    constexpr Vec3 operator*(float k) const { return Vec3(x * k, y * k, z * k); }
    Vec3 operator/(float k) const { return *this * (1.0f / k); }

    Vec3& operator+=(const Vec3& o) { return *this = *this + o; }
    Vec3& operator-=(const Vec3& o) { return *this = *this - o; }
    Vec3& operator*=(float k) { return *this = *this * k; }
    Vec3& operator/=(float k) { return *this = *this / k; }

    friend constexpr Vec3 operator*(float k, const Vec3& v) { return Vec3(k * v.x, k * v.y, k * v.z); }
    constexpr bool operator==(const Vec3& o) const { return x == o.x && y == o.y && z == o.z; }
    constexpr bool operator!=(const Vec3& o) const { return !(*this == o); }

    constexpr float dot(const Vec3& o) const { return x * o.x + y * o.y + z * o.z; }
    constexpr Vec3 cross(const Vec3& o) const {
        return Vec3(y * o.z - z * o.y, z * o.x - x * o.z, x * o.y - y * o.x);
    }
    constexpr float lengthSquared() const { return dot(*this); }
    float length() const { return std::sqrt(lengthSquared()); }
    Vec3 normalize() const {
        const float len = length();
        return len < 1e-8f ? Vec3() : *this / len;
    }
};
