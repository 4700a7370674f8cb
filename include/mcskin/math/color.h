// RGBA colour in floats; alpha defaults to 1 (reference: src/math/color.h:5-42).
#pragma once

struct Color {
    float r = 0.0f, g = 0.0f, b = 0.0f, a = 1.0f;

    constexpr Color() = default;
    constexpr Color(float rr, float gg, float bb, float aa = 1.0f) : r(rr), g(gg), b(bb), a(aa) {}

    constexpr Color operator+(const Color& o) const { return Color(r + o.r, g + o.g, b + o.b, a + o.a); }
    constexpr Color operator-(const Color& o) const { return Color(r - o.r, g - o.g, b - o.b, a - o.a); }
    constexpr Color operator*(const Color& o) const { return Color(r * o.r, g * o.g, b * o.b, a * o.a); }
    constexpr Color operator*(float k) const { return Color(r * k, g * k, b * k, a * k); }
    Color operator/(float k) const { return *this * (1.0f / k); }

    Color& operator+=(const Color& o) { return *this = *this + o; }
    Color& operator-=(const Color& o) { return *this = *this - o; }
    Color& operator*=(const Color& o) { return *this = *this * o; }
    Color& operator*=(float k) { return *this = *this * k; }

    friend constexpr Color operator*(float k, const Color& c) { return Color(k * c.r, k * c.g, k * c.b, k * c.a); }
    constexpr bool operator==(const Color& o) const { return r == o.r && g == o.g && b == o.b && a == o.a; }
    constexpr bool operator!=(const Color& o) const { return !(*this == o); }

    Color clamp() const {
        auto unit = [](float v) { return v < 0.0f ? 0.0f : (1.0f < v ? 1.0f : v); };
        return Color(unit(r), unit(g), unit(b), unit(a));
    }
};
