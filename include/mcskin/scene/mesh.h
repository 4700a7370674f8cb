// Mesh: one body-part box with its six owned face textures and optional pose
// (reference: src/scene/mesh.h:12-120).  Triangle::texture pointers that point into a
// mesh's own ownedTextures are re-pointed on copy and move; pointers elsewhere are kept.
#pragma once

#include <array>
#include <utility>
#include <vector>

#include "scene/triangle.h"
#include "skin/texture_region.h"

struct Mesh {
    std::vector<Triangle> triangles;
    bool isOuterLayer = false;
    std::array<TextureRegion, 6> ownedTextures;  // front, back, left, right, top, bottom

    bool hasRotation = false;
    Vec3 pivot;
    float rotX = 0.0f;  // pitch, degrees
    float rotZ = 0.0f;  // roll, degrees
    std::vector<Triangle> localTriangles;  // unposed copy the ray tracer intersects

    Mesh() = default;
    Mesh(const Mesh& o) { assign(o); }
    Mesh(Mesh&& o) noexcept { take(std::move(o)); }
    Mesh& operator=(const Mesh& o) {
        if (this != &o) assign(o);
        return *this;
    }
    Mesh& operator=(Mesh&& o) noexcept {
        if (this != &o) take(std::move(o));
        return *this;
    }

private:
    void copyScalars(const Mesh& o) {
        isOuterLayer = o.isOuterLayer;
        hasRotation = o.hasRotation;
        pivot = o.pivot;
        rotX = o.rotX;
        rotZ = o.rotZ;
    }
    void rebind(const Mesh& from) {
        for (Triangle& t : triangles)
            for (int i = 0; i < 6 && t.texture; ++i)
                if (t.texture == &from.ownedTextures[i]) {
                    t.texture = &ownedTextures[i];
                    break;
                }
    }
    void assign(const Mesh& o) {
        triangles = o.triangles;
        localTriangles = o.localTriangles;
        ownedTextures = o.ownedTextures;
        copyScalars(o);
        rebind(o);
    }
    void take(Mesh&& o) {
        triangles = std::move(o.triangles);
        localTriangles = std::move(o.localTriangles);
        ownedTextures = std::move(o.ownedTextures);
        copyScalars(o);
        rebind(o);
    }
};
