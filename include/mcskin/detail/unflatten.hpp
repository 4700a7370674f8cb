// unflatten.hpp — McScene -> Scene: builds the C++ scene objects (one 12-triangle box Mesh
// per McBox, six owned TextureRegions each) that flattenScene() maps back to the same
// McScene.  Used by the headless CLI and the tests to drive TileRenderer::render through the
// reference-shaped API starting from mcskin_build_skin_scene's output.
#pragma once

#include <vector>

#include "mcskin_cuda.h"

namespace mcskin {

template <class SceneT, class MeshT, class TriangleT, class TextureT, class Vec3T, class ColorT>
inline SceneT unflattenSceneAs(const McScene& f) {
    SceneT sc;
    sc.meshes.reserve(static_cast<size_t>(f.n_boxes));
    // corner index = x + 2y + 4z; quads in reference face order -Z,+Z,+X,-X,+Y,-Y
    static const int quad[6][4] = {{2, 3, 1, 0}, {7, 6, 4, 5}, {3, 7, 5, 1}, {6, 2, 0, 4}, {6, 7, 3, 2}, {0, 1, 5, 4}};
    static const float nrm[6][3] = {{0, 0, -1}, {0, 0, 1}, {1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}};
    // ownedTextures order is front, back, left, right, top, bottom = faces 1,0,2,3,4,5
    static const int ownedOfFace[6] = {1, 0, 2, 3, 4, 5};
    for (int b = 0; b < f.n_boxes; ++b) {
        const McBox& box = f.boxes[b];
        sc.meshes.emplace_back();
        MeshT& m = sc.meshes.back();
        m.isOuterLayer = box.is_outer_layer != 0;
        m.hasRotation = box.has_rotation != 0;
        m.pivot = Vec3T(box.pivot[0], box.pivot[1], box.pivot[2]);
        m.rotX = box.rot_x_deg;
        m.rotZ = box.rot_z_deg;
        const float* lo = box.bounds_min;
        const float* hi = box.bounds_max;
        const Vec3T c[8] = {Vec3T(lo[0], lo[1], lo[2]), Vec3T(hi[0], lo[1], lo[2]), Vec3T(lo[0], hi[1], lo[2]),
                            Vec3T(hi[0], hi[1], lo[2]), Vec3T(lo[0], lo[1], hi[2]), Vec3T(hi[0], lo[1], hi[2]),
                            Vec3T(lo[0], hi[1], hi[2]), Vec3T(hi[0], hi[1], hi[2])};
        for (int face = 0; face < 6; ++face) {
            const McFaceTex& ft = box.face[face];
            TextureT& tex = m.ownedTextures[ownedOfFace[face]];
            const bool hasTexels = ft.texel_offset >= 0 && ft.width > 0 && ft.height > 0;
            if (hasTexels) {
                tex = TextureT(ft.width, ft.height);
                for (int i = 0; i < ft.width * ft.height; ++i) {
                    const float* p = f.texels_rgba + (static_cast<size_t>(ft.texel_offset) + i) * 4;
                    tex.pixels[i] = ColorT(p[0], p[1], p[2], p[3]);
                }
            }
            for (int half = 0; half < 2 && box.n_triangles > 0; ++half) {
                TriangleT t;
                t.v0 = c[quad[face][0]];
                t.v1 = c[quad[face][half ? 2 : 1]];
                t.v2 = c[quad[face][half ? 3 : 2]];
                t.normal = Vec3T(nrm[face][0], nrm[face][1], nrm[face][2]);
                t.texture = ft.texel_offset < 0 ? nullptr : &tex;
                m.triangles.push_back(t);
            }
        }
        if (m.hasRotation) m.localTriangles = m.triangles;
    }
    sc.light.position = Vec3T(f.light_pos[0], f.light_pos[1], f.light_pos[2]);
    sc.light.color = ColorT(f.light_color[0], f.light_color[1], f.light_color[2], f.light_color[3]);
    sc.light.radius = f.light_radius;
    sc.camera.position = Vec3T(f.cam_pos[0], f.cam_pos[1], f.cam_pos[2]);
    sc.camera.target = Vec3T(f.cam_target[0], f.cam_target[1], f.cam_target[2]);
    sc.camera.up = Vec3T(f.cam_up[0], f.cam_up[1], f.cam_up[2]);
    sc.camera.fov = f.cam_fov_deg;
    sc.backgroundColor = ColorT(f.background[0], f.background[1], f.background[2], f.background[3]);
    return sc;
}

}  // namespace mcskin

// With the scene headers in scope: Scene unflattenScene(const McScene&)
#define MCSKIN_DEFINE_UNFLATTEN()                                                                          \
    inline Scene unflattenScene(const McScene& f) {                                                        \
        return mcskin::unflattenSceneAs<Scene, Mesh, Triangle, TextureRegion, Vec3, Color>(f);              \
    }
