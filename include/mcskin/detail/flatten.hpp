// flatten.hpp — Scene / RayTracer::Config  ->  the C-ABI PODs of mcskin_cuda.h.
//
// Written as templates over the scene types so the same walk serves the
// re-authored headers in include/mcskin/ and (in oracle/ref_shim.cpp, test
// infrastructure only) the reference's own structs; it includes neither.
//
// What the walk has to preserve (reference file:line):
//  * bounds come from Mesh::localTriangles for posed meshes, Mesh::triangles
//    otherwise, as a min/max over all 36 vertices (intersection.cpp:45-64,374-395);
//    an empty list means the mesh can never be hit (intersection.cpp:205).
//  * the texture of face f is whatever Mesh::triangles[2*f].texture points at —
//    not ownedTextures[f]; tests hang meshes on an external static TextureRegion
//    (tests/test_raytracer.cpp:100-146).  A missing triangle or a null pointer
//    means magenta (intersection.cpp:124-129,303-306).
//  * a TextureRegion with no area or no pixels samples as (0,0,0,1)
//    (texture_region.h:20-22).
#pragma once

#include <cstddef>
#include <cstdint>
#include <limits>
#include <map>
#include <vector>

#include "mcskin_cuda.h"

namespace mcskin {

struct FlatScene {
    std::vector<McBox> boxes;
    std::vector<float> texels;  // RGBA float
    McScene scene{};            // pointers refer to the two vectors above

    void rebind() {
        scene.n_boxes = static_cast<int32_t>(boxes.size());
        scene.boxes = boxes.empty() ? nullptr : boxes.data();
        scene.n_texels = static_cast<int32_t>(texels.size() / 4);
        scene.texels_rgba = texels.empty() ? nullptr : texels.data();
    }
};

template <class TriangleList>
inline void triangleBounds(const TriangleList& tris, float lo[3], float hi[3]) {
    const float big = std::numeric_limits<float>::max();
    lo[0] = lo[1] = lo[2] = big;
    hi[0] = hi[1] = hi[2] = -big;
    for (const auto& t : tris) {
        const float vx[3] = {t.v0.x, t.v1.x, t.v2.x};
        const float vy[3] = {t.v0.y, t.v1.y, t.v2.y};
        const float vz[3] = {t.v0.z, t.v1.z, t.v2.z};
        for (int k = 0; k < 3; ++k) {
            if (vx[k] < lo[0]) lo[0] = vx[k];
            if (vy[k] < lo[1]) lo[1] = vy[k];
            if (vz[k] < lo[2]) lo[2] = vz[k];
            if (vx[k] > hi[0]) hi[0] = vx[k];
            if (vy[k] > hi[1]) hi[1] = vy[k];
            if (vz[k] > hi[2]) hi[2] = vz[k];
        }
    }
}

template <class SceneT>
inline void flattenScene(const SceneT& src, FlatScene& out) {
    out.boxes.clear();
    out.texels.clear();
    std::map<const void*, McFaceTex> seen;  // one pool window per distinct TextureRegion

    for (const auto& mesh : src.meshes) {
        McBox box{};
        const auto& boundsList = mesh.hasRotation ? mesh.localTriangles : mesh.triangles;
        box.n_triangles = static_cast<int32_t>(boundsList.size());
        if (box.n_triangles > 0) {
            triangleBounds(boundsList, box.bounds_min, box.bounds_max);
        }
        box.pivot[0] = mesh.pivot.x;
        box.pivot[1] = mesh.pivot.y;
        box.pivot[2] = mesh.pivot.z;
        box.rot_x_deg = mesh.rotX;
        box.rot_z_deg = mesh.rotZ;
        box.has_rotation = mesh.hasRotation ? 1 : 0;
        box.is_outer_layer = mesh.isOuterLayer ? 1 : 0;

        for (int f = 0; f < 6; ++f) {
            McFaceTex ft{-1, 0, 0};
            const std::size_t triIndex = static_cast<std::size_t>(f) * 2;
            const auto* tex = triIndex < mesh.triangles.size() ? mesh.triangles[triIndex].texture : nullptr;
            if (tex != nullptr) {
                auto it = seen.find(static_cast<const void*>(tex));
                if (it != seen.end()) {
                    ft = it->second;
                } else {
                    const bool blank = tex->width <= 0 || tex->height <= 0 || tex->pixels.empty();
                    ft.texel_offset = static_cast<int32_t>(out.texels.size() / 4);
                    ft.width = blank ? 0 : tex->width;
                    ft.height = blank ? 0 : tex->height;
                    if (!blank) {
                        const std::size_t n = static_cast<std::size_t>(tex->width) * tex->height;
                        for (std::size_t i = 0; i < n && i < tex->pixels.size(); ++i) {
                            const auto& c = tex->pixels[i];
                            out.texels.insert(out.texels.end(), {c.r, c.g, c.b, c.a});
                        }
                        // a region shorter than width*height would be out-of-bounds reads in
                        // the reference; pad with the default Color so the window is complete
                        for (std::size_t i = tex->pixels.size(); i < n; ++i)
                            out.texels.insert(out.texels.end(), {0.f, 0.f, 0.f, 1.f});
                    }
                    seen.emplace(static_cast<const void*>(tex), ft);
                }
            }
            box.face[f] = ft;
        }
        out.boxes.push_back(box);
    }

    McScene& s = out.scene;
    s.light_pos[0] = src.light.position.x;
    s.light_pos[1] = src.light.position.y;
    s.light_pos[2] = src.light.position.z;
    s.light_color[0] = src.light.color.r;
    s.light_color[1] = src.light.color.g;
    s.light_color[2] = src.light.color.b;
    s.light_color[3] = src.light.color.a;
    s.light_radius = src.light.radius;
    s.cam_pos[0] = src.camera.position.x;
    s.cam_pos[1] = src.camera.position.y;
    s.cam_pos[2] = src.camera.position.z;
    s.cam_target[0] = src.camera.target.x;
    s.cam_target[1] = src.camera.target.y;
    s.cam_target[2] = src.camera.target.z;
    s.cam_up[0] = src.camera.up.x;
    s.cam_up[1] = src.camera.up.y;
    s.cam_up[2] = src.camera.up.z;
    s.cam_fov_deg = src.camera.fov;
    s.background[0] = src.backgroundColor.r;
    s.background[1] = src.backgroundColor.g;
    s.background[2] = src.backgroundColor.b;
    s.background[3] = src.backgroundColor.a;
    out.rebind();
}

template <class ConfigT>
inline McConfig flattenConfig(const ConfigT& c) {
    McConfig f{};
    // render() always shades with ShadingParams{} (tile_renderer.cpp:106, shading.h:9-14)
    f.kd = 0.75f;
    f.ks = 0.15f;
    f.ambient = 0.20f;
    f.shininess = 16.0f;
    f.width = c.width;
    f.height = c.height;
    f.max_bounces = c.maxBounces;
    f.samples_per_pixel = c.samplesPerPixel;
    f.tile_size = c.tileSize;
    f.thread_count = c.threadCount;
    f.soft_shadows = c.softShadows ? 1 : 0;
    f.shadow_samples = c.shadowSamples;
    f.ao_enabled = c.aoEnabled ? 1 : 0;
    f.ao_samples = c.aoSamples;
    f.ao_radius = c.aoRadius;
    f.ao_intensity = c.aoIntensity;
    f.dof_enabled = c.dofEnabled ? 1 : 0;
    f.aperture = c.aperture;
    f.focus_distance = c.focusDistance;
    f.gradient_bg = c.gradientBg ? 1 : 0;
    f.gradient_scale = c.gradientScale;
    const float ctr[4] = {c.bgCenter.r, c.bgCenter.g, c.bgCenter.b, c.bgCenter.a};
    const float edg[4] = {c.bgEdge.r, c.bgEdge.g, c.bgEdge.b, c.bgEdge.a};
    for (int i = 0; i < 4; ++i) {
        f.bg_center[i] = ctr[i];
        f.bg_edge[i] = edg[i];
    }
    return f;
}

}  // namespace mcskin
