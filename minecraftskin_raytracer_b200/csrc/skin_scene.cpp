// skin_scene.cpp — RGBA8 skin atlas + pose -> flat scene, on the host.
//
// The callers either side of the hot path: what
// SkinParser::parse (skin_parser.cpp:11-132) and MeshBuilder::buildScene
// (mesh_builder.cpp:66-202) produce, emitted directly as McBox records and one
// texel pool instead of vectors of triangles the ray tracer never reads.
// Texels are byte / 255.0f (image.cpp:14-21).  The pool order (per box: back, front,
// left, right, top, bottom = reference faces 0..5) is the order a walk over
// Mesh::triangles[2*f].texture meets the regions, so the result is identical to
// flattening the reference's own Scene.
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "host_prep.hpp"
#include "mcskin_cuda.h"

namespace {

struct Region {
    int w = 0, h = 0;
    std::vector<float> px;  // RGBA
    bool fullyTransparent() const {  // mesh_builder.cpp:8-13 (vacuously true when empty)
        for (size_t i = 3; i < px.size(); i += 4)
            if (px[i] != 0.0f) return false;
        return true;
    }
};

struct PartTextures {
    Region top, bottom, front, back, left, right;
    bool fullyTransparent() const {
        return top.fullyTransparent() && bottom.fullyTransparent() && front.fullyTransparent() &&
               back.fullyTransparent() && left.fullyTransparent() && right.fullyTransparent();
    }
};

struct Atlas {
    const uint8_t* rgba;
    int w, h;
};

// Image::extractRegion (image.h:21-33): texels outside the atlas keep Color() = (0,0,0,1)
Region cut(const Atlas& a, int x, int y, int w, int h) {
    Region r;
    r.w = w;
    r.h = h;
    r.px.resize(static_cast<size_t>(w) * h * 4);
    for (int row = 0; row < h; ++row)
        for (int col = 0; col < w; ++col) {
            float* o = &r.px[(static_cast<size_t>(row) * w + col) * 4];
            const int sx = x + col, sy = y + row;
            if (sx >= 0 && sx < a.w && sy >= 0 && sy < a.h) {
                const uint8_t* s = a.rgba + (static_cast<size_t>(sy) * a.w + sx) * 4;
                for (int c = 0; c < 4; ++c) o[c] = s[c] / 255.0f;
            } else {
                o[0] = o[1] = o[2] = 0.0f;
                o[3] = 1.0f;
            }
        }
    return r;
}

// box unwrap of one body part (skin_parser.cpp:11-20)
PartTextures unwrap(const Atlas& a, int ox, int oy, int w, int h, int d) {
    PartTextures p;
    p.top = cut(a, ox + d, oy, w, d);
    p.bottom = cut(a, ox + d + w, oy, w, d);
    p.left = cut(a, ox, oy + d, d, h);
    p.front = cut(a, ox + d, oy + d, w, h);
    p.right = cut(a, ox + d + w, oy + d, d, h);
    p.back = cut(a, ox + 2 * d + w, oy + d, w, h);
    return p;
}

Region mirrored(const Region& r) {  // skin_parser.cpp:22-31
    Region m;
    m.w = r.w;
    m.h = r.h;
    m.px.resize(r.px.size());
    for (int y = 0; y < r.h; ++y)
        for (int x = 0; x < r.w; ++x)
            std::memcpy(&m.px[(static_cast<size_t>(y) * r.w + x) * 4],
                        &r.px[(static_cast<size_t>(y) * r.w + (r.w - 1 - x)) * 4], 4 * sizeof(float));
    return m;
}

PartTextures mirroredPart(const PartTextures& p) {  // skin_parser.cpp:33-43
    PartTextures m;
    m.top = mirrored(p.top);
    m.bottom = mirrored(p.bottom);
    m.front = mirrored(p.front);
    m.back = mirrored(p.back);
    m.left = mirrored(p.right);
    m.right = mirrored(p.left);
    return m;
}

struct PartDef {
    const PartTextures* inner;
    const PartTextures* outer;
    float pos[3], size[3], pivot[3];
    float rotX, rotZ;
};

}  // namespace

extern "C" int32_t mcskin_build_skin_scene(const uint8_t* atlasRgba, int32_t atlasW, int32_t atlasH,
                                           const float* pose12, McBox* boxesOut, float* texelsOut,
                                           McScene* sceneOut) {
    if (!atlasRgba || !boxesOut || !texelsOut || !sceneOut) {
        mcskin::set_last_error("mcskin_build_skin_scene: null argument");
        return MC_ERR_INVALID;
    }
    const bool isNew = atlasW == 64 && atlasH == 64;
    const bool isOld = atlasW == 64 && atlasH == 32;
    if (!isNew && !isOld) {  // skin_parser.cpp:122-131
        mcskin::set_last_error("Invalid skin dimensions: " + std::to_string(atlasW) + "x" + std::to_string(atlasH) +
                               " (expected 64x64 or 64x32)");
        return MC_ERR_INVALID;
    }
    const Atlas a{atlasRgba, atlasW, atlasH};
    PartTextures head, body, rArm, lArm, rLeg, lLeg, headO, bodyO, rArmO, lArmO, rLegO, lLegO;
    head = unwrap(a, 0, 0, 8, 8, 8);
    headO = unwrap(a, 32, 0, 8, 8, 8);
    body = unwrap(a, 16, 16, 8, 12, 4);
    rArm = unwrap(a, 40, 16, 4, 12, 4);
    rLeg = unwrap(a, 0, 16, 4, 12, 4);
    if (isNew) {  // skin_parser.cpp:45-80
        bodyO = unwrap(a, 16, 32, 8, 12, 4);
        rArmO = unwrap(a, 40, 32, 4, 12, 4);
        lArm = unwrap(a, 32, 48, 4, 12, 4);
        lArmO = unwrap(a, 48, 48, 4, 12, 4);
        rLegO = unwrap(a, 0, 32, 4, 12, 4);
        lLeg = unwrap(a, 16, 48, 4, 12, 4);
        lLegO = unwrap(a, 0, 48, 4, 12, 4);
    } else {  // skin_parser.cpp:82-110: limbs mirrored, no outer layers but the head's
        lArm = mirroredPart(rArm);
        lLeg = mirroredPart(rLeg);
    }

    float pose[12] = {0};
    if (pose12) std::memcpy(pose, pose12, sizeof(pose));
    // mesh_builder.cpp:163-170: head, body, rightArm, leftArm, rightLeg, leftLeg
    const PartDef parts[6] = {
        {&head, &headO, {0, 28, 0}, {8, 8, 8}, {0, 24, 0}, pose[0], pose[1]},
        {&body, &bodyO, {0, 18, 0}, {8, 12, 4}, {0, 18, 0}, pose[2], pose[3]},
        {&rArm, &rArmO, {-6, 18, 0}, {4, 12, 4}, {-6, 24, 0}, pose[4], pose[5]},
        {&lArm, &lArmO, {6, 18, 0}, {4, 12, 4}, {6, 24, 0}, pose[6], pose[7]},
        {&rLeg, &rLegO, {-2, 6, 0}, {4, 12, 4}, {-2, 12, 0}, pose[8], pose[9]},
        {&lLeg, &lLegO, {2, 6, 0}, {4, 12, 4}, {2, 12, 0}, pose[10], pose[11]},
    };

    int nBoxes = 0, nTexels = 0;
    auto emit = [&](const PartTextures& tex, const PartDef& part, float offset, bool posed) {
        McBox box;
        std::memset(&box, 0, sizeof(box));
        for (int k = 0; k < 3; ++k) {  // mesh_builder.cpp:83-91
            const float half = part.size[k] / 2.0f + offset;
            box.bounds_min[k] = part.pos[k] - half;
            box.bounds_max[k] = part.pos[k] + half;
        }
        box.is_outer_layer = offset > 0.0f ? 1 : 0;
        box.n_triangles = 12;
        if (posed) {  // mesh_builder.cpp:125-143
            box.has_rotation = 1;
            std::memcpy(box.pivot, part.pivot, sizeof(box.pivot));
            box.rot_x_deg = part.rotX;
            box.rot_z_deg = part.rotZ;
        }
        // reference face order -Z,+Z,+X,-X,+Y,-Y = back, front, left, right, top, bottom (mesh_builder.cpp:115-120)
        const Region* faces[6] = {&tex.back, &tex.front, &tex.left, &tex.right, &tex.top, &tex.bottom};
        for (int f = 0; f < 6; ++f) {
            const Region& r = *faces[f];
            const bool blank = r.w <= 0 || r.h <= 0 || r.px.empty();
            box.face[f].texel_offset = nTexels;
            box.face[f].width = blank ? 0 : r.w;
            box.face[f].height = blank ? 0 : r.h;
            if (!blank) {
                std::memcpy(texelsOut + static_cast<size_t>(nTexels) * 4, r.px.data(), r.px.size() * sizeof(float));
                nTexels += r.w * r.h;
            }
        }
        boxesOut[nBoxes++] = box;
    };
    for (const PartDef& part : parts) {
        const bool posed = std::fabs(part.rotX) > 0.01f || std::fabs(part.rotZ) > 0.01f;  // mesh_builder.cpp:173
        emit(*part.inner, part, 0.0f, posed);
        if (!part.outer->fullyTransparent()) emit(*part.outer, part, 0.5f, posed);
    }

    McScene s;
    std::memset(&s, 0, sizeof(s));
    s.n_boxes = nBoxes;
    s.boxes = boxesOut;
    s.n_texels = nTexels;
    s.texels_rgba = texelsOut;
    // mesh_builder.cpp:190-199, scene.h:10-15
    s.light_pos[0] = 0; s.light_pos[1] = 40; s.light_pos[2] = 30;
    s.light_color[0] = s.light_color[1] = s.light_color[2] = s.light_color[3] = 1.0f;
    s.light_radius = 3.0f;
    s.cam_pos[0] = 0; s.cam_pos[1] = 18; s.cam_pos[2] = 50;
    s.cam_target[0] = 0; s.cam_target[1] = 18; s.cam_target[2] = 0;
    s.cam_up[0] = 0; s.cam_up[1] = 1; s.cam_up[2] = 0;
    s.cam_fov_deg = 60.0f;
    s.background[0] = 0.2f; s.background[1] = 0.3f; s.background[2] = 0.5f; s.background[3] = 1.0f;
    *sceneOut = s;
    return MC_OK;
}
