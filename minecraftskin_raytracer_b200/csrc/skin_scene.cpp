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

namespace mcskin {

namespace {

struct Window {
    int x, y, w, h;
    bool mirror;
};
// the six windows of one body part in reference face order: back, front, left, right, top, bottom
// (skin_parser.cpp:11-20 for the unwrap, mesh_builder.cpp:115-120 for the order)
void part_windows(int ox, int oy, int w, int h, int d, bool mirroredPart, Window out[6]) {
    const Window top{ox + d, oy, w, d, false}, bottom{ox + d + w, oy, w, d, false};
    const Window left{ox, oy + d, d, h, false}, front{ox + d, oy + d, w, h, false};
    const Window right{ox + d + w, oy + d, d, h, false}, back{ox + 2 * d + w, oy + d, w, h, false};
    if (!mirroredPart) {
        out[0] = back; out[1] = front; out[2] = left; out[3] = right; out[4] = top; out[5] = bottom;
    } else {  // skin_parser.cpp:33-43: every region flipped horizontally, left and right swapped
        out[0] = back; out[1] = front; out[2] = right; out[3] = left; out[4] = top; out[5] = bottom;
        for (int f = 0; f < 6; ++f) out[f].mirror = true;
    }
}
// alpha of texel (col, row) of a window as Image::extractRegion sees it: 255 outside the atlas (Color() has a = 1)
inline int window_alpha(const uint8_t* rgba, int aw, int ah, const Window& wd, int col, int row) {
    const int sx = wd.x + (wd.mirror ? wd.w - 1 - col : col), sy = wd.y + row;
    if (sx < 0 || sx >= aw || sy < 0 || sy >= ah) return 255;
    return rgba[(static_cast<size_t>(sy) * aw + sx) * 4 + 3];
}

}  // namespace

// What mcskin_build_skin_scene produces, minus the float texels: the boxes (with their face windows into a pool
// that is never materialised on the host), where every face's texels come from in the atlas, and whether a box
// has a texel with alpha 0 at all.  Only alpha bytes are read.  Returns MC_OK or MC_ERR_INVALID (bad atlas size).
int skin_layout(const uint8_t* atlasRgba, int atlasW, int atlasH, const float* pose12, McBox* boxesOut, SkinFaceSource* facesOut,
                int* nFacesOut, uint8_t* boxOpaqueOut, McScene* sceneOut) {
    const bool isNew = atlasW == 64 && atlasH == 64;
    const bool isOld = atlasW == 64 && atlasH == 32;
    if (!isNew && !isOld) return MC_ERR_INVALID;
    struct PartSrc { int ox, oy, w, h, d; bool mirror; bool present; };
    // inner, outer per part: head, body, rightArm, leftArm, rightLeg, leftLeg (skin_parser.cpp:45-110)
    const PartSrc inner[6] = {{0, 0, 8, 8, 8, false, true}, {16, 16, 8, 12, 4, false, true}, {40, 16, 4, 12, 4, false, true},
                              isNew ? PartSrc{32, 48, 4, 12, 4, false, true} : PartSrc{40, 16, 4, 12, 4, true, true},
                              {0, 16, 4, 12, 4, false, true},
                              isNew ? PartSrc{16, 48, 4, 12, 4, false, true} : PartSrc{0, 16, 4, 12, 4, true, true}};
    const PartSrc outer[6] = {{32, 0, 8, 8, 8, false, true}, {16, 32, 8, 12, 4, false, isNew}, {40, 32, 4, 12, 4, false, isNew},
                              {48, 48, 4, 12, 4, false, isNew}, {0, 32, 4, 12, 4, false, isNew}, {0, 48, 4, 12, 4, false, isNew}};
    static const float kPos[6][3] = {{0, 28, 0}, {0, 18, 0}, {-6, 18, 0}, {6, 18, 0}, {-2, 6, 0}, {2, 6, 0}};
    static const float kSize[6][3] = {{8, 8, 8}, {8, 12, 4}, {4, 12, 4}, {4, 12, 4}, {4, 12, 4}, {4, 12, 4}};
    static const float kPivot[6][3] = {{0, 24, 0}, {0, 18, 0}, {-6, 24, 0}, {6, 24, 0}, {-2, 12, 0}, {2, 12, 0}};
    float pose[12] = {0};
    if (pose12) std::memcpy(pose, pose12, sizeof(pose));
    int nBoxes = 0, nTexels = 0, nFaces = 0;
    auto emit = [&](const PartSrc& src, int part, float offset, bool posed, const Window wins[6], bool opaque) {
        McBox box;
        std::memset(&box, 0, sizeof(box));
        for (int k = 0; k < 3; ++k) {  // mesh_builder.cpp:83-91
            const float half = kSize[part][k] / 2.0f + offset;
            box.bounds_min[k] = kPos[part][k] - half;
            box.bounds_max[k] = kPos[part][k] + half;
        }
        box.is_outer_layer = offset > 0.0f ? 1 : 0;
        box.n_triangles = 12;
        if (posed) {  // mesh_builder.cpp:125-143
            box.has_rotation = 1;
            std::memcpy(box.pivot, kPivot[part], sizeof(box.pivot));
            box.rot_x_deg = pose[2 * part];
            box.rot_z_deg = pose[2 * part + 1];
        }
        for (int f = 0; f < 6; ++f) {
            box.face[f].texel_offset = nTexels;
            box.face[f].width = wins[f].w;
            box.face[f].height = wins[f].h;
            facesOut[nFaces++] = SkinFaceSource{nTexels, static_cast<int16_t>(wins[f].x), static_cast<int16_t>(wins[f].y), static_cast<int16_t>(wins[f].w),
                                                static_cast<int16_t>(wins[f].h), wins[f].mirror ? 1 : 0};
            nTexels += wins[f].w * wins[f].h;
        }
        boxOpaqueOut[nBoxes] = opaque ? 1 : 0;
        boxesOut[nBoxes++] = box;
        (void)src;
    };
    for (int part = 0; part < 6; ++part) {
        const bool posed = std::fabs(pose[2 * part]) > 0.01f || std::fabs(pose[2 * part + 1]) > 0.01f;  // mesh_builder.cpp:173
        for (int layer = 0; layer < 2; ++layer) {
            const PartSrc& src = layer == 0 ? inner[part] : outer[part];
            // an absent outer layer (legacy skins) has empty textures: vacuously fully transparent (mesh_builder.cpp:8-13)
            if (!src.present) continue;
            Window wins[6];
            part_windows(src.ox, src.oy, src.w, src.h, src.d, src.mirror, wins);
            bool anyVisible = false, anyHole = false;
            for (int f = 0; f < 6; ++f)
                for (int row = 0; row < wins[f].h; ++row)
                    for (int col = 0; col < wins[f].w; ++col) {
                        const int a = window_alpha(atlasRgba, atlasW, atlasH, wins[f], col, row);
                        anyVisible = anyVisible || a != 0;
                        anyHole = anyHole || a == 0;
                    }
            if (layer == 1 && !anyVisible) continue;  // isFullyTransparent: the outer mesh is not built (mesh_builder.cpp:176-187)
            emit(src, part, layer == 0 ? 0.0f : 0.5f, posed, wins, !anyHole);
        }
    }
    *nFacesOut = nFaces;
    McScene s;
    std::memset(&s, 0, sizeof(s));
    s.n_boxes = nBoxes;
    s.boxes = boxesOut;
    s.n_texels = nTexels;
    s.texels_rgba = nullptr;
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

}  // namespace mcskin

// skin_layout through the C ABI (host code): faces as 6 int32 each — dst, x, y, w, h, mirror.
extern "C" int32_t mcskin_skin_layout(const uint8_t* atlasRgba, int32_t atlasW, int32_t atlasH, const float* pose12, McBox* boxesOut,
                                      int32_t* facesOut, int32_t* nFacesOut, uint8_t* boxOpaqueOut, McScene* sceneOut) {
    if (!atlasRgba || !boxesOut || !facesOut || !nFacesOut || !boxOpaqueOut || !sceneOut) {
        mcskin::set_last_error("mcskin_skin_layout: null argument");
        return MC_ERR_INVALID;
    }
    mcskin::SkinFaceSource faces[mcskin::kSkinMaxFaces];
    int n = 0;
    if (mcskin::skin_layout(atlasRgba, atlasW, atlasH, pose12, boxesOut, faces, &n, boxOpaqueOut, sceneOut) != MC_OK) {
        mcskin::set_last_error("Invalid skin dimensions: " + std::to_string(atlasW) + "x" + std::to_string(atlasH) +
                               " (expected 64x64 or 64x32)");
        return MC_ERR_INVALID;
    }
    for (int i = 0; i < n; ++i) {
        const int32_t rec[6] = {faces[i].dst, faces[i].x, faces[i].y, faces[i].w, faces[i].h, faces[i].mirror};
        std::memcpy(facesOut + 6 * i, rec, sizeof(rec));
    }
    *nFacesOut = n;
    return MC_OK;
}

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
