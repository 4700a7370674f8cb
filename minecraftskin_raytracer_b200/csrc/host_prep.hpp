// host_prep.hpp — host-side preparation of a frame: McScene + McConfig -> DevFrame,
// DevBox records and the texel pool uploaded to the device.
#pragma once
#include <string>
#include <vector>

#include "dev_types.cuh"
#include "mcskin_cuda.h"

namespace mcskin {

struct float4h {
    float x, y, z, w;
};

struct PreparedFrame {
    DevFrame frame;
    std::vector<DevBox> boxes;
    std::vector<unsigned char> blob;  // SceneBlobLayout image of `boxes`, 16-byte granular
    std::vector<float4h> texels;  // scene pool + two synthetic 1x1 textures (magenta, default Color)
};

// Texels that never exist on the host (skins sliced on the device): the scene's texels_rgba may be null; what
// prepare_frame would have learnt from them is handed in (per box: no texel of any face has alpha == 0).
struct ExternalTexels {
    const uint8_t* boxOpaque;
};

// Returns MC_OK or a negative MC_ERR_* with `error` filled.  cfg may be null for
// scene-only queries (intersect, in_shadow ...): reference defaults are used then.
// ext: see ExternalTexels; out.texels then holds only the two synthetic texels (which follow the scene's pool).
int prepare_frame(const McScene* scene, const McConfig* cfg, int useConfig, float aspectOverride,
                  PreparedFrame& out, std::string& error, const ExternalTexels* ext = nullptr);

// Where the texels of one face of a skin scene come from: `dst` = first texel in the scene's pool, (x, y, w, h) =
// window in the atlas, mirror = flipped horizontally (the left limbs of 64x32 skins).  16 bytes.
struct SkinFaceSource {
    int32_t dst;
    int16_t x, y, w, h;
    int32_t mirror;
};
constexpr int kSkinMaxBoxes = 12;
constexpr int kSkinMaxFaces = kSkinMaxBoxes * 6;
constexpr int kSkinMaxTexels = 3264;  // 64x64 skin, every outer layer present
// skin_scene.cpp: the flat scene of a skin without its float texels (see there).
int skin_layout(const uint8_t* atlasRgba, int atlasW, int atlasH, const float* pose12, McBox* boxesOut, SkinFaceSource* facesOut,
                int* nFacesOut, uint8_t* boxOpaqueOut, McScene* sceneOut);

void set_last_error(const std::string& message);

}  // namespace mcskin
