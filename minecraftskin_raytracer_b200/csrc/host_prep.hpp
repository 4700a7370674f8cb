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

// Returns MC_OK or a negative MC_ERR_* with `error` filled.  cfg may be null for
// scene-only queries (intersect, in_shadow ...): reference defaults are used then.
int prepare_frame(const McScene* scene, const McConfig* cfg, int useConfig, float aspectOverride,
                  PreparedFrame& out, std::string& error);

void set_last_error(const std::string& message);

}  // namespace mcskin
