// host_prep.cpp — frame preparation on the host (no CUDA calls).
//
// Everything here is evaluated with the host's float arithmetic and glibc libm in
// the reference's own operation order, so that values the reference recomputes per
// ray from frame constants come out bit-identical:
//   * camera look-at basis and tan(fov/2)          camera.cpp:10-16
//   * thin-lens focus distance                     tile_renderer.cpp:82-85
//   * pose sines / cosines, rad = deg*PI_f/180     intersection.cpp:16-19,26-29
//   * UV axis sizes with the 1e-8 guard            intersection.cpp:138-143
// Compiled with -ffp-contract=off; x86-64 baseline has no FMA to contract anyway.
#include "host_prep.hpp"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>
#include <limits>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace mcskin {

namespace {

thread_local std::string g_lastError;

struct H3 {
    float x, y, z;
};
H3 sub(H3 a, H3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
H3 crossh(H3 a, H3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
float lenh(H3 a) { return std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
H3 normh(H3 a) {  // vec3.h:46-50
    const float l = lenh(a);
    if (l < 1e-8f) return {0.0f, 0.0f, 0.0f};
    const float inv = 1.0f / l;
    return {a.x * inv, a.y * inv, a.z * inv};
}
float radians(float deg) { return deg * static_cast<float>(M_PI) / 180.0f; }

}  // namespace

void set_last_error(const std::string& message) { g_lastError = message; }

extern "C" const char* mcskin_cuda_last_error(void) { return g_lastError.c_str(); }
extern "C" int32_t mcskin_cuda_abi_version(void) { return MCSKIN_ABI_VERSION; }
extern "C" uint32_t mcskin_counter_word(uint32_t seed, uint32_t k) { return mc_rng_counter_word(seed, k); }

extern "C" void mcskin_config_defaults(McConfig* c) {
    if (!c) return;
    std::memset(c, 0, sizeof(*c));
    // raytracer.h:10-38
    c->width = 256;
    c->height = 256;
    c->max_bounces = 3;
    c->samples_per_pixel = 1;
    c->tile_size = 32;
    c->thread_count = 0;
    c->soft_shadows = 1;
    c->shadow_samples = 8;
    c->ao_enabled = 0;
    c->ao_samples = 8;
    c->ao_radius = 3.0f;
    c->ao_intensity = 0.5f;
    c->dof_enabled = 0;
    c->aperture = 0.5f;
    c->focus_distance = 0.0f;
    c->gradient_bg = 1;
    c->gradient_scale = 1.0f;
    const float center[4] = {0.91f, 0.89f, 0.86f, 1.0f}, edge[4] = {0.56f, 0.63f, 0.71f, 1.0f};
    std::memcpy(c->bg_center, center, sizeof(center));
    std::memcpy(c->bg_edge, edge, sizeof(edge));
    // shading.h:9-14
    c->kd = 0.75f;
    c->ks = 0.15f;
    c->ambient = 0.20f;
    c->shininess = 16.0f;
}

extern "C" int32_t mcskin_generate_tiles(int32_t w, int32_t h, int32_t ts, McTile* out, int32_t capacity) {
    if (w <= 0 || h <= 0 || ts <= 0) return 0;  // tile_renderer.cpp:19-21
    const long long cols = (static_cast<long long>(w) + ts - 1) / ts, rows = (static_cast<long long>(h) + ts - 1) / ts;
    const long long total = cols * rows;
    if (total > std::numeric_limits<int32_t>::max()) return 0;
    if (out) {
        long long n = 0;
        for (long long ty = 0; ty < rows && n < capacity; ++ty)
            for (long long tx = 0; tx < cols && n < capacity; ++tx, ++n) {
                McTile t;
                t.x = static_cast<int32_t>(tx * ts);
                t.y = static_cast<int32_t>(ty * ts);
                t.width = ts < w - t.x ? ts : w - t.x;
                t.height = ts < h - t.y ? ts : h - t.y;
                out[n] = t;
            }
    }
    return static_cast<int32_t>(total);
}

int prepare_frame(const McScene* scene, const McConfig* cfgIn, int useConfig, float aspectOverride,
                  PreparedFrame& out, std::string& error, const ExternalTexels* ext) {
    if (!scene) {
        error = "scene is null";
        return MC_ERR_INVALID;
    }
    if (scene->n_boxes < 0 || scene->n_texels < 0 || (scene->n_boxes > 0 && !scene->boxes) ||
        (scene->n_texels > 0 && !scene->texels_rgba && !ext)) {
        error = "scene has negative counts or null arrays";
        return MC_ERR_INVALID;
    }
    McConfig cfg;
    if (cfgIn) cfg = *cfgIn;
    else mcskin_config_defaults(&cfg);
    if (cfg.rng_mode != MC_RNG_MT19937 && cfg.rng_mode != MC_RNG_COUNTER) {
        error = "config: rng_mode must be 0 (mt19937) or 1 (counter-based)";
        return MC_ERR_INVALID;
    }

    DevFrame& f = out.frame;
    std::memset(&f, 0, sizeof(f));
    f.width = cfg.width;
    f.height = cfg.height;
    f.spp = cfg.samples_per_pixel > 1 ? cfg.samples_per_pixel : 1;  // tile_renderer.cpp:76
    f.max_bounces = cfg.max_bounces;
    f.tile_size = cfg.tile_size;
    if (cfg.width > 0 && cfg.height > 0 && cfg.tile_size > 0) {
        f.tiles_x = (cfg.width + cfg.tile_size - 1) / cfg.tile_size;
        f.tiles_y = (cfg.height + cfg.tile_size - 1) / cfg.tile_size;
    }
    f.inv_spp = 1.0f / static_cast<float>(f.spp);
    f.width_f = static_cast<float>(cfg.width);
    f.height_f = static_cast<float>(cfg.height);
    f.uv_recip = (cfg.width >= 1 && cfg.width <= 65535 && cfg.height >= 1 && cfg.height <= 65535) ? 1 : 0;
    f.inv_width_f = f.uv_recip ? 1.0f / f.width_f : 0.0f;
    f.inv_height_f = f.uv_recip ? 1.0f / f.height_f : 0.0f;
    f.aspect = aspectOverride > 0.0f ? aspectOverride
                                     : static_cast<float>(cfg.width) / static_cast<float>(cfg.height);

    // camera.cpp:10-16
    const H3 pos{scene->cam_pos[0], scene->cam_pos[1], scene->cam_pos[2]};
    const H3 tgt{scene->cam_target[0], scene->cam_target[1], scene->cam_target[2]};
    const H3 upv{scene->cam_up[0], scene->cam_up[1], scene->cam_up[2]};
    const H3 fwd = normh(sub(tgt, pos));
    const H3 right = normh(crossh(fwd, upv));
    const H3 trueUp = crossh(right, fwd);
    const float halfH = std::tan(scene->cam_fov_deg * 0.5f * static_cast<float>(M_PI) / 180.0f);
    const float halfW = halfH * f.aspect;
    f.cam_pos[0] = pos.x; f.cam_pos[1] = pos.y; f.cam_pos[2] = pos.z;
    f.cam_fwd[0] = fwd.x; f.cam_fwd[1] = fwd.y; f.cam_fwd[2] = fwd.z;
    f.cam_right[0] = right.x; f.cam_right[1] = right.y; f.cam_right[2] = right.z;
    f.cam_up[0] = trueUp.x; f.cam_up[1] = trueUp.y; f.cam_up[2] = trueUp.z;
    f.half_w = halfW;
    f.half_h = halfH;

    // tile_renderer.cpp:82-85, :99
    f.dof_on = (cfg.dof_enabled && cfg.aperture > 1e-6f) ? 1 : 0;
    f.aperture = cfg.aperture;
    f.focus_dist = cfg.focus_distance;
    if (f.focus_dist <= 0.0f) f.focus_dist = lenh(sub(tgt, pos));
    f.draws_per_sample = (f.spp > 1 ? 2 : 0) + (f.dof_on ? 2 : 0);

    f.n_boxes = scene->n_boxes;
    std::memcpy(f.light_pos, scene->light_pos, sizeof(f.light_pos));
    std::memcpy(f.light_color, scene->light_color, sizeof(f.light_color));
    f.light_radius = scene->light_radius;
    std::memcpy(f.background, scene->background, sizeof(f.background));

    f.rng_mode = cfg.rng_mode == MC_RNG_COUNTER ? 1 : 0;
    f.use_config = useConfig ? 1 : 0;
    f.soft_on = (cfg.soft_shadows && cfg.shadow_samples > 1) ? 1 : 0;  // raytracer.cpp:109
    f.shadow_samples = cfg.shadow_samples;
    f.ao_on = cfg.ao_enabled ? 1 : 0;
    f.ao_samples = cfg.ao_samples;
    f.ao_radius = cfg.ao_radius;
    f.ao_intensity = cfg.ao_intensity;
    f.gradient_bg = cfg.gradient_bg ? 1 : 0;
    f.gradient_scale = cfg.gradient_scale;
    std::memcpy(f.bg_center, cfg.bg_center, sizeof(f.bg_center));
    std::memcpy(f.bg_edge, cfg.bg_edge, sizeof(f.bg_edge));
    f.kd = cfg.kd;
    f.ks = cfg.ks;
    f.ambient = cfg.ambient;
    f.shininess = cfg.shininess;

    // texel pool + the two synthetic 1x1 textures
    const int magentaTexel = scene->n_texels;      // null Triangle::texture (intersection.cpp:303-306)
    const int blankTexel = scene->n_texels + 1;    // empty TextureRegion -> Color() (texture_region.h:20-22)
    if (ext) {  // the pool is filled on the device; only the two synthetic texels that follow it live here
        out.texels.resize(2);
        out.texels[0] = {1.0f, 0.0f, 1.0f, 1.0f};
        out.texels[1] = {0.0f, 0.0f, 0.0f, 1.0f};
    } else {
        out.texels.resize(static_cast<size_t>(scene->n_texels) + 2);
        if (scene->n_texels > 0) std::memcpy(out.texels.data(), scene->texels_rgba, sizeof(float4h) * scene->n_texels);
        out.texels[magentaTexel] = {1.0f, 0.0f, 1.0f, 1.0f};
        out.texels[blankTexel] = {0.0f, 0.0f, 0.0f, 1.0f};
    }

    out.boxes.resize(scene->n_boxes);
    std::vector<std::array<float, 3>> rejectLo(scene->n_boxes), rejectHi(scene->n_boxes);
    double cullLo[3] = {1e300, 1e300, 1e300}, cullHi[3] = {-1e300, -1e300, -1e300};
    bool anyBox = false;
    for (int b = 0; b < scene->n_boxes; ++b) {
        const McBox& src = scene->boxes[b];
        DevBox& d = out.boxes[b];
        std::memset(&d, 0, sizeof(d));
        uint32_t flags = 0;
        if (src.is_outer_layer) flags |= kBoxOuter;
        if (src.n_triangles <= 0) flags |= kBoxEmpty;
        for (int k = 0; k < 3; ++k) {
            d.lo[k] = src.bounds_min[k];
            d.hi[k] = src.bounds_max[k];
            const float s = src.bounds_max[k] - src.bounds_min[k];
            d.size[k] = (s > 1e-8f) ? s : 1.0f;
            d.pivot[k] = src.pivot[k];
        }
        {
            // Quotients by size[] are formed on the device as two Newton steps on a*RN(1/size) with
            // exact (fused) residuals, which is the correctly rounded a/size unless size's
            // significand is all ones (Markstein); also stay clear of the extreme exponents.
            bool ok = true;
            for (int k = 0; k < 3; ++k) {
                uint32_t bits;
                std::memcpy(&bits, &d.size[k], sizeof(bits));
                const uint32_t expo = (bits >> 23) & 0xffu;
                if ((bits & 0x7fffffu) == 0x7fffffu || expo < 64u || expo > 190u) ok = false;
            }
            d.inv_size_x = ok ? 1.0f / d.size[0] : 0.0f;
            d.inv_size_y = ok ? 1.0f / d.size[1] : 0.0f;
            d.inv_size_z = ok ? 1.0f / d.size[2] : 0.0f;
            if (ok) flags |= kBoxRecip;
        }
        d.inv_cx = d.inv_cz = d.fwd_cx = d.fwd_cz = 1.0f;
        if (src.has_rotation) {
            flags |= kBoxRotated;
            // the inverse passes -rot (intersection.cpp:388-391); |−x| > 0.01 is the same test
            if (std::fabs(src.rot_x_deg) > 0.01f) flags |= kBoxRotX;
            if (std::fabs(src.rot_z_deg) > 0.01f) flags |= kBoxRotZ;
            const float ix = radians(-src.rot_x_deg), iz = radians(-src.rot_z_deg);
            const float fx = radians(src.rot_x_deg), fz = radians(src.rot_z_deg);
            d.inv_cx = std::cos(ix); d.inv_sx = std::sin(ix);
            d.inv_cz = std::cos(iz); d.inv_sz = std::sin(iz);
            d.fwd_cx = std::cos(fx); d.fwd_sx = std::sin(fx);
            d.fwd_cz = std::cos(fz); d.fwd_sz = std::sin(fz);
        }
        d.flags = flags;
        if ((flags & kBoxRotated) && !(flags & kBoxEmpty)) f.any_rotated = 1;
        if (b < 32) {
            if (!(flags & kBoxEmpty)) f.usable_mask |= 1u << b;
            if ((flags & kBoxRotated) && !(flags & kBoxEmpty)) f.rotated_mask |= 1u << b;
        }
        bool opaque = true;  // no texel of any face has alpha == 0 (the pass-through rule, intersection.cpp:311)
        for (int k = 0; k < kFaceCount; ++k) {
            const McFaceTex& ft = src.face[k];
            int offset, w, h;
            if (ft.texel_offset < 0) {
                offset = magentaTexel; w = 1; h = 1;
            } else if (ft.width <= 0 || ft.height <= 0) {
                offset = blankTexel; w = 1; h = 1;
            } else {
                offset = ft.texel_offset; w = ft.width; h = ft.height;
                if (w > 32767 || h > 32767) {
                    error = "face texture larger than 32767 texels on a side";
                    return MC_ERR_LIMIT;
                }
                if (static_cast<long long>(offset) + static_cast<long long>(w) * h > scene->n_texels) {
                    error = "face texture window runs past the texel pool";
                    return MC_ERR_INVALID;
                }
            }
            d.face[k].x = offset;
            d.face[k].y = w | (h << 16);
            if (ext) {
                // (windows into the device-side pool; the synthetic texels are opaque)
                if (offset < scene->n_texels) opaque = opaque && ext->boxOpaque[b] != 0;
            } else {
                for (long long t = 0; opaque && t < static_cast<long long>(w) * h; ++t)
                    if (out.texels[static_cast<size_t>(offset + t)].w == 0.0f) opaque = false;
            }
        }
        // an unposed box without see-through texels is hit by exactly the rays that pass the slab test, at
        // the slab distance: occlusion queries need no face / texel evaluation for it (occluded_among)
        if (opaque && !src.has_rotation && !(flags & kBoxEmpty)) {
            flags |= kBoxOpaque;
            d.flags = flags;
            if (b < 32) f.opaque_mask |= 1u << b;
        }
        // (posed: no exact shortcut, but a ray well inside such a box needs no exact evaluation to be called occluded)
        if (opaque && src.has_rotation && !(flags & kBoxEmpty) && b < 32) f.opaque_posed_mask |= 1u << b;
        // conservative world bounds of this box (posed boxes: rotate the 8 corners in double)
        double boxLo[3] = {1e300, 1e300, 1e300}, boxHi[3] = {-1e300, -1e300, -1e300};
        if (!(flags & kBoxEmpty)) {
            anyBox = true;
            for (int corner = 0; corner < 8; ++corner) {
                double p[3] = {(corner & 1) ? d.hi[0] : d.lo[0], (corner & 2) ? d.hi[1] : d.lo[1],
                               (corner & 4) ? d.hi[2] : d.lo[2]};
                if (src.has_rotation) {
                    double x = p[0] - d.pivot[0], y = p[1] - d.pivot[1], z = p[2] - d.pivot[2];
                    if (flags & kBoxRotX) {
                        const double a = static_cast<double>(src.rot_x_deg) * M_PI / 180.0, c = std::cos(a), s = std::sin(a);
                        const double ny = y * c - z * s, nz = y * s + z * c;
                        y = ny; z = nz;
                    }
                    if (flags & kBoxRotZ) {
                        const double a = static_cast<double>(src.rot_z_deg) * M_PI / 180.0, c = std::cos(a), s = std::sin(a);
                        const double nx = x * c - y * s, ny = x * s + y * c;
                        x = nx; y = ny;
                    }
                    p[0] = x + d.pivot[0]; p[1] = y + d.pivot[1]; p[2] = z + d.pivot[2];
                }
                for (int k = 0; k < 3; ++k) {
                    if (p[k] < cullLo[k]) cullLo[k] = p[k];
                    if (p[k] > cullHi[k]) cullHi[k] = p[k];
                    if (p[k] < boxLo[k]) boxLo[k] = p[k];
                    if (p[k] > boxHi[k]) boxHi[k] = p[k];
                }
            }
        }
        // Bounds the reject pass tests.  Unposed box: the box itself — the pass then compares
        // exactly the floats the reference compares.  Posed box: a world-space box around its
        // rotated corners, inflated far beyond any rounding of the reference's local-space test
        // (which the exact evaluation repeats for every survivor), so rejection stays conservative.
        for (int k = 0; k < 3; ++k) {
            float lo = d.lo[k], hi = d.hi[k];
            if (src.has_rotation && !(flags & kBoxEmpty)) {
                const double margin = 1e-3 * (std::fabs(boxLo[k]) + std::fabs(boxHi[k]) + (boxHi[k] - boxLo[k])) + 1e-2;
                lo = static_cast<float>(boxLo[k] - margin);
                hi = static_cast<float>(boxHi[k] + margin);
                if (!std::isfinite(lo) || !std::isfinite(hi)) {  // degenerate pose: never pre-reject
                    lo = -std::numeric_limits<float>::max();
                    hi = std::numeric_limits<float>::max();
                }
            }
            rejectLo[b][k] = lo;
            rejectHi[b][k] = hi;
        }
    }
    // shared-memory image of the boxes
    {
        const SceneBlobLayout lay(scene->n_boxes);
        out.blob.assign(std::max<size_t>(16, lay.bytes()), 0);
        float* lo = reinterpret_cast<float*>(out.blob.data() + lay.loOffset());
        float* hi = reinterpret_cast<float*>(out.blob.data() + lay.hiOffset());
        // Two-level reject pass: box i is ENCLOSED by box j when its reject bounds lie inside j's with 1e-3 to spare on
        // every side (an inner body part inside its outer layer: 0.5) — a ray the slab test of j rejects then misses
        // i in any float arithmetic.  Enclosing boxes ("roots") must not be enclosed themselves; every enclosed box
        // hangs on one root.  Only among the first 32 boxes, and only when it saves tests.
        std::vector<uint32_t> children(scene->n_boxes, 0u);
        if (scene->n_boxes <= 32) {
            std::vector<int> parent(scene->n_boxes, -1);
            auto inside = [&](int i, int j) {
                for (int k = 0; k < 3; ++k)
                    if (!(rejectLo[i][k] >= rejectLo[j][k] + 1e-3f && rejectHi[i][k] <= rejectHi[j][k] - 1e-3f)) return false;
                return true;
            };
            auto usable = [&](int b) { return !(out.boxes[b].flags & kBoxEmpty); };
            for (int i = 0; i < scene->n_boxes; ++i) {
                if (!usable(i)) continue;
                for (int j = 0; j < scene->n_boxes && parent[i] < 0; ++j)
                    if (j != i && usable(j) && inside(i, j)) parent[i] = j;
            }
            // one level only: a parent that is itself enclosed hands its children to its own root
            for (int i = 0; i < scene->n_boxes; ++i) {
                int p = parent[i], guard = 0;
                while (p >= 0 && parent[p] >= 0 && guard++ < 64) p = parent[p];
                parent[i] = p;
            }
            int nRoots = 0, nEnclosed = 0;
            for (int i = 0; i < scene->n_boxes; ++i) {
                if (!usable(i)) continue;
                if (parent[i] < 0) { f.root_mask |= 1u << i; ++nRoots; }
                else { children[parent[i]] |= 1u << i; ++nEnclosed; }
            }
            if (nEnclosed < 2) {  // nothing to gain
                f.root_mask = 0u;
                std::fill(children.begin(), children.end(), 0u);
            }
            (void)nRoots;
        }
        for (int b = 0; b < scene->n_boxes; ++b) {
            const DevBox& d = out.boxes[b];
            std::memcpy(lo + 4 * b, rejectLo[b].data(), 3 * sizeof(float));
            std::memcpy(lo + 4 * b + 3, &d.flags, sizeof(uint32_t));
            std::memcpy(hi + 4 * b, rejectHi[b].data(), 3 * sizeof(float));
            std::memcpy(hi + 4 * b + 3, &children[b], sizeof(uint32_t));
        }
        if (scene->n_boxes > 0)
            std::memcpy(out.blob.data() + lay.boxOffset(), out.boxes.data(), sizeof(DevBox) * scene->n_boxes);
    }
    f.cull_valid = 0;
    if (anyBox) {
        bool finite = true;
        for (int k = 0; k < 3; ++k) {
            const double ext = cullHi[k] - cullLo[k];
            const double margin = 1e-3 * (std::fabs(cullLo[k]) + std::fabs(cullHi[k]) + ext) + 1e-2;
            f.cull_lo[k] = static_cast<float>(cullLo[k] - margin);
            f.cull_hi[k] = static_cast<float>(cullHi[k] + margin);
            if (!std::isfinite(f.cull_lo[k]) || !std::isfinite(f.cull_hi[k])) finite = false;
        }
        f.cull_valid = finite ? 1 : 0;
    } else {
        // nothing hittable: an empty box rejects every ray
        for (int k = 0; k < 3; ++k) { f.cull_lo[k] = 1.0f; f.cull_hi[k] = -1.0f; }
        f.cull_valid = 1;
    }
    // screen-space rectangle of a world-space box for pinhole rays: the pixels (inclusive, with a
    // 2-pixel margin) whose rays can reach it; false if a corner is not in front of the camera
    const bool canProject = !f.dof_on && cfg.width > 0 && cfg.height > 0 && halfW > 1e-6f && halfH > 1e-6f &&
                            lenh(right) > 0.5f && lenh(fwd) > 0.5f;
    auto project_box = [&](const float* lo, const float* hi, int* rect) {
        double x0 = 1e300, y0 = 1e300, x1 = -1e300, y1 = -1e300;
        for (int corner = 0; corner < 8; ++corner) {
            const double p[3] = {(corner & 1) ? hi[0] : lo[0], (corner & 2) ? hi[1] : lo[1], (corner & 4) ? hi[2] : lo[2]};
            const double rel[3] = {p[0] - pos.x, p[1] - pos.y, p[2] - pos.z};
            const double depth = rel[0] * fwd.x + rel[1] * fwd.y + rel[2] * fwd.z;
            if (!(depth > 1e-3)) return false;
            const double su = (rel[0] * right.x + rel[1] * right.y + rel[2] * right.z) / depth;
            const double sv = (rel[0] * trueUp.x + rel[1] * trueUp.y + rel[2] * trueUp.z) / depth;
            const double px = (su / halfW + 1.0) * 0.5 * cfg.width;
            const double py = (1.0 - (sv / halfH + 1.0) * 0.5) * cfg.height;
            x0 = std::min(x0, px); x1 = std::max(x1, px);
            y0 = std::min(y0, py); y1 = std::max(y1, py);
        }
        if (!(std::isfinite(x0) && std::isfinite(x1) && std::isfinite(y0) && std::isfinite(y1))) return false;
        const double big = 1e9;
        rect[0] = static_cast<int>(std::floor(std::max(-big, x0))) - 2;
        rect[1] = static_cast<int>(std::floor(std::max(-big, y0))) - 2;
        rect[2] = static_cast<int>(std::ceil(std::min(big, x1))) + 2;
        rect[3] = static_cast<int>(std::ceil(std::min(big, y1))) + 2;
        return true;
    };
    f.rect_valid = 0;
    if (f.cull_valid && canProject) {
        if (!anyBox) {
            f.rect_valid = 1;  // nothing to hit: an empty rectangle
            f.rect_x0 = f.rect_y0 = 1;
            f.rect_x1 = f.rect_y1 = 0;
        } else {
            int r[4];
            if (project_box(f.cull_lo, f.cull_hi, r)) {
                f.rect_valid = 1;
                f.rect_x0 = r[0]; f.rect_y0 = r[1]; f.rect_x1 = r[2]; f.rect_y1 = r[3];
            }
        }
    }
    // and of every box (its reject bounds, inflated like the cull box)
    f.box_rects_valid = 0;
    if (f.rect_valid && scene->n_boxes > 0 && scene->n_boxes <= 32) {
        const SceneBlobLayout lay(scene->n_boxes);
        int* rects = reinterpret_cast<int*>(out.blob.data() + lay.rectOffset());
        bool ok = true;
        for (int b = 0; b < scene->n_boxes && ok; ++b) {
            int* r = rects + 4 * b;
            if (out.boxes[b].flags & kBoxEmpty) {
                r[0] = r[1] = 1; r[2] = r[3] = 0;
                continue;
            }
            float lo[3], hi[3];
            for (int k = 0; k < 3; ++k) {
                const double l = rejectLo[b][k], h = rejectHi[b][k];
                const double margin = 1e-3 * (std::fabs(l) + std::fabs(h) + (h - l)) + 1e-2;
                lo[k] = static_cast<float>(l - margin);
                hi[k] = static_cast<float>(h + margin);
                if (!std::isfinite(lo[k]) || !std::isfinite(hi[k])) ok = false;
            }
            if (ok) ok = project_box(lo, hi, r);
        }
        f.box_rects_valid = ok ? 1 : 0;
    }
    return MC_OK;
}

}  // namespace mcskin

// Cost-balanced deal of a frame's tiles (see mcskin_cuda.h).  The weight of a tile: 1 for its primary pass plus
// kCoveredCost times the fraction of it that box rectangles cover (every box counted: overlapping inner and
// outer layers do mean more work), or that the figure's rectangle covers when the boxes have none.
extern "C" int32_t mcskin_partition_tiles(const McScene* scene, const McConfig* cfg, int32_t nParts, int32_t part,
                                          int32_t rootPart, int32_t* outTiles, int32_t capacity) {
    using namespace mcskin;
    if (!scene || !cfg || nParts <= 0 || part < 0 || part >= nParts || capacity < 0 || rootPart >= nParts) {
        set_last_error("partition_tiles: bad argument");
        return MC_ERR_INVALID;
    }
    if (cfg->width <= 0 || cfg->height <= 0 || cfg->tile_size <= 0) return 0;
    PreparedFrame pf;
    std::string err;
    const int rc = prepare_frame(scene, cfg, 1, 0.0f, pf, err);
    if (rc != MC_OK) {
        set_last_error(err);
        return rc;
    }
    const DevFrame& f = pf.frame;
    const int ts = f.tile_size;
    const long long nTiles = static_cast<long long>(f.tiles_x) * f.tiles_y;
    if (nTiles > 0x7fffffff) {
        set_last_error("partition_tiles: too many tiles");
        return MC_ERR_LIMIT;
    }
    constexpr double kCoveredCost = 50.0;
    std::vector<std::array<int, 4>> rects;
    if (f.box_rects_valid) {
        const SceneBlobLayout lay(f.n_boxes);
        const int* r = reinterpret_cast<const int*>(pf.blob.data() + lay.rectOffset());
        for (int b = 0; b < f.n_boxes; ++b)
            if (!(pf.boxes[b].flags & kBoxEmpty)) rects.push_back({r[4 * b], r[4 * b + 1], r[4 * b + 2], r[4 * b + 3]});
    } else if (f.rect_valid) {
        rects.push_back({f.rect_x0, f.rect_y0, f.rect_x1, f.rect_y1});
    } else {
        rects.push_back({0, 0, f.width - 1, f.height - 1});  // any pixel may hit: cost follows the tile's area
    }
    // Tiles the figure's screen rectangle touches ("hot": the kernels write them through BandView::hot_*) are dealt
    // one by one, heaviest first, to the least loaded part.  The others cost the same (their primary pass) and are
    // final after it: they are dealt as CONTIGUOUS runs in frame order, so that a part's background pixels are a
    // few rectangles (one DMA each on the way to a host frame) — each part gets the run that fills it up to the mean.
    std::vector<std::pair<double, int>> hot;
    std::vector<int> light;
    double total = 0.0;
    for (int id = 0; id < nTiles; ++id) {
        const int tileY = id / f.tiles_x, tileX = id - tileY * f.tiles_x;
        const int left = tileX * ts, top = tileY * ts;
        const int right = std::min(f.width, left + ts) - 1, bottom = std::min(f.height, top + ts) - 1;
        double covered = 0.0;
        for (const auto& r : rects) {
            const long long w = static_cast<long long>(std::min(right, r[2])) - std::max(left, r[0]) + 1;
            const long long h = static_cast<long long>(std::min(bottom, r[3])) - std::max(top, r[1]) + 1;
            if (w > 0 && h > 0) covered += static_cast<double>(w) * static_cast<double>(h);
        }
        const double area = static_cast<double>(right - left + 1) * (bottom - top + 1);
        const double weight = area / (static_cast<double>(ts) * ts) + kCoveredCost * covered / (static_cast<double>(ts) * ts);
        total += weight;
        // the primary kernels' own test (tileCanHit)
        const bool canHit = !f.rect_valid || !(left > f.rect_x1 || right < f.rect_x0 || top > f.rect_y1 || bottom < f.rect_y0);
        if (canHit) hot.push_back({weight, id});
        else light.push_back(id);
    }
    std::stable_sort(hot.begin(), hot.end(), [](const std::pair<double, int>& a, const std::pair<double, int>& b) {
        return a.first > b.first;  // heaviest first; equal weights keep frame order
    });
    std::vector<double> load(nParts, 0.0);
    std::vector<int32_t> mine;
    for (const auto& t : hot) {
        int best = 0;
        for (int p = 1; p < nParts; ++p)
            if (load[p] < load[best]) best = p;
        load[best] += t.first;
        if (best == part) mine.push_back(t.second);
    }
    {
        auto weight_of = [&](int id) {
            const int tileY = id / f.tiles_x, tileX = id - tileY * f.tiles_x;
            return static_cast<double>(std::min(f.width, (tileX + 1) * ts) - tileX * ts) *
                   (std::min(f.height, (tileY + 1) * ts) - tileY * ts) / (static_cast<double>(ts) * ts);
        };
        // What a background tile costs a part: 1, or kRemoteLight when its pixels go to ANOTHER device's memory (the
        // parts other than rootPart, when one device holds the frame and the others store into it over NVLink: a
        // background tile is little more than its stores, and remote stores measured ~30 % dearer on B200).
        constexpr double kRemoteLight = 1.3;
        auto cost_of = [&](int p) { return (rootPart >= 0 && p != rootPart) ? kRemoteLight : 1.0; };
        double lightTotal = 0.0;
        for (int id : light) lightTotal += weight_of(id);
        // the level T every part is filled to: sum over parts of max(0, T - load) / cost = the background tiles
        double lo = 0.0, hi = 0.0;
        for (int p = 0; p < nParts; ++p) hi = std::max(hi, load[p]);
        hi += lightTotal * kRemoteLight + 1.0;
        for (int it = 0; it < 100; ++it) {
            const double T = 0.5 * (lo + hi);
            double fit = 0.0;
            for (int p = 0; p < nParts; ++p) fit += std::max(0.0, T - load[p]) / cost_of(p);
            (fit < lightTotal ? lo : hi) = T;
        }
        const double T = hi;
        int p = 0;
        double room = std::max(0.0, T - load[0]) / cost_of(0);
        for (int id : light) {
            const double w = weight_of(id);
            while (p < nParts - 1 && room < 0.5 * w) {  // this part is full: its run ends here
                ++p;
                room = std::max(0.0, T - load[p]) / cost_of(p);
            }
            room -= w;
            if (p == part) mine.push_back(id);
        }
    }
    std::sort(mine.begin(), mine.end());
    if (outTiles)
        for (size_t i = 0; i < mine.size() && i < static_cast<size_t>(capacity); ++i) outTiles[i] = mine[i];
    return static_cast<int32_t>(mine.size());
}

// Host model of the device's sincos_ref (dev_shade.cuh): glibc 2.39 sincosf as built for x86-64 with
// FMA.  Same operations in the same order; std::fma is a correctly rounded fused multiply-add
// whether or not this translation unit is compiled with -mfma.
extern "C" void mcskin_sincos_model(const float* angles, int32_t n, float* outSin, float* outCos) {
    for (int32_t i = 0; i < n; ++i) {
        const float a = angles[i];
        uint32_t bits;
        std::memcpy(&bits, &a, sizeof(bits));
        const uint32_t top12 = (bits >> 20) & 0x7ffu;
        if (top12 > 0x42eu) {
            outSin[i] = std::sin(a);
            outCos[i] = std::cos(a);
            continue;
        }
        if (top12 < 0x398u) {
            outSin[i] = a;
            outCos[i] = 1.0f;
            continue;
        }
        double x = static_cast<double>(a), xs = x;
        int nq = 0;
        if (top12 >= 0x3f4u) {
            const double r = x * 0x1.45F306DC9C883p+23;
            nq = (static_cast<int32_t>(r) + 0x800000) >> 24;
            x = std::fma(-static_cast<double>(nq), 0x1.921FB54442D18p0, x);
            xs = ((nq + 1) & 2) ? -x : x;
        }
        const bool flipCos = (nq & 2) != 0;
        const double c0 = flipCos ? -0x1p0 : 0x1p0;
        const double c1 = flipCos ? 0x1.ffffffd0c621cp-2 : -0x1.ffffffd0c621cp-2;
        const double c2 = flipCos ? -0x1.55553e1068f19p-5 : 0x1.55553e1068f19p-5;
        const double c3 = flipCos ? 0x1.6c087e89a359dp-10 : -0x1.6c087e89a359dp-10;
        const double c4 = flipCos ? -0x1.99343027bf8c3p-16 : 0x1.99343027bf8c3p-16;
        const double s1 = -0x1.555545995a603p-3, s2 = 0x1.1107605230bc4p-7, s3 = -0x1.994eb3774cf24p-13;
        // volatile-free, but every product below is a separate statement: this file is built with
        // -ffp-contract=off, so only the std::fma calls fuse
        const double x2 = x * x;
        const double x3 = xs * x2, x4 = x2 * x2;
        const double sq = std::fma(x2, s3, s2), cq = std::fma(x2, c4, c3), cl = std::fma(x2, c1, c0);
        const double x5 = x3 * x2, x6 = x4 * x2;
        const double sl = std::fma(x3, s1, xs), cm = std::fma(x4, c2, cl);
        const float fs = static_cast<float>(std::fma(x5, sq, sl));
        const float fc = static_cast<float>(std::fma(x6, cq, cm));
        outSin[i] = (nq & 1) ? fc : fs;
        outCos[i] = (nq & 1) ? fs : fc;
    }
}

// Host model of the device's powf_ref (dev_shade.cuh): glibc 2.39 powf as built for x86-64 with FMA.
namespace {
const double kPowLog2Tab[16][2] = {
    {0x1.661ec79f8f3bep+0, -0x1.efec65b963019p-2}, {0x1.571ed4aaf883dp+0, -0x1.b0b6832d4fca4p-2},
    {0x1.49539f0f010b0p+0, -0x1.7418b0a1fb77bp-2}, {0x1.3c995b0b80385p+0, -0x1.39de91a6dcf7bp-2},
    {0x1.30d190c8864a5p+0, -0x1.01d9bf3f2b631p-2}, {0x1.25e227b0b8ea0p+0, -0x1.97c1d1b3b7af0p-3},
    {0x1.1bb4a4a1a343fp+0, -0x1.2f9e393af3c9fp-3}, {0x1.12358f08ae5bap+0, -0x1.960cbbf788d5cp-4},
    {0x1.0953f419900a7p+0, -0x1.a6f9db6475fcep-5}, {0x1.0000000000000p+0, 0x0.0p+0},
    {0x1.e608cfd9a47acp-1, 0x1.338ca9f24f53dp-4},  {0x1.ca4b31f026aa0p-1, 0x1.476a9543891bap-3},
    {0x1.b2036576afce6p-1, 0x1.e840b4ac4e4d2p-3},  {0x1.9c2d163a1aa2dp-1, 0x1.40645f0c6651cp-2},
    {0x1.886e6037841edp-1, 0x1.88e9c2c1b9ff8p-2},  {0x1.767dcf5534862p-1, 0x1.ce0a44eb17bccp-2}};
const uint64_t kPowExp2Tab[32] = {
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,
    0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,
    0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,
    0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,
    0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,
    0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,
    0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull};
uint32_t bits_of(float f) { uint32_t u; std::memcpy(&u, &f, sizeof(u)); return u; }
float float_of(uint32_t u) { float f; std::memcpy(&f, &u, sizeof(f)); return f; }
uint64_t bits_of(double f) { uint64_t u; std::memcpy(&u, &f, sizeof(u)); return u; }
float powf_model(float x, float y) {
    uint32_t ix = bits_of(x);
    const uint32_t iy = bits_of(y);
    if (2u * iy - 1u > 0xfefffffeu) return std::pow(x, y);
    if (ix - 0x00800000u > 0x7effffffu) {
        if (ix == 0u || ix >= 0x00800000u) return std::pow(x, y);
        ix = bits_of(x * 0x1p23f) & 0x7fffffffu;
        ix -= 23u << 23;
    }
    const uint32_t tmp = ix - 0x3f330000u;
    const int i = (tmp >> 19) & 15;
    const uint32_t top = tmp & 0xff800000u;
    const int k = static_cast<int32_t>(top) >> 23;
    const double z = static_cast<double>(float_of(ix - top));
    const double r = std::fma(z, kPowLog2Tab[i][0], -1.0);
    const double y0 = kPowLog2Tab[i][1] + static_cast<double>(k);
    const double r2 = r * r;
    double p = std::fma(0x1.27616c9496e0bp-2, r, -0x1.71969a075c67ap-2);
    const double p1 = std::fma(0x1.ec70a6ca7baddp-2, r, -0x1.7154748bef6c8p-1);
    double q = std::fma(0x1.71547652ab82bp+0, r, y0);
    const double r4 = r2 * r2;
    q = std::fma(p1, r2, q);
    p = std::fma(p, r4, q);
    const double ylogx = static_cast<double>(y) * p;
    if (((bits_of(ylogx) >> 47) & 0xffffull) > 0x80beull) {
        if (ylogx > 0x1.fffffffa3aae2p+6) return std::pow(x, y);
        if (ylogx <= -150.0) return 0.0f;
        if (ylogx < -149.0) return float_of(1u);
    }
    double kd = ylogx + 0x1.8p+47;
    const uint64_t ki = bits_of(kd);
    kd -= 0x1.8p+47;
    const double rr = ylogx - kd;
    const uint64_t sb = kPowExp2Tab[ki & 31ull] + (ki << 47);
    double s;
    std::memcpy(&s, &sb, sizeof(s));
    const double zz = std::fma(0x1.c6af84b912394p-5, rr, 0x1.ebfce50fac4f3p-3);
    const double rr2 = rr * rr;
    double e = std::fma(0x1.62e42ff0c52d6p-1, rr, 1.0);
    e = std::fma(zz, rr2, e);
    return static_cast<float>(e * s);
}
}  // namespace

extern "C" void mcskin_powf_model(const float* x, const float* y, int32_t n, float* out) {
    for (int32_t i = 0; i < n; ++i) out[i] = powf_model(x[i], y[i]);
}
