// dev_types.cuh — device-side layout of one frame's read-only inputs.
//
// Everything the kernels read per frame, pre-digested on the host by
// scene_prep.cpp with the host's glibc so that every libm call the reference makes
// per ray with frame-constant arguments (tanf of the fov, cosf/sinf of pose
// angles; camera.cpp:15, intersection.cpp:17-19,27-29) yields the bit-identical
// value on the device.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>  // int2, __align__, __host__ __device__

// The kernels are built twice: MCSKIN_POSED=1 (kernels.cu, wavefront.cu), the general form, and MCSKIN_POSED=0
// (kernels_plain.cu, wavefront_plain.cu, into namespace mcskin::plain) for scenes in which no box is posed — the
// standing figure of every BASELINE config — where the pose transform of the exact box test, the box-space second
// opinions on posed boxes and what they cost in registers are compiled out.  The host picks the build per scene
// (DevFrame::any_rotated).
#ifndef MCSKIN_POSED
#define MCSKIN_POSED 1
#endif
// A third build (kernels_counter.cu, wavefront_counter.cu, namespace mcskin::counter; the general, pose-capable form)
// replaces every std::mt19937 stream by the counter-based streams of McConfig::rng_mode 1 (dev_mt19937.cuh).
#ifndef MCSKIN_COUNTER_RNG
#define MCSKIN_COUNTER_RNG 0
#endif
#if MCSKIN_COUNTER_RNG
#define MCSKIN_VARIANT_BEGIN namespace counter {
#define MCSKIN_VARIANT_END }
#define MCSKIN_VARIANT_NS ::mcskin::counter
#elif MCSKIN_POSED
#define MCSKIN_VARIANT_BEGIN
#define MCSKIN_VARIANT_END
#define MCSKIN_VARIANT_NS ::mcskin
#else
#define MCSKIN_VARIANT_BEGIN namespace plain {
#define MCSKIN_VARIANT_END }
#define MCSKIN_VARIANT_NS ::mcskin::plain
#endif

namespace mcskin {

// memo of the engines' seeding (FreshStream::seed_memo): seeds in [-2^20, 2^23 - 2^20), 64 MB
constexpr uint32_t kSeedMemoEntries = 1u << 23;
constexpr uint32_t kSeedMemoOffset = 1u << 20;
constexpr uint32_t kSeedMemoTag = 0x5bd1e995u;  // not a seed of the window: an all-zero entry is never valid


constexpr bool kPosedScenes = MCSKIN_POSED != 0;
constexpr bool kCounterRng = MCSKIN_COUNTER_RNG != 0;
constexpr int kFaceCount = 6;

// One Mesh == one box (intersection.cpp:200-406).  144 bytes, 16-byte aligned so a
// warp-uniform read is a handful of broadcast LDG.128s.
struct __align__(16) DevBox {
    float lo[3];      // bounds_min
    uint32_t flags;   // kBox* bits
    float hi[3];      // bounds_max
    float inv_size_x; // RN(1 / size[k]) (host division): operands of the exact quotients in face_texel,
    float size[3];    // hi-lo per axis, replaced by 1 where <= 1e-8 (intersection.cpp:141-143)
    float inv_size_y; //   valid when kBoxRecip is set
    float pivot[3];
    float inv_size_z;
    float inv_cx, inv_sx, inv_cz, inv_sz;  // cosf/sinf of rad(-rotX), rad(-rotZ): world -> local
    float fwd_cx, fwd_sx, fwd_cz, fwd_sz;  // cosf/sinf of rad(+rotX), rad(+rotZ): local -> world
    // face f: x = first texel in the pool, y = width | (height << 16)
    int2 face[kFaceCount];
};

enum : uint32_t {
    kBoxOuter = 1u,      // Mesh::isOuterLayer
    kBoxRotated = 2u,    // Mesh::hasRotation
    kBoxRotX = 4u,       // |rotX| > 0.01 (intersection.cpp:16)
    kBoxRotZ = 8u,       // |rotZ| > 0.01 (intersection.cpp:26)
    kBoxEmpty = 16u,     // no triangles: never hit (intersection.cpp:205)
    kBoxRecip = 32u,     // inv_size_* may replace the divisions by size[] (see div_exact)
    kBoxOpaque = 64u     // unposed, and no texel of any face has alpha == 0: every ray that passes the slab test hits
};

// The scene blob: what a CTA stages into shared memory with one bulk copy.
//   [ float4 lo[nPad] ][ float4 hi[nPad] ][ int4 rect[nPad] ][ DevBox boxes[n] ]      nPad = n rounded up to 4
// lo[i] = (bounds_min, flags as bit pattern), hi[i] = (bounds_max, bit mask of the boxes box i encloses): the compact
// operands of the reject pass; rect[i] = (x0, y0, x1, y1), the pixels (inclusive, 2-pixel margin) a pinhole
// ray must pass through to reach box i (valid when DevFrame::box_rects_valid); the full records serve the
// exact evaluation.
struct SceneBlobLayout {
    int n, nPad;
    __host__ __device__ explicit SceneBlobLayout(int nBoxes) : n(nBoxes), nPad((nBoxes + 3) & ~3) {}
    __host__ __device__ unsigned loOffset() const { return 0u; }
    __host__ __device__ unsigned hiOffset() const { return 16u * nPad; }
    __host__ __device__ unsigned rectOffset() const { return 32u * nPad; }
    __host__ __device__ unsigned boxOffset() const { return 48u * nPad; }
    __host__ __device__ unsigned bytes() const { return 48u * nPad + static_cast<unsigned>(sizeof(DevBox)) * n; }
};

// Frame constants, passed by value as a kernel parameter (constant bank).
struct DevFrame {
    // image / sampling
    int width, height, spp, max_bounces;
    int tile_size, tiles_x, tiles_y;
    int draws_per_sample;     // 0 (spp==1, no DOF), 2 (jitter or lens), 4 (jitter + lens)
    float inv_spp;            // 1.0f / float(spp) (tile_renderer.cpp:122)
    float width_f, height_f, aspect;
    // 1.0f / width_f, 1.0f / height_f rounded to nearest on the host: operands of the exact
    // quotient in sample_uv (valid when uv_recip != 0, i.e. both sizes are in 1..65535)
    float inv_width_f, inv_height_f;
    int uv_recip;
    // camera (camera.cpp:10-16 evaluated once on the host)
    float cam_pos[3], cam_fwd[3], cam_right[3], cam_up[3];
    float half_w, half_h;
    int dof_on;               // dofEnabled && aperture > 1e-6f (tile_renderer.cpp:99)
    float aperture, focus_dist;
    // scene
    int n_boxes;
    uint32_t posed_mask, usable_mask;   // over boxes 0..31: posed / has triangles
    uint32_t opaque_mask;               // over boxes 0..31: kBoxOpaque
    uint32_t rotated_mask;              // over boxes 0..31: kBoxRotated
    uint32_t opaque_posed_mask;         // over boxes 0..31: posed, and no texel of any face has alpha == 0
    uint32_t root_mask;                 // over boxes 0..31: boxes no other box encloses; 0 = no box encloses another
    int any_rotated;                    // some box (of any index) is posed: needs the MCSKIN_POSED=1 build of the kernels
    float light_pos[3], light_color[4], light_radius;
    float background[4];
    int rng_mode;             // McConfig::rng_mode: 0 std::mt19937, 1 counter-based streams (the MCSKIN_COUNTER_RNG build)
    // integrator
    int use_config;           // traceRay's config pointer non-null (always 1 for render())
    int soft_on;              // softShadows && shadowSamples > 1 (raytracer.cpp:109)
    int shadow_samples;
    int ao_on, ao_samples;
    float ao_radius, ao_intensity;
    int gradient_bg;
    float gradient_scale;
    float bg_center[4], bg_edge[4];
    float kd, ks, ambient, shininess;
    // conservative bounds of everything hittable (world space, slightly inflated): rays
    // that miss it cannot hit any box
    float cull_lo[3], cull_hi[3];
    int cull_valid;
    // the same bounds projected through the pinhole camera, in pixels (inclusive, with a
    // 2-pixel margin): pinhole rays through pixels outside this rectangle miss everything.
    // rect_valid == 0 (a corner behind the camera, degenerate basis, DOF): test every pixel.
    int rect_valid;
    int rect_x0, rect_y0, rect_x1, rect_y1;
    // the same per box (SceneBlobLayout::rect): primary rays only test the boxes whose rectangle holds their pixel
    int box_rects_valid;
};

}  // namespace mcskin
