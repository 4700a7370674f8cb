/* mcskin_oracle.h — TEST INFRASTRUCTURE ONLY: C API of the CPU parity oracle
 * (oracle/mcskin_oracle.c).  Same flat PODs as the product's C ABI so one set of
 * inputs drives the reference shim, the oracle and the CUDA path. */
#ifndef MCSKIN_ORACLE_H
#define MCSKIN_ORACLE_H
#include <stdint.h>
#include "mcskin_cuda.h"
#ifdef __cplusplus
extern "C" {
#endif

/* Exact event counts of one render; the inputs of the algorithmic work model
 * (DESIGN.md "Work model").  All int64, summed over threads.
 * n_intersect_scene counts every invocation (what ld --wrap sees on the reference);
 * the box-level counters below it cover UNIQUE rays only, i.e. they exclude the
 * redundant re-test of each primary ray (tile_renderer.cpp:111). */
typedef struct McOracleCounters {
    int64_t n_intersect_scene;   /* intersectScene invocations (incl. the redundant re-test) */
    int64_t n_primary_rays;      /* pixel-samples */
    int64_t n_retests;           /* tile_renderer.cpp:111 redundant re-intersections */
    int64_t n_shadow_rays;       /* isInShadow rays actually cast */
    int64_t n_ao_rays;
    int64_t n_reflect_rays;
    int64_t n_box_tests_plain;   /* intersectMesh on unposed boxes */
    int64_t n_box_tests_rotated; /* intersectMesh on posed boxes */
    int64_t n_slab_pass;         /* slab tests that reached the face/UV/texel evaluation */
    int64_t n_backface_eval;     /* outer-layer exit-face evaluations */
    int64_t n_rotated_hits;      /* accepted hits on posed boxes (back-transform) */
    int64_t n_background_primary;/* primary samples resolved to the gradient background */
    int64_t n_shade;             /* shade() calls */
    int64_t n_soft_shadow;       /* computeSoftShadow calls from traceRay */
    int64_t n_hard_shadow;       /* hard shadow tests inside shade() */
} McOracleCounters;

int32_t mcorc_generate_tiles(int32_t w, int32_t h, int32_t tile_size, McTile* out, int32_t capacity);
int32_t mcorc_render(const McScene* sc, const McConfig* cfg, int32_t threads, float* out_rgba, McOracleCounters* counters);
int32_t mcorc_render_tile(const McScene* sc, const McConfig* cfg, const McTile* tile, float* image_rgba);
int32_t mcorc_hardware_threads(void);
int32_t mcorc_intersect(const McScene* sc, int32_t box, const McRay* rays, int32_t n, McHit* out);
int32_t mcorc_trace(const McScene* sc, const McConfig* cfg, int32_t use_config, int32_t depth, const McRay* rays, int32_t n, float* out_rgba);
int32_t mcorc_shade(const McScene* sc, const McConfig* cfg, const McHit* hits, const float* view_dirs, const float* shadow_factors, int32_t n, float* out_rgba);
int32_t mcorc_in_shadow(const McScene* sc, const float* points, const float* normals, const float* lights, int32_t n, int32_t* out);
int32_t mcorc_soft_shadow(const McScene* sc, const float* points, const float* normals, const uint32_t* seeds, int32_t samples, int32_t n, float* out);
int32_t mcorc_ambient_occlusion(const McScene* sc, const float* points, const float* normals, const uint32_t* seeds, int32_t samples, float radius, int32_t n, float* out);
int32_t mcorc_generate_rays(const McScene* sc, float aspect, const float* uv, int32_t n, McRay* out);
int32_t mcorc_background(const McScene* sc, const McConfig* cfg, int32_t use_config, const float* uv, int32_t n, float* out_rgba);
int32_t mcorc_aov(const McScene* sc, const McConfig* cfg, int32_t* out_tri_id);
void mcorc_quantize(const float* rgba, int64_t n_floats, uint8_t* out);
void mcorc_mt19937(uint32_t seed, int32_t n, uint32_t* out_u32, float* out_canonical);
uint32_t mcorc_seed_cast(float f);

#ifdef __cplusplus
}
#endif
#endif
