/*
 * mcskin_oracle.c — TEST INFRASTRUCTURE ONLY (the parity oracle).
 *
 * A plain-C restatement of the reference's render hot path
 *   TileRenderer::render -> renderTile -> Camera::generateRay / generateDOFRay ->
 *   RayTracer::traceRay -> intersectScene/intersectMesh/intersectAABB ->
 *   computeSoftShadow/isInShadow/shade -> computeAO -> backgroundColor
 * written against the flat PODs of include/mcskin_cuda.h.  It deliberately keeps
 * the reference's evaluation order operation by operation (no FMA, no
 * re-association, per-ray recomputation where that changes rounding) so that its
 * float image is BIT-IDENTICAL to the unmodified reference compiled with its
 * Release flags; tests/test_oracle_vs_reference.py pins that against
 * oracle/_ref/libmcskin_ref.so and the committed fixtures in tests/golden/.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load
 * this.  The product (minecraftskin_raytracer_b200/) never does.
 *
 * Third-party arithmetic that is NOT under /root/reference and is restated or
 * relied on here:
 *  - libstdc++ 13.3 std::mt19937 (ISO 26.5.3.2) — restated in mt_seed/mt_next;
 *  - libstdc++ 13.3 std::uniform_real_distribution<float>(0,1) ==
 *    generate_canonical<float,24>: one engine call, float(u32)/2^32, a result
 *    >= 1 becomes nextafterf(1,0) (bits/random.tcc:3349-3381) — canonical_float();
 *  - glibc 2.39 libm cosf/sinf/tanf/powf/sqrtf — called, not restated.
 *
 * Build: gcc -O2 -std=c99 -fno-fast-math -ffp-contract=off (oracle/Makefile).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#include "mcskin_cuda.h"
#include "mcskin_oracle.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#define PI_F ((float)M_PI)

/* ------------------------------------------------------------------ math/vec3.h */
typedef struct { float x, y, z; } V3;
typedef struct { float r, g, b, a; } Col;

static V3 v3(float x, float y, float z) { V3 v = {x, y, z}; return v; }
static V3 vadd(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
static V3 vsub(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
static V3 vmul(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
/* vec3.h:22 — division is a multiply by the rounded reciprocal */
static V3 vdiv(V3 a, float s) { float inv = 1.0f / s; return v3(a.x * inv, a.y * inv, a.z * inv); }
static float vdot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static V3 vcross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static float vlen(V3 a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }
/* vec3.h:46-50 */
static V3 vnorm(V3 a) { float l = vlen(a); if (l < 1e-8f) return v3(0, 0, 0); return vdiv(a, l); }

/* std::clamp(v, lo, hi) */
static float clampf(float v, float lo, float hi) { return (v < lo) ? lo : (hi < v) ? hi : v; }
static Col col(float r, float g, float b, float a) { Col c = {r, g, b, a}; return c; }
static Col cclamp(Col c) { return col(clampf(c.r, 0, 1), clampf(c.g, 0, 1), clampf(c.b, 0, 1), clampf(c.a, 0, 1)); }
static Col cscale(Col c, float s) { return col(c.r * s, c.g * s, c.b * s, c.a * s); }
static Col cadd(Col a, Col b) { return col(a.r + b.r, a.g + b.g, a.b + b.b, a.a + b.a); }
static Col cmul(Col a, Col b) { return col(a.r * b.r, a.g * b.g, a.b * b.b, a.a * b.a); }

/* ------------------------------------------------- std::mt19937 (ISO 26.5.3.2) */
/* counter != 0: the stream is McConfig::rng_mode 1's — word k = mc_rng_counter_word(seed, k) (mcskin_cuda.h); this is
 * the "identically patched oracle" the product's counter-based mode is compared with, not the reference's algorithm. */
typedef struct { uint32_t s[624]; int idx; int counter; uint32_t seed, k; } Mt;

static void mt_seed_mode(Mt* m, uint32_t seed, int counter) {
    m->counter = counter;
    m->seed = seed;
    m->k = 0;
    if (counter) return;
    m->s[0] = seed;
    for (int i = 1; i < 624; ++i) m->s[i] = 1812433253u * (m->s[i - 1] ^ (m->s[i - 1] >> 30)) + (uint32_t)i;
    m->idx = 624;
}
static void mt_seed(Mt* m, uint32_t seed) { mt_seed_mode(m, seed, 0); }

static uint32_t mt_next(Mt* m) {
    if (m->counter) return mc_rng_counter_word(m->seed, m->k++);
    if (m->idx >= 624) {
        for (int i = 0; i < 624; ++i) {
            uint32_t y = (m->s[i] & 0x80000000u) | (m->s[(i + 1) % 624] & 0x7fffffffu);
            m->s[i] = m->s[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        m->idx = 0;
    }
    uint32_t y = m->s[m->idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

/* libstdc++ generate_canonical<float,24,mt19937> as used by uniform_real_distribution<float>(0,1) */
static float canonical_float(Mt* m) {
    float r = (float)mt_next(m) / 4294967296.0f;
    if (r >= 1.0f) r = nextafterf(1.0f, 0.0f);
    return r;
}

/* ------------------------------------------------------------------ counters */
typedef struct Ctx {
    const McScene* sc;
    const McConfig* cfg; /* may be NULL (traceRay without config) */
    McOracleCounters cnt;
    McOracleCounters scratch; /* receives the inner-work counts of the redundant re-test */
    int in_retest;
} Ctx;
/* inner-work counters describe UNIQUE rays only: the re-test of tile_renderer.cpp:111
 * repeats the primary ray exactly, so its box tests are counted into `scratch` */
#define CNT(cx) ((cx)->in_retest ? &(cx)->scratch : &(cx)->cnt)

/* ------------------------------------------------ intersection.cpp:12-42 rotatePoint */
static V3 rotate_point(V3 point, V3 pivot, float rotXDeg, float rotZDeg) {
    V3 p = vsub(point, pivot);
    if (fabsf(rotXDeg) > 0.01f) {
        float rad = rotXDeg * PI_F / 180.0f;
        float c = cosf(rad), s = sinf(rad);
        float ny = p.y * c - p.z * s;
        float nz = p.y * s + p.z * c;
        p.y = ny; p.z = nz;
    }
    if (fabsf(rotZDeg) > 0.01f) {
        float rad = rotZDeg * PI_F / 180.0f;
        float c = cosf(rad), s = sinf(rad);
        float nx = p.x * c - p.y * s;
        float ny = p.x * s + p.y * c;
        p.x = nx; p.y = ny;
    }
    return vadd(p, pivot);
}
static V3 rotate_dir(V3 d, float rx, float rz) { return rotate_point(d, v3(0, 0, 0), rx, rz); }

/* ------------------------------------------------------- texture_region.h:19-26 */
static Col sample_face(const McScene* sc, const McFaceTex* ft, float u, float v) {
    if (ft->texel_offset < 0) return col(1, 0, 1, 1);              /* intersection.cpp:303-306 */
    if (ft->width <= 0 || ft->height <= 0) return col(0, 0, 0, 1); /* texture_region.h:20-22 */
    int x = (int)(u * ft->width);
    int y = (int)(v * ft->height);
    x = x < 0 ? 0 : (x > ft->width - 1 ? ft->width - 1 : x);
    y = y < 0 ? 0 : (y > ft->height - 1 ? ft->height - 1 : y);
    const float* p = sc->texels_rgba + 4 * ((size_t)ft->texel_offset + (size_t)y * ft->width + x);
    return col(p[0], p[1], p[2], p[3]);
}

/* intersection.cpp:86-122: (axis, negSide) -> face index and outward normal */
static int face_of(int axis, int neg, V3* n) {
    if (axis == 2) { if (neg) { *n = v3(0, 0, -1); return 0; } *n = v3(0, 0, 1); return 1; }
    if (axis == 0) { if (!neg) { *n = v3(1, 0, 0); return 2; } *n = v3(-1, 0, 0); return 3; }
    if (!neg) { *n = v3(0, 1, 0); return 4; }
    *n = v3(0, -1, 0); return 5;
}

/* intersection.cpp:136-196 computeFaceUV */
static void face_uv(V3 p, const float lo[3], const float hi[3], int axis, int neg, float* u, float* v) {
    float sx = hi[0] - lo[0], sy = hi[1] - lo[1], sz = hi[2] - lo[2];
    sx = (sx > 1e-8f) ? sx : 1.0f;
    sy = (sy > 1e-8f) ? sy : 1.0f;
    sz = (sz > 1e-8f) ? sz : 1.0f;
    if (axis == 2) {
        float lx = (p.x - lo[0]) / sx, ly = (p.y - lo[1]) / sy;
        *u = neg ? 1.0f - lx : lx;
        *v = 1.0f - ly;
    } else if (axis == 0) {
        float lz = (p.z - lo[2]) / sz, ly = (p.y - lo[1]) / sy;
        *u = !neg ? 1.0f - lz : lz;
        *v = 1.0f - ly;
    } else {
        float lx = (p.x - lo[0]) / sx, lz = (p.z - lo[2]) / sz;
        *u = lx;
        *v = !neg ? lz : 1.0f - lz;
    }
    *u = clampf(*u, 0.0f, 1.0f);
    *v = clampf(*v, 0.0f, 1.0f);
}

/* exit face = axis with the smallest far distance, first axis wins ties (intersection.cpp:265-285,326-337) */
static void exit_face(const float o[3], const float d[3], const float lo[3], const float hi[3], int* axis, int* neg) {
    float best = FLT_MAX;
    *axis = 0; *neg = 0;
    for (int i = 0; i < 3; ++i) {
        if (fabsf(d[i]) < 1e-8f) continue;
        float inv = 1.0f / d[i];
        float t0 = (lo[i] - o[i]) * inv, t1 = (hi[i] - o[i]) * inv;
        int en = 0;
        if (t0 > t1) { float t = t0; t0 = t1; t1 = t; en = 1; }
        if (t1 < best) { best = t1; *axis = i; *neg = en; }
    }
}

/* intersection.cpp:200-371 intersectAABB, on precomputed bounds (computeAABB's
 * min/max over the triangle list is what McBox::bounds_* hold). */
static McHit intersect_box(Ctx* cx, V3 ro, V3 rd, const McBox* box, int boxIndex) {
    McHit res;
    memset(&res, 0, sizeof(res));
    res.box = -1; res.face = -1;
    if (box->n_triangles <= 0) return res;
    const float* lo = box->bounds_min;
    const float* hi = box->bounds_max;
    const float o[3] = {ro.x, ro.y, ro.z}, d[3] = {rd.x, rd.y, rd.z};

    float tmin = -FLT_MAX, tmax = FLT_MAX;
    int axis = 0, neg = 0;
    for (int i = 0; i < 3; ++i) {
        if (fabsf(d[i]) < 1e-8f) {
            if (o[i] < lo[i] || o[i] > hi[i]) return res;
        } else {
            float inv = 1.0f / d[i];
            float t0 = (lo[i] - o[i]) * inv, t1 = (hi[i] - o[i]) * inv;
            int en = 1;
            if (t0 > t1) { float t = t0; t0 = t1; t1 = t; en = 0; }
            if (t0 > tmin) { tmin = t0; axis = i; neg = en; }
            tmax = (t1 < tmax) ? t1 : tmax; /* std::min(tmax, t1) */
            if (tmin > tmax || tmax < 0.0f) return res;
        }
    }
    float tHit = tmin;
    if (tHit < 0.0f) {
        tHit = tmax;
        if (tHit < 0.0f) return res;
        exit_face(o, d, lo, hi, &axis, &neg);
    }
    CNT(cx)->n_slab_pass++;
    V3 p = vadd(ro, vmul(rd, tHit)); /* ray.h:14 */
    V3 n;
    int f = face_of(axis, neg, &n);
    float u, v;
    face_uv(p, lo, hi, axis, neg, &u, &v);
    Col tc = sample_face(cx->sc, &box->face[f], u, v);

    if (tc.a == 0.0f) { /* intersection.cpp:311-361 */
        if (!box->is_outer_layer) return res;
        if (tmax > tHit) {
            int ea, en;
            exit_face(o, d, lo, hi, &ea, &en);
            CNT(cx)->n_backface_eval++;
            V3 bp = vadd(ro, vmul(rd, tmax));
            V3 bn;
            int bf = face_of(ea, en, &bn);
            float bu, bv;
            face_uv(bp, lo, hi, ea, en, &bu, &bv);
            Col bc = sample_face(cx->sc, &box->face[bf], bu, bv);
            if (bc.a > 0.0f) {
                res.hit = 1; res.t = tmax;
                res.point[0] = bp.x; res.point[1] = bp.y; res.point[2] = bp.z;
                V3 fn = vmul(bn, -1.0f);
                res.normal[0] = fn.x; res.normal[1] = fn.y; res.normal[2] = fn.z;
                res.tex_color[0] = bc.r; res.tex_color[1] = bc.g; res.tex_color[2] = bc.b; res.tex_color[3] = bc.a;
                res.is_outer_layer = 1;
                res.box = boxIndex; res.face = bf;
                return res;
            }
        }
        return res;
    }
    res.hit = 1; res.t = tHit;
    res.point[0] = p.x; res.point[1] = p.y; res.point[2] = p.z;
    res.normal[0] = n.x; res.normal[1] = n.y; res.normal[2] = n.z;
    res.tex_color[0] = tc.r; res.tex_color[1] = tc.g; res.tex_color[2] = tc.b; res.tex_color[3] = tc.a;
    res.is_outer_layer = box->is_outer_layer ? 1 : 0;
    res.box = boxIndex; res.face = f;
    return res;
}

/* intersection.cpp:373-406 intersectMesh */
static McHit intersect_mesh(Ctx* cx, V3 ro, V3 rd, int b) {
    const McBox* box = &cx->sc->boxes[b];
    if (!box->has_rotation) {
        CNT(cx)->n_box_tests_plain++;
        return intersect_box(cx, ro, rd, box, b);
    }
    CNT(cx)->n_box_tests_rotated++;
    V3 pivot = v3(box->pivot[0], box->pivot[1], box->pivot[2]);
    V3 lo = rotate_point(ro, pivot, 0, -box->rot_z_deg);
    lo = rotate_point(lo, pivot, -box->rot_x_deg, 0);
    V3 ld = rotate_dir(rd, 0, -box->rot_z_deg);
    ld = rotate_dir(ld, -box->rot_x_deg, 0);
    McHit h = intersect_box(cx, lo, vnorm(ld), box, b);
    if (h.hit) {
        CNT(cx)->n_rotated_hits++;
        V3 p = rotate_point(v3(h.point[0], h.point[1], h.point[2]), pivot, box->rot_x_deg, box->rot_z_deg);
        V3 n = vnorm(rotate_dir(v3(h.normal[0], h.normal[1], h.normal[2]), box->rot_x_deg, box->rot_z_deg));
        h.point[0] = p.x; h.point[1] = p.y; h.point[2] = p.z;
        h.normal[0] = n.x; h.normal[1] = n.y; h.normal[2] = n.z;
        h.t = vdot(vsub(p, ro), rd);
    }
    return h;
}

/* intersection.cpp:408-421 intersectScene */
static McHit intersect_scene(Ctx* cx, V3 ro, V3 rd) {
    McHit best;
    memset(&best, 0, sizeof(best));
    best.t = FLT_MAX; best.box = -1; best.face = -1;
    cx->cnt.n_intersect_scene++;
    for (int b = 0; b < cx->sc->n_boxes; ++b) {
        McHit h = intersect_mesh(cx, ro, rd, b);
        if (h.hit && h.t < best.t) best = h;
    }
    return best;
}

/* ---------------------------------------------------------------- shading.cpp */
#define SHADOW_EPSILON 1e-3f

/* shading.cpp:14-26 */
static int in_shadow(Ctx* cx, V3 point, V3 normal, V3 lightPos) {
    V3 origin = vadd(point, vmul(normal, SHADOW_EPSILON));
    V3 toLight = vsub(lightPos, origin);
    float dist = vlen(toLight);
    if (dist < 1e-6f) return 0;
    V3 dir = vdiv(toLight, dist);
    cx->cnt.n_shadow_rays++;
    McHit h = intersect_scene(cx, origin, dir);
    return h.hit && h.t < dist;
}

/* shading.cpp:28-60 */
static float soft_shadow(Ctx* cx, V3 point, V3 normal, int samples, uint32_t seed) {
    const McScene* sc = cx->sc;
    V3 lp = v3(sc->light_pos[0], sc->light_pos[1], sc->light_pos[2]);
    if (samples <= 1 || sc->light_radius < 1e-4f) return in_shadow(cx, point, normal, lp) ? 0.0f : 1.0f;
    V3 toPoint = vnorm(vsub(point, lp));
    V3 tangent;
    if (fabsf(toPoint.x) < 0.9f) tangent = vnorm(vcross(v3(1, 0, 0), toPoint));
    else tangent = vnorm(vcross(v3(0, 1, 0), toPoint));
    V3 bitangent = vcross(toPoint, tangent);
    Mt rng;
    mt_seed_mode(&rng, seed, cx->cfg != NULL && cx->cfg->rng_mode == MC_RNG_COUNTER);
    int lit = 0;
    for (int i = 0; i < samples; ++i) {
        float angle = 2.0f * PI_F * canonical_float(&rng);
        float r = sc->light_radius * sqrtf(canonical_float(&rng));
        V3 offset = vadd(vmul(tangent, r * cosf(angle)), vmul(bitangent, r * sinf(angle)));
        V3 samplePos = vadd(lp, offset);
        if (!in_shadow(cx, point, normal, samplePos)) ++lit;
    }
    return (float)lit / (float)samples;
}

/* shading.cpp:62-96 */
static Col shade_hit(Ctx* cx, const McHit* hit, V3 viewDir, float kd, float ks, float ambientK, float shininess,
                     float shadowFactor) {
    const McScene* sc = cx->sc;
    cx->cnt.n_shade++;
    Col tex = col(hit->tex_color[0], hit->tex_color[1], hit->tex_color[2], hit->tex_color[3]);
    float alpha = tex.a;
    Col ambient = cscale(tex, ambientK);
    V3 lp = v3(sc->light_pos[0], sc->light_pos[1], sc->light_pos[2]);
    V3 P = v3(hit->point[0], hit->point[1], hit->point[2]);
    V3 L = vnorm(vsub(lp, P));
    V3 N = vnorm(v3(hit->normal[0], hit->normal[1], hit->normal[2]));
    V3 V = vnorm(viewDir);
    float vis = shadowFactor;
    if (vis < 0.0f) {
        cx->cnt.n_hard_shadow++;
        vis = in_shadow(cx, P, N, lp) ? 0.0f : 1.0f;
    }
    Col lc = col(sc->light_color[0], sc->light_color[1], sc->light_color[2], sc->light_color[3]);
    float ndl = vdot(N, L); ndl = (ndl > 0.0f) ? ndl : 0.0f; /* std::max(0.0f, x) */
    Col diffuse = cscale(cmul(tex, lc), kd * ndl * vis);
    V3 H = vnorm(vadd(L, V));
    float ndh = vdot(N, H); ndh = (ndh > 0.0f) ? ndh : 0.0f;
    float spec = powf(ndh, shininess);
    Col specular = cscale(lc, ks * spec * vis);
    Col result = cadd(cadd(ambient, diffuse), specular);
    result.a = alpha;
    return cclamp(result);
}

/* --------------------------------------------------------------- raytracer.cpp */
#define SKIN_REFLECTIVITY 0.1f
#define REFLECT_EPSILON 1e-3f

/* raytracer.cpp:16-34 */
static Col background(const McScene* sc, float u, float v, const McConfig* cfg) {
    if (cfg && cfg->gradient_bg) {
        float cxx = u - 0.5f, cyy = v - 0.5f;
        float dist = sqrtf(cxx * cxx + cyy * cyy) * 2.0f * cfg->gradient_scale;
        dist = clampf(dist, 0.0f, 1.0f);
        float t = dist * dist;
        Col c;
        c.r = cfg->bg_center[0] * (1.0f - t) + cfg->bg_edge[0] * t;
        c.g = cfg->bg_center[1] * (1.0f - t) + cfg->bg_edge[1] * t;
        c.b = cfg->bg_center[2] * (1.0f - t) + cfg->bg_edge[2] * t;
        c.a = 1.0f;
        return c;
    }
    return col(sc->background[0], sc->background[1], sc->background[2], sc->background[3]);
}

/* raytracer.cpp:38-78 */
static float ambient_occlusion(Ctx* cx, V3 point, V3 normal, int samples, float radius, uint32_t seed) {
    V3 N = vnorm(normal);
    V3 T;
    if (fabsf(N.x) < 0.9f) T = vnorm(vcross(v3(1, 0, 0), N));
    else T = vnorm(vcross(v3(0, 1, 0), N));
    V3 B = vcross(N, T);
    Mt rng;
    mt_seed_mode(&rng, seed, cx->cfg != NULL && cx->cfg->rng_mode == MC_RNG_COUNTER);
    int occluded = 0;
    for (int i = 0; i < samples; ++i) {
        float r1 = canonical_float(&rng);
        float r2 = canonical_float(&rng);
        float sinT = sqrtf(1.0f - r1);
        float cosT = sqrtf(r1);
        float phi = 2.0f * PI_F * r2;
        V3 ld = v3(sinT * cosf(phi), cosT, sinT * sinf(phi));
        V3 wd = vadd(vadd(vmul(T, ld.x), vmul(N, ld.y)), vmul(B, ld.z));
        wd = vnorm(wd);
        cx->cnt.n_ao_rays++;
        McHit h = intersect_scene(cx, vadd(point, vmul(N, 1e-3f)), wd);
        if (h.hit && h.t < radius) ++occluded;
    }
    return 1.0f - (float)occluded / (float)samples;
}

/* x86-64 gcc lowers static_cast<unsigned>(float) to a 64-bit cvttss2si and keeps the
 * low 32 bits (negative values wrap, out-of-range gives 0).  gcc compiles this C
 * cast the same way; written through int64 to make the intent explicit. */
static uint32_t seed_cast(float f) {
    if (!(f > -9.2233720368547758e18f && f < 9.2233720368547758e18f)) return 0u;
    return (uint32_t)(int64_t)f;
}

/* raytracer.cpp:82-148 */
static Col trace_ray(Ctx* cx, V3 ro, V3 rd, int depth, int maxBounces, float kd, float ks, float amb, float shin) {
    const McScene* sc = cx->sc;
    const McConfig* cfg = cx->cfg;
    Col flat = col(sc->background[0], sc->background[1], sc->background[2], sc->background[3]);
    if (depth > maxBounces) return cfg ? background(sc, 0.5f, 0.5f, cfg) : flat;
    McHit hit = intersect_scene(cx, ro, rd);
    if (!hit.hit) {
        if (depth == 0 && cfg) return background(sc, 0.5f, 0.5f, cfg);
        return flat;
    }
    V3 P = v3(hit.point[0], hit.point[1], hit.point[2]);
    V3 Nraw = v3(hit.normal[0], hit.normal[1], hit.normal[2]);
    V3 viewDir = vnorm(vsub(ro, P));
    float shadowFactor = -1.0f;
    if (cfg && cfg->soft_shadows && cfg->shadow_samples > 1) {
        uint32_t seed = seed_cast(P.x * 12345.0f + P.y * 67890.0f + P.z * 11111.0f + (float)depth * 99999.0f);
        cx->cnt.n_soft_shadow++;
        shadowFactor = soft_shadow(cx, P, Nraw, cfg->shadow_samples, seed);
    }
    Col shaded = shade_hit(cx, &hit, viewDir, kd, ks, amb, shin, shadowFactor);
    float alpha = shaded.a;
    if (cfg && cfg->ao_enabled && depth == 0) {
        uint32_t seed = seed_cast(P.x * 73856093.0f + P.y * 19349663.0f + P.z * 83492791.0f);
        float ao = ambient_occlusion(cx, P, Nraw, cfg->ao_samples, cfg->ao_radius, seed);
        float f = 1.0f - cfg->ao_intensity * (1.0f - ao);
        shaded.r *= f; shaded.g *= f; shaded.b *= f;
    }
    if (depth < maxBounces) {
        V3 N = vnorm(Nraw);
        V3 D = vnorm(rd);
        V3 R = vsub(D, vmul(N, 2.0f * vdot(D, N)));
        R = vnorm(R);
        V3 origin = vadd(P, vmul(N, REFLECT_EPSILON));
        cx->cnt.n_reflect_rays++;
        Col refl = trace_ray(cx, origin, R, depth + 1, maxBounces, kd, ks, amb, shin);
        shaded = cadd(cscale(shaded, 1.0f - SKIN_REFLECTIVITY), cscale(refl, SKIN_REFLECTIVITY));
    }
    shaded.a = alpha;
    return cclamp(shaded);
}

/* ------------------------------------------------------------------ camera.cpp */
/* camera.cpp:8-26 */
static void camera_ray(const McScene* sc, float u, float v, float aspect, V3* ro, V3* rd) {
    V3 pos = v3(sc->cam_pos[0], sc->cam_pos[1], sc->cam_pos[2]);
    V3 tgt = v3(sc->cam_target[0], sc->cam_target[1], sc->cam_target[2]);
    V3 up = v3(sc->cam_up[0], sc->cam_up[1], sc->cam_up[2]);
    V3 fwd = vnorm(vsub(tgt, pos));
    V3 right = vnorm(vcross(fwd, up));
    V3 trueUp = vcross(right, fwd);
    float halfH = tanf(sc->cam_fov_deg * 0.5f * PI_F / 180.0f);
    float halfW = halfH * aspect;
    float su = (2.0f * u - 1.0f) * halfW;
    float sv = (2.0f * (1.0f - v) - 1.0f) * halfH;
    *rd = vnorm(vadd(vadd(fwd, vmul(right, su)), vmul(trueUp, sv)));
    *ro = pos;
}

/* tile_renderer.cpp:42-69 generateDOFRay */
static void dof_ray(const McScene* sc, float u, float v, float aspect, float aperture, float focusDist, Mt* rng,
                    V3* ro, V3* rd) {
    V3 po, pd;
    camera_ray(sc, u, v, aspect, &po, &pd);
    if (aperture < 1e-6f) { *ro = po; *rd = pd; return; }
    V3 pos = v3(sc->cam_pos[0], sc->cam_pos[1], sc->cam_pos[2]);
    V3 tgt = v3(sc->cam_target[0], sc->cam_target[1], sc->cam_target[2]);
    V3 up = v3(sc->cam_up[0], sc->cam_up[1], sc->cam_up[2]);
    V3 fwd = vnorm(vsub(tgt, pos));
    V3 right = vnorm(vcross(fwd, up));
    V3 camUp = vcross(right, fwd);
    V3 focus = vadd(po, vmul(pd, focusDist));
    float angle = 2.0f * PI_F * canonical_float(rng);
    float radius = aperture * sqrtf(canonical_float(rng));
    float lx = radius * cosf(angle);
    float ly = radius * sinf(angle);
    V3 lens = vadd(vmul(right, lx), vmul(camUp, ly));
    *ro = vadd(pos, lens);
    *rd = vnorm(vsub(focus, *ro));
}

/* ------------------------------------------------------------ tile_renderer.cpp */
int32_t mcorc_generate_tiles(int32_t w, int32_t h, int32_t ts, McTile* out, int32_t cap) {
    if (w <= 0 || h <= 0 || ts <= 0) return 0; /* tile_renderer.cpp:19-21 */
    int cols = (w + ts - 1) / ts, rows = (h + ts - 1) / ts, n = 0;
    for (int ty = 0; ty < rows; ++ty)
        for (int tx = 0; tx < cols; ++tx, ++n) {
            if (out && n < cap) {
                out[n].x = tx * ts; out[n].y = ty * ts;
                out[n].width = (ts < w - tx * ts) ? ts : w - tx * ts;
                out[n].height = (ts < h - ty * ts) ? ts : h - ty * ts;
            }
        }
    return n;
}

/* tile_renderer.cpp:71-127 renderTile */
static void render_tile(Ctx* cx, const McTile* tile, float* image) {
    const McScene* sc = cx->sc;
    const McConfig* cfg = cx->cfg;
    float aspect = (float)cfg->width / (float)cfg->height;
    int spp = cfg->samples_per_pixel > 1 ? cfg->samples_per_pixel : 1;
    Mt rng;
    mt_seed_mode(&rng, (uint32_t)(tile->y * cfg->width + tile->x), cfg->rng_mode == MC_RNG_COUNTER);
    float focusDist = cfg->focus_distance;
    if (focusDist <= 0.0f) {
        V3 pos = v3(sc->cam_pos[0], sc->cam_pos[1], sc->cam_pos[2]);
        V3 tgt = v3(sc->cam_target[0], sc->cam_target[1], sc->cam_target[2]);
        focusDist = vlen(vsub(tgt, pos));
    }
    for (int py = tile->y; py < tile->y + tile->height; ++py)
        for (int px = tile->x; px < tile->x + tile->width; ++px) {
            Col acc = col(0, 0, 0, 0);
            for (int s = 0; s < spp; ++s) {
                float jx = (spp == 1) ? 0.5f : canonical_float(&rng);
                float jy = (spp == 1) ? 0.5f : canonical_float(&rng);
                float u = ((float)px + jx) / (float)cfg->width;
                float v = ((float)py + jy) / (float)cfg->height;
                V3 ro, rd;
                if (cfg->dof_enabled && cfg->aperture > 1e-6f) dof_ray(sc, u, v, aspect, cfg->aperture, focusDist, &rng, &ro, &rd);
                else camera_ray(sc, u, v, aspect, &ro, &rd);
                cx->cnt.n_primary_rays++;
                /* tile_renderer.cpp:106 always passes ShadingParams{}; McConfig carries the same defaults */
                Col c = trace_ray(cx, ro, rd, 0, cfg->max_bounces, cfg->kd, cfg->ks, cfg->ambient, cfg->shininess);
                cx->cnt.n_retests++;
                cx->in_retest = 1;
                McHit again = intersect_scene(cx, ro, rd); /* tile_renderer.cpp:111 */
                cx->in_retest = 0;
                if (!again.hit) { c = background(sc, u, v, cfg); cx->cnt.n_background_primary++; }
                acc.r += c.r; acc.g += c.g; acc.b += c.b; acc.a += c.a;
            }
            float inv = 1.0f / (float)spp;
            float* o = image + 4 * ((size_t)py * cfg->width + px);
            o[0] = acc.r * inv; o[1] = acc.g * inv; o[2] = acc.b * inv; o[3] = acc.a * inv;
        }
}

static void add_counters(McOracleCounters* a, const McOracleCounters* b) {
    int64_t* x = (int64_t*)a;
    const int64_t* y = (const int64_t*)b;
    for (size_t i = 0; i < sizeof(McOracleCounters) / sizeof(int64_t); ++i) x[i] += y[i];
}

typedef struct Job {
    const McScene* sc; const McConfig* cfg; const McTile* tiles; int nTiles; float* image;
    volatile int next; pthread_mutex_t mu; McOracleCounters total;
} Job;

static void* worker(void* arg) {
    Job* job = (Job*)arg;
    Ctx cx;
    memset(&cx, 0, sizeof(cx));
    cx.sc = job->sc; cx.cfg = job->cfg;
    for (;;) {
        int idx = __sync_fetch_and_add(&job->next, 1); /* tile_renderer.cpp:155 */
        if (idx >= job->nTiles) break;
        render_tile(&cx, &job->tiles[idx], job->image);
    }
    pthread_mutex_lock(&job->mu);
    add_counters(&job->total, &cx.cnt);
    pthread_mutex_unlock(&job->mu);
    return NULL;
}

/* tile_renderer.cpp:129-189 render.  threads <= 0: cfg->thread_count, then all cores. */
int32_t mcorc_render(const McScene* sc, const McConfig* cfg, int32_t threads, float* out, McOracleCounters* counters) {
    if (!sc || !cfg || !out) return -1;
    int nTiles = mcorc_generate_tiles(cfg->width, cfg->height, cfg->tile_size, NULL, 0);
    if (counters) memset(counters, 0, sizeof(*counters));
    /* Image(w,h) default-constructs every pixel to (0,0,0,1) (image.h:15, color.h:8) */
    if (cfg->width > 0 && cfg->height > 0)
        for (size_t i = 0; i < (size_t)cfg->width * cfg->height; ++i) { out[4 * i] = out[4 * i + 1] = out[4 * i + 2] = 0.0f; out[4 * i + 3] = 1.0f; }
    if (nTiles == 0) return 0;
    McTile* tiles = (McTile*)malloc(sizeof(McTile) * (size_t)nTiles);
    mcorc_generate_tiles(cfg->width, cfg->height, cfg->tile_size, tiles, nTiles);
    if (threads <= 0) threads = cfg->thread_count;
    if (threads <= 0) threads = mcorc_hardware_threads();
    if (threads > nTiles) threads = nTiles;
    if (threads > 256) threads = 256;
    Job job;
    memset(&job, 0, sizeof(job));
    job.sc = sc; job.cfg = cfg; job.tiles = tiles; job.nTiles = nTiles; job.image = out;
    pthread_mutex_init(&job.mu, NULL);
    pthread_t th[256];
    for (int i = 0; i < threads; ++i) pthread_create(&th[i], NULL, worker, &job);
    for (int i = 0; i < threads; ++i) pthread_join(th[i], NULL);
    pthread_mutex_destroy(&job.mu);
    if (counters) *counters = job.total;
    free(tiles);
    return 0;
}

int32_t mcorc_render_tile(const McScene* sc, const McConfig* cfg, const McTile* tile, float* image) {
    Ctx cx;
    memset(&cx, 0, sizeof(cx));
    cx.sc = sc; cx.cfg = cfg;
    render_tile(&cx, tile, image);
    return 0;
}

#include <unistd.h>
int32_t mcorc_hardware_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int32_t)n : 1;
}

/* ------------------------------------------- single-call views (reference's free functions) */
int32_t mcorc_intersect(const McScene* sc, int32_t box, const McRay* rays, int32_t n, McHit* out) {
    Ctx cx;
    memset(&cx, 0, sizeof(cx));
    cx.sc = sc;
    for (int i = 0; i < n; ++i) {
        V3 ro = v3(rays[i].origin[0], rays[i].origin[1], rays[i].origin[2]);
        V3 rd = v3(rays[i].dir[0], rays[i].dir[1], rays[i].dir[2]);
        out[i] = box >= 0 ? intersect_mesh(&cx, ro, rd, box) : intersect_scene(&cx, ro, rd);
        if (!out[i].hit) { memset(&out[i], 0, sizeof(McHit)); out[i].box = -1; out[i].face = -1; }
    }
    return 0;
}

int32_t mcorc_trace(const McScene* sc, const McConfig* cfg, int32_t useConfig, int32_t depth, const McRay* rays,
                    int32_t n, float* out) {
    Ctx cx;
    memset(&cx, 0, sizeof(cx));
    cx.sc = sc; cx.cfg = useConfig ? cfg : NULL;
    for (int i = 0; i < n; ++i) {
        Col c = trace_ray(&cx, v3(rays[i].origin[0], rays[i].origin[1], rays[i].origin[2]),
                          v3(rays[i].dir[0], rays[i].dir[1], rays[i].dir[2]), depth, cfg->max_bounces, cfg->kd,
                          cfg->ks, cfg->ambient, cfg->shininess);
        out[4 * i] = c.r; out[4 * i + 1] = c.g; out[4 * i + 2] = c.b; out[4 * i + 3] = c.a;
    }
    return 0;
}

int32_t mcorc_shade(const McScene* sc, const McConfig* cfg, const McHit* hits, const float* viewDirs,
                    const float* shadowFactors, int32_t n, float* out) {
    Ctx cx;
    memset(&cx, 0, sizeof(cx));
    cx.sc = sc;
    for (int i = 0; i < n; ++i) {
        Col c = shade_hit(&cx, &hits[i], v3(viewDirs[3 * i], viewDirs[3 * i + 1], viewDirs[3 * i + 2]), cfg->kd,
                          cfg->ks, cfg->ambient, cfg->shininess, shadowFactors ? shadowFactors[i] : -1.0f);
        out[4 * i] = c.r; out[4 * i + 1] = c.g; out[4 * i + 2] = c.b; out[4 * i + 3] = c.a;
    }
    return 0;
}

int32_t mcorc_in_shadow(const McScene* sc, const float* p, const float* nr, const float* l, int32_t n, int32_t* out) {
    Ctx cx;
    memset(&cx, 0, sizeof(cx));
    cx.sc = sc;
    for (int i = 0; i < n; ++i)
        out[i] = in_shadow(&cx, v3(p[3 * i], p[3 * i + 1], p[3 * i + 2]), v3(nr[3 * i], nr[3 * i + 1], nr[3 * i + 2]),
                           v3(l[3 * i], l[3 * i + 1], l[3 * i + 2]));
    return 0;
}

int32_t mcorc_soft_shadow(const McScene* sc, const float* p, const float* nr, const uint32_t* seeds, int32_t samples,
                          int32_t n, float* out) {
    Ctx cx;
    memset(&cx, 0, sizeof(cx));
    cx.sc = sc;
    for (int i = 0; i < n; ++i)
        out[i] = soft_shadow(&cx, v3(p[3 * i], p[3 * i + 1], p[3 * i + 2]),
                             v3(nr[3 * i], nr[3 * i + 1], nr[3 * i + 2]), samples, seeds[i]);
    return 0;
}

int32_t mcorc_ambient_occlusion(const McScene* sc, const float* p, const float* nr, const uint32_t* seeds,
                                int32_t samples, float radius, int32_t n, float* out) {
    Ctx cx;
    memset(&cx, 0, sizeof(cx));
    cx.sc = sc;
    for (int i = 0; i < n; ++i)
        out[i] = ambient_occlusion(&cx, v3(p[3 * i], p[3 * i + 1], p[3 * i + 2]),
                                   v3(nr[3 * i], nr[3 * i + 1], nr[3 * i + 2]), samples, radius, seeds[i]);
    return 0;
}

int32_t mcorc_generate_rays(const McScene* sc, float aspect, const float* uv, int32_t n, McRay* out) {
    for (int i = 0; i < n; ++i) {
        V3 ro, rd;
        camera_ray(sc, uv[2 * i], uv[2 * i + 1], aspect, &ro, &rd);
        out[i].origin[0] = ro.x; out[i].origin[1] = ro.y; out[i].origin[2] = ro.z;
        out[i].dir[0] = rd.x; out[i].dir[1] = rd.y; out[i].dir[2] = rd.z;
    }
    return 0;
}

int32_t mcorc_background(const McScene* sc, const McConfig* cfg, int32_t useConfig, const float* uv, int32_t n,
                         float* out) {
    for (int i = 0; i < n; ++i) {
        Col c = background(sc, uv[2 * i], uv[2 * i + 1], useConfig ? cfg : NULL);
        out[4 * i] = c.r; out[4 * i + 1] = c.g; out[4 * i + 2] = c.b; out[4 * i + 3] = c.a;
    }
    return 0;
}

int32_t mcorc_aov(const McScene* sc, const McConfig* cfg, int32_t* outTriId) {
    Ctx cx;
    memset(&cx, 0, sizeof(cx));
    cx.sc = sc;
    float aspect = (float)cfg->width / (float)cfg->height;
    for (int py = 0; py < cfg->height; ++py)
        for (int px = 0; px < cfg->width; ++px) {
            float u = ((float)px + 0.5f) / (float)cfg->width;
            float v = ((float)py + 0.5f) / (float)cfg->height;
            V3 ro, rd;
            camera_ray(sc, u, v, aspect, &ro, &rd);
            McHit h = intersect_scene(&cx, ro, rd);
            outTriId[(size_t)py * cfg->width + px] = h.hit ? h.box * 12 + h.face * 2 : -1;
        }
    return 0;
}

/* image_writer.cpp:18-22 */
void mcorc_quantize(const float* rgba, int64_t nFloats, uint8_t* out) {
    for (int64_t i = 0; i < nFloats; ++i) out[i] = (uint8_t)(clampf(rgba[i], 0.0f, 1.0f) * 255.0f + 0.5f);
}

/* raw generator views, to pin the libstdc++ restatement */
void mcorc_mt19937(uint32_t seed, int32_t n, uint32_t* outU32, float* outCanonical) {
    Mt a, b;
    mt_seed(&a, seed);
    mt_seed(&b, seed);
    for (int i = 0; i < n; ++i) {
        if (outU32) outU32[i] = mt_next(&a);
        if (outCanonical) outCanonical[i] = canonical_float(&b);
    }
}
uint32_t mcorc_seed_cast(float f) { return seed_cast(f); }
