// dev_shade.cuh — camera rays, background, shadows, Blinn-Phong, AO and the bounce
// loop: Camera::generateRay (camera.cpp:8-26), generateDOFRay (tile_renderer.cpp:42-69),
// RayTracer::backgroundColor / computeAO / traceRay (raytracer.cpp:16-148),
// isInShadow / computeSoftShadow / shade (shading.cpp:14-96).
//
// The recursion of traceRay is flattened: the loop walks down the bounce chain
// keeping each level's shaded colour on a small per-thread stack, then folds the
// chain back-to-front with the reference's own mix (c*0.9f + r*0.1f, alpha restored,
// clamp) so every level rounds exactly as the recursive code does.
#pragma once
#include "dev_intersect.cuh"
#include "dev_mt19937.cuh"

namespace mcskin {

constexpr float kShadowEpsilon = 1e-3f;    // shading.cpp:12
constexpr float kReflectEpsilon = 1e-3f;   // raytracer.cpp:12
constexpr float kReflectivity = 0.1f;      // raytracer.cpp:11
constexpr int kMaxStackDepth = 64;         // deeper levels weigh < 1e-64: invisible in float
// 2.0f * static_cast<float>(M_PI), evaluated in float like the reference (shading.cpp:49)
#define MCSKIN_TWO_PI_F (2.0f * 3.14159274101257324219f)

// Camera::generateRay with the look-at basis and tan(fov/2) hoisted to the host.
__device__ __forceinline__ Ray camera_ray(const DevFrame& fr, float u, float v) {
    const float su = (2.0f * u - 1.0f) * fr.half_w;
    const float sv = (2.0f * (1.0f - v) - 1.0f) * fr.half_h;
    const V3 fwd = ld3(fr.cam_fwd), right = ld3(fr.cam_right), up = ld3(fr.cam_up);
    Ray r;
    r.o = ld3(fr.cam_pos);
    r.d = normalize3((fwd + right * su) + up * sv);
    return r;
}

// sinf / cosf of a float angle, bit-identical to glibc 2.39's sinf / cosf / sincosf on an
// x86-64 host with FMA (the variant its ifunc resolver picks on every CPU since Haswell / Zen):
// the angle is widened to double, reduced by a multiple of pi/2 with one fused step, and the
// result is a degree-7 / degree-8 polynomial in double rounded once to float (glibc
// sysdeps/ieee754/flt-32/s_sincosf.h, sincosf_poly, reduce_fast; coefficients are
// __sincosf_table's).  Every a + b*c of that source is a fused multiply-add in the host binary
// (gcc contracts them under -mfma), hence the explicit fma() calls; the plain products stay
// separate (--fmad=false).  Checked on the host against libm for all 1.12e9 floats in [0, 120)
// (tests/test_host_side.py keeps a sampled version of that check).  |angle| >= 120, inf and NaN
// (never produced here: angles are 2*pi*u, u in [0,1)) take CUDA's sincosf.
__device__ __forceinline__ void sincos_ref(float a, float* sn, float* cs) {
    const uint32_t top12 = (__float_as_uint(a) >> 20) & 0x7ffu;
    if (top12 > 0x42eu) {  // |a| >= 120: out of the fast-reduction range
        sincosf(a, sn, cs);
        return;
    }
    if (top12 < 0x398u) {  // |a| < 2^-12
        *sn = a;
        *cs = 1.0f;
        return;
    }
    double x = static_cast<double>(a);
    double xs = x;
    int n = 0;
    if (top12 >= 0x3f4u) {  // |a| >= pi/4: x -= n * pi/2, n = round(x * 2/pi)
        const double r = x * 0x1.45F306DC9C883p+23;
        n = (__double2int_rz(r) + 0x800000) >> 24;
        x = fma(-static_cast<double>(n), 0x1.921FB54442D18p0, x);
        xs = ((n + 1) & 2) ? -x : x;  // sign[n & 3] = {1, -1, -1, 1}
    }
    const bool flipCos = (n & 2) != 0;  // second table: cosine coefficients negated
    const double c0 = flipCos ? -0x1p0 : 0x1p0;
    const double c1 = flipCos ? 0x1.ffffffd0c621cp-2 : -0x1.ffffffd0c621cp-2;
    const double c2 = flipCos ? -0x1.55553e1068f19p-5 : 0x1.55553e1068f19p-5;
    const double c3 = flipCos ? 0x1.6c087e89a359dp-10 : -0x1.6c087e89a359dp-10;
    const double c4 = flipCos ? -0x1.99343027bf8c3p-16 : 0x1.99343027bf8c3p-16;
    const double s1 = -0x1.555545995a603p-3, s2 = 0x1.1107605230bc4p-7, s3 = -0x1.994eb3774cf24p-13;
    const double x2 = x * x;
    const double x3 = xs * x2, x4 = x2 * x2;
    const double sq = fma(x2, s3, s2), cq = fma(x2, c4, c3), cl = fma(x2, c1, c0);
    const double x5 = x3 * x2, x6 = x4 * x2;
    const double sl = fma(x3, s1, xs), cm = fma(x4, c2, cl);
    const float fs = static_cast<float>(fma(x5, sq, sl));
    const float fc = static_cast<float>(fma(x6, cq, cm));
    *sn = (n & 1) ? fc : fs;
    *cs = (n & 1) ? fs : fc;
}

// powf(x, y) bit-identical to glibc 2.39's (x86-64 FMA variant; sysdeps/ieee754/flt-32/e_powf.c):
// log2(x) from a 16-entry table of reciprocals plus a degree-5 polynomial, y*log2(x) in double,
// exp2 from a 32-entry table plus a cubic, one rounding to float.  Tables are __powf_log2_data and
// __exp2f_data; every a*b+c of the source is fused in the host binary, hence the explicit fma().
// Checked on the host against libm for every positive float x <= 2 and nine exponents.  Zero,
// negative, infinite or NaN x, zero / infinite / NaN y and results beyond 2^127.99 take CUDA's powf
// (same special values).
static __device__ const double kPowLog2Tab[16][2] = {
    {0x1.661ec79f8f3bep+0, -0x1.efec65b963019p-2}, {0x1.571ed4aaf883dp+0, -0x1.b0b6832d4fca4p-2},
    {0x1.49539f0f010b0p+0, -0x1.7418b0a1fb77bp-2}, {0x1.3c995b0b80385p+0, -0x1.39de91a6dcf7bp-2},
    {0x1.30d190c8864a5p+0, -0x1.01d9bf3f2b631p-2}, {0x1.25e227b0b8ea0p+0, -0x1.97c1d1b3b7af0p-3},
    {0x1.1bb4a4a1a343fp+0, -0x1.2f9e393af3c9fp-3}, {0x1.12358f08ae5bap+0, -0x1.960cbbf788d5cp-4},
    {0x1.0953f419900a7p+0, -0x1.a6f9db6475fcep-5}, {0x1.0000000000000p+0, 0x0.0p+0},
    {0x1.e608cfd9a47acp-1, 0x1.338ca9f24f53dp-4},  {0x1.ca4b31f026aa0p-1, 0x1.476a9543891bap-3},
    {0x1.b2036576afce6p-1, 0x1.e840b4ac4e4d2p-3},  {0x1.9c2d163a1aa2dp-1, 0x1.40645f0c6651cp-2},
    {0x1.886e6037841edp-1, 0x1.88e9c2c1b9ff8p-2},  {0x1.767dcf5534862p-1, 0x1.ce0a44eb17bccp-2}};
static __device__ const unsigned long long kPowExp2Tab[32] = {
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,
    0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,
    0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,
    0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,
    0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,
    0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,
    0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull};
static __device__ __noinline__ float powf_ref(float x, float y) {
    uint32_t ix = __float_as_uint(x);
    const uint32_t iy = __float_as_uint(y);
    if (2u * iy - 1u > 0xfefffffeu) return powf(x, y);           // y is 0, inf or NaN
    if (ix - 0x00800000u > 0x7effffffu) {
        if (ix == 0u || ix >= 0x00800000u) return powf(x, y);    // x is 0, negative, inf or NaN
        ix = __float_as_uint(x * 0x1p23f) & 0x7fffffffu;         // subnormal: normalise
        ix -= 23u << 23;
    }
    const uint32_t tmp = ix - 0x3f330000u;
    const int i = (tmp >> 19) & 15;
    const uint32_t top = tmp & 0xff800000u;
    const int k = static_cast<int>(top) >> 23;
    const double z = static_cast<double>(__uint_as_float(ix - top));
    const double r = fma(z, kPowLog2Tab[i][0], -1.0);
    const double y0 = kPowLog2Tab[i][1] + static_cast<double>(k);
    const double r2 = r * r;
    double p = fma(0x1.27616c9496e0bp-2, r, -0x1.71969a075c67ap-2);
    const double p1 = fma(0x1.ec70a6ca7baddp-2, r, -0x1.7154748bef6c8p-1);
    double q = fma(0x1.71547652ab82bp+0, r, y0);
    const double r4 = r2 * r2;
    q = fma(p1, r2, q);
    p = fma(p, r4, q);                                           // log2(x)
    const double ylogx = static_cast<double>(y) * p;
    if (((static_cast<unsigned long long>(__double_as_longlong(ylogx)) >> 47) & 0xffffull) > 0x80beull) {  // |y log2 x| >= 126
        if (ylogx > 0x1.fffffffa3aae2p+6) return powf(x, y);     // overflow side
        if (ylogx <= -150.0) return 0.0f;                        // __math_uflowf
        if (ylogx < -149.0) return __uint_as_float(1u);          // __math_may_uflowf: 0x1.4p-75f squared = 2^-149
    }
    double kd = ylogx + 0x1.8p+47;
    const unsigned long long ki = static_cast<unsigned long long>(__double_as_longlong(kd));
    kd -= 0x1.8p+47;
    const double rr = ylogx - kd;
    const double s = __longlong_as_double(static_cast<long long>(kPowExp2Tab[ki & 31ull] + (ki << 47)));
    const double zz = fma(0x1.c6af84b912394p-5, rr, 0x1.ebfce50fac4f3p-3);
    const double rr2 = rr * rr;
    double e = fma(0x1.62e42ff0c52d6p-1, rr, 1.0);
    e = fma(zz, rr2, e);
    return static_cast<float>(e * s);
}

// generateDOFRay; r1, r2 are the two lens draws that follow the jitter draws.
__device__ __forceinline__ Ray dof_ray(const DevFrame& fr, float u, float v, float r1, float r2) {
    const Ray pin = camera_ray(fr, u, v);
    const V3 focus = pin.o + pin.d * fr.focus_dist;
    const float angle = MCSKIN_TWO_PI_F * r1;
    const float radius = fr.aperture * sqrtf(r2);
    float sn, cs;
    sincos_ref(angle, &sn, &cs);
    const float lensX = radius * cs;
    const float lensY = radius * sn;
    const V3 lens = ld3(fr.cam_right) * lensX + ld3(fr.cam_up) * lensY;
    Ray r;
    r.o = ld3(fr.cam_pos) + lens;
    r.d = normalize3(focus - r.o);
    return r;
}

__device__ __forceinline__ float4 flat_background(const DevFrame& fr) {
    return make_float4(fr.background[0], fr.background[1], fr.background[2], fr.background[3]);
}
// RayTracer::backgroundColor with a config
__device__ __forceinline__ float4 config_background(const DevFrame& fr, float u, float v) {
    if (!fr.gradient_bg) return flat_background(fr);
    const float cx = u - 0.5f, cy = v - 0.5f;
    float dist = sqrtf(cx * cx + cy * cy) * 2.0f * fr.gradient_scale;
    dist = clamp01(dist);
    const float t = dist * dist;
    const float k = 1.0f - t;
    return make_float4(fr.bg_center[0] * k + fr.bg_edge[0] * t, fr.bg_center[1] * k + fr.bg_edge[1] * t,
                       fr.bg_center[2] * k + fr.bg_edge[2] * t, 1.0f);
}

// isInShadow from the offset origin point + normal*eps, examining only the boxes of `allow` among
// the first 32 (see bundle_box_mask)
__device__ __forceinline__ bool in_shadow_from(const SceneView& sc, V3 origin, V3 lightPos, uint32_t allow) {
    Ray r;
    r.o = origin;
    const V3 toLight = lightPos - r.o;
    const float dist = len3(toLight);
    if (dist < 1e-6f) return false;
    r.d = div3(toLight, dist);
    return occluded_among(sc, r, dist, allow);
}
__device__ __forceinline__ bool in_shadow_among(const SceneView& sc, V3 point, V3 normal, V3 lightPos, uint32_t allow) {
    return in_shadow_from(sc, point + normal * kShadowEpsilon, lightPos, allow);
}

// isInShadow
__device__ __forceinline__ bool in_shadow(const SceneView& sc, V3 point, V3 normal, V3 lightPos) {
    Ray r;
    r.o = point + normal * kShadowEpsilon;
    const V3 toLight = lightPos - r.o;
    const float dist = len3(toLight);
    if (dist < 1e-6f) return false;
    r.d = div3(toLight, dist);
    return occluded(sc, r, dist);
}

// Orthonormal frame of computeSoftShadow / computeAO: t = normalize(e x n), b = n x t
__device__ __forceinline__ void frame_about(V3 n, V3* t, V3* b) {
    if (fabsf(n.x) < 0.9f) *t = normalize3(cross3(mk3(1.0f, 0.0f, 0.0f), n));
    else *t = normalize3(cross3(mk3(0.0f, 1.0f, 0.0f), n));
    *b = cross3(n, *t);
}

template <class Engine>
__device__ __forceinline__ int soft_shadow_count(const SceneView& sc, const DevFrame& fr, V3 point, V3 normal,
                                                 int samples, Engine& rng, uint32_t allow) {
    const V3 lp = ld3(fr.light_pos);
    const V3 toPoint = normalize3(point - lp);
    V3 tangent, bitangent;
    frame_about(toPoint, &tangent, &bitangent);
    int lit = 0;
    for (int i = 0; i < samples; ++i) {
        const float angle = MCSKIN_TWO_PI_F * rng.next();
        const float r = fr.light_radius * sqrtf(rng.next());
        float sn, cs;
        sincos_ref(angle, &sn, &cs);
        const V3 offset = tangent * (r * cs) + bitangent * (r * sn);
        const V3 samplePos = lp + offset;
        if (!in_shadow_among(sc, point, normal, samplePos, allow)) ++lit;
    }
    return lit;
}

// computeSoftShadow
static __device__ __noinline__ float soft_shadow_large(const SceneView& sc, const DevFrame& fr, V3 point, V3 normal,
                                                int samples, uint32_t seed) {
    if (samples <= 1 || fr.light_radius < 1e-4f) return in_shadow(sc, point, normal, ld3(fr.light_pos)) ? 0.0f : 1.0f;
    LocalEngine rng;
    rng.seed(seed);
    return static_cast<float>(soft_shadow_count(sc, fr, point, normal, samples, rng, 0xffffffffu)) / static_cast<float>(samples);
}
__device__ __forceinline__ float soft_shadow(const SceneView& sc, const DevFrame& fr, V3 point, V3 normal,
                                             int samples, uint32_t seed) {
    if (samples <= 1 || fr.light_radius < 1e-4f || 2 * samples > kFreshStreamMaxDraws)
        return soft_shadow_large(sc, fr, point, normal, samples, seed);  // rare configurations, out of line
    // all shadow rays of this hit start at the same point and end on the light's disk: pre-select once
    // the boxes that bundle can reach; with none in reach every ray is lit and no sample is drawn
    const uint32_t allow = bundle_box_mask(sc, point + normal * kShadowEpsilon, ld3(fr.light_pos), fr.light_radius);
    if (allow == 0u && sc.n_boxes <= 32) return 1.0f;  // samples / samples
    FreshStream rng;
    rng.seed(seed);
    return static_cast<float>(soft_shadow_count(sc, fr, point, normal, samples, rng, allow)) / static_cast<float>(samples);
}

template <class Engine>
__device__ __forceinline__ int ao_count(const SceneView& sc, V3 point, V3 normal, int samples, float radius,
                                        Engine& rng) {
    const V3 N = normalize3(normal);
    V3 T, B;
    frame_about(N, &T, &B);
    Ray r;
    r.o = point + N * 1e-3f;
    int occludedCount = 0;
    for (int i = 0; i < samples; ++i) {
        const float r1 = rng.next();
        const float r2 = rng.next();
        const float sinT = sqrtf(1.0f - r1);
        const float cosT = sqrtf(r1);
        const float phi = MCSKIN_TWO_PI_F * r2;
        float sn, cs;
        sincos_ref(phi, &sn, &cs);
        const V3 l = mk3(sinT * cs, cosT, sinT * sn);
        r.d = normalize3((T * l.x + N * l.y) + B * l.z);
        if (occluded(sc, r, radius)) ++occludedCount;  // hit.hit && hit.t < radius
    }
    return occludedCount;
}
static __device__ __noinline__ float ambient_occlusion_large(const SceneView& sc, V3 point, V3 normal, int samples,
                                                      float radius, uint32_t seed) {
    LocalEngine rng;
    rng.seed(seed);
    return 1.0f - static_cast<float>(ao_count(sc, point, normal, samples, radius, rng)) / static_cast<float>(samples);
}
// RayTracer::computeAO
static __device__ __noinline__ float ambient_occlusion(const SceneView& sc, V3 point, V3 normal, int samples,
                                                   float radius, uint32_t seed) {
    if (2 * samples > kFreshStreamMaxDraws) return ambient_occlusion_large(sc, point, normal, samples, radius, seed);
    FreshStream rng;
    rng.seed(seed);
    return 1.0f - static_cast<float>(ao_count(sc, point, normal, samples, radius, rng)) / static_cast<float>(samples);
}

// shade() with the visibility already known (shading.cpp:62-96 after its shadow test)
__device__ __forceinline__ float4 shade_lit(const DevFrame& fr, V3 P, V3 normal, float4 tex, V3 viewDir, float vis) {
    const V3 lp = ld3(fr.light_pos);
    const V3 L = normalize3(lp - P);
    const V3 N = normalize3(normal);
    const V3 V = normalize3(viewDir);
    const float ndl = fmaxf(0.0f, dot3(N, L));
    const float kdiff = fr.kd * ndl * vis;
    const V3 H = normalize3(L + V);
    const float ndh = fmaxf(0.0f, dot3(N, H));
    const float spec = powf_ref(ndh, fr.shininess);  // std::pow(NdotH, shininess), shading.cpp:90
    const float kspec = fr.ks * spec * vis;
    float4 out;
    out.x = (tex.x * fr.ambient + tex.x * fr.light_color[0] * kdiff) + fr.light_color[0] * kspec;
    out.y = (tex.y * fr.ambient + tex.y * fr.light_color[1] * kdiff) + fr.light_color[1] * kspec;
    out.z = (tex.z * fr.ambient + tex.z * fr.light_color[2] * kdiff) + fr.light_color[2] * kspec;
    out.w = tex.w;
    return clamp4(out);
}

// shade(): Blinn-Phong with the visibility factor (negative = hard shadow test inside)
__device__ __forceinline__ float4 shade_hit(const SceneView& sc, const DevFrame& fr, V3 P, V3 normal, float4 tex,
                                            V3 viewDir, float shadowFactor) {
    float vis = shadowFactor;
    if (vis < 0.0f) vis = in_shadow(sc, P, normalize3(normal), ld3(fr.light_pos)) ? 0.0f : 1.0f;
    return shade_lit(fr, P, normal, tex, viewDir, vis);
}

// The points on the light's disk computeSoftShadow samples for one hit, in order
// (shading.cpp:35-52): frame at the light facing the point, then per sample
// angle = 2*pi*draw, r = radius*sqrt(draw), position = light + t*(r cos) + b*(r sin).
template <class Engine>
__device__ __forceinline__ void soft_shadow_positions(const DevFrame& fr, V3 point, int samples, Engine& rng,
                                                      float* out /* 3 floats per sample */) {
    const V3 lp = ld3(fr.light_pos);
    const V3 toPoint = normalize3(point - lp);
    V3 tangent, bitangent;
    frame_about(toPoint, &tangent, &bitangent);
    for (int i = 0; i < samples; ++i) {
        const float angle = MCSKIN_TWO_PI_F * rng.next();
        const float r = fr.light_radius * sqrtf(rng.next());
        float sn, cs;
        sincos_ref(angle, &sn, &cs);
        const V3 offset = tangent * (r * cs) + bitangent * (r * sn);
        const V3 samplePos = lp + offset;
        out[3 * i] = samplePos.x;
        out[3 * i + 1] = samplePos.y;
        out[3 * i + 2] = samplePos.z;
    }
}

// The soft-shadow / AO seeds of traceRay (raytracer.cpp:110-112, :122-123)
__device__ __forceinline__ uint32_t shadow_seed(V3 P, int depth) {
    return seed_cast(P.x * 12345.0f + P.y * 67890.0f + P.z * 11111.0f + static_cast<float>(depth) * 99999.0f);
}
__device__ __forceinline__ uint32_t ao_seed(V3 P) {
    return seed_cast(P.x * 73856093.0f + P.y * 19349663.0f + P.z * 83492791.0f);
}

// traceRay(ray, scene, startDepth, maxBounces, params, config) including the
// tile renderer's primary-miss override when (u, v) are supplied (tile_renderer.cpp:106-114).
struct TraceOptions {
    int start_depth;     // 0 for camera rays
    bool primary_uv;     // true: a depth-0 miss resolves to backgroundColor(u, v)
    float u, v;
};

__device__ __forceinline__ float4 trace_path(const SceneView& sc, const DevFrame& fr, Ray ray,
                                             const TraceOptions& opt) {
    const bool cfg = fr.use_config != 0;
    const int maxB = fr.max_bounces;
    int depth = opt.start_depth;

    if (depth > maxB) {  // raytracer.cpp:86-90
        float4 c = cfg ? config_background(fr, 0.5f, 0.5f) : flat_background(fr);
        if (opt.primary_uv && closest_hit(sc, ray).box < 0) c = config_background(fr, opt.u, opt.v);
        return c;
    }

    float stack[kMaxStackDepth][3];  // shaded rgb of every level that spawned a reflection
    float alpha0 = 1.0f;             // only the outermost level's alpha survives the fold
    int top = 0;
    float4 tail;                     // colour returned by the deepest call
    for (;;) {
        const Hit hit = closest_hit(sc, ray);
        if (hit.box < 0) {  // raytracer.cpp:94-102
            if (depth == 0 && opt.primary_uv) tail = config_background(fr, opt.u, opt.v);
            else if (depth == 0 && cfg) tail = config_background(fr, 0.5f, 0.5f);
            else tail = flat_background(fr);
            break;
        }
        const V3 P = hit.p;
        const V3 nrm = hit_normal(sc, hit);
        const float4 tex = hit_texel(sc, hit);
        const V3 viewDir = normalize3(ray.o - P);
        float shadowFactor = -1.0f;
        if (cfg && fr.soft_on) {
            const uint32_t seed = shadow_seed(P, depth);
            shadowFactor = soft_shadow(sc, fr, P, nrm, fr.shadow_samples, seed);
        }
        float4 shaded = shade_hit(sc, fr, P, nrm, tex, viewDir, shadowFactor);
        const float alpha = shaded.w;
        if (cfg && fr.ao_on && depth == 0) {
            const uint32_t seed = ao_seed(P);
            const float ao = ambient_occlusion(sc, P, nrm, fr.ao_samples, fr.ao_radius, seed);
            const float f = 1.0f - fr.ao_intensity * (1.0f - ao);
            shaded.x *= f;
            shaded.y *= f;
            shaded.z *= f;
        }
        if (depth < maxB && top < kMaxStackDepth) {
            stack[top][0] = shaded.x;
            stack[top][1] = shaded.y;
            stack[top][2] = shaded.z;
            if (top == 0) alpha0 = alpha;
            ++top;
            const V3 N = normalize3(nrm);
            const V3 D = normalize3(ray.d);
            V3 R = D - N * (2.0f * dot3(D, N));
            R = normalize3(R);
            ray.o = P + N * kReflectEpsilon;
            ray.d = R;
            ++depth;
            continue;
        }
        shaded.w = alpha;
        tail = clamp4(shaded);
        break;
    }
    // fold back: shaded*(1-0.1f) + reflected*0.1f, alpha restored, clamp (raytracer.cpp:143-147)
    const float keep = 1.0f - kReflectivity;
    while (top > 0) {
        --top;
        float4 c;
        c.x = stack[top][0] * keep + tail.x * kReflectivity;
        c.y = stack[top][1] * keep + tail.y * kReflectivity;
        c.z = stack[top][2] * keep + tail.z * kReflectivity;
        c.w = alpha0;  // deeper levels' alpha is overwritten by the caller's (raytracer.cpp:146)
        tail = clamp4(c);
    }
    return tail;
}

}  // namespace mcskin
