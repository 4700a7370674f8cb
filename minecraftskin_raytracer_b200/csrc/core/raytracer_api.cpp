// raytracer_api.cpp — the reference's free functions of the hot path, evaluated on the GPU.
//
// intersectMesh / intersectScene (intersection.h:14-18), isInShadow / computeSoftShadow /
// shade (shading.h:20-37), RayTracer::traceRay / backgroundColor / computeAO
// (raytracer.h:41-54) and Camera::generateRay (scene.h:25) keep their signatures and run the
// same device code as the frame renderer, one query per call, through the single-ray entry
// points of include/mcskin_cuda.h.  They exist so the reference's unit tests and tools link
// and behave the same; a frame is rendered with TileRenderer::render, not by calling these in
// a loop.  On failure (no device) they return the reference's "miss" values.
#include <cstring>

#include "mcskin/detail/flatten.hpp"
#include "mcskin_cuda.h"
#include "raytracer/intersection.h"
#include "raytracer/raytracer.h"
#include "raytracer/shading.h"

namespace {

McRay toRay(const Ray& r) {
    return McRay{{r.origin.x, r.origin.y, r.origin.z}, {r.direction.x, r.direction.y, r.direction.z}};
}
HitResult toHit(const McHit& h) {
    HitResult r;
    r.hit = h.hit != 0;
    if (!r.hit) return r;
    r.t = h.t;
    r.point = Vec3(h.point[0], h.point[1], h.point[2]);
    r.normal = Vec3(h.normal[0], h.normal[1], h.normal[2]);
    r.textureColor = Color(h.tex_color[0], h.tex_color[1], h.tex_color[2], h.tex_color[3]);
    r.isOuterLayer = h.is_outer_layer != 0;
    return r;
}
McHit fromHit(const HitResult& h) {
    McHit m;
    std::memset(&m, 0, sizeof(m));
    m.hit = h.hit ? 1 : 0;
    m.t = h.t;
    m.point[0] = h.point.x; m.point[1] = h.point.y; m.point[2] = h.point.z;
    m.normal[0] = h.normal.x; m.normal[1] = h.normal.y; m.normal[2] = h.normal.z;
    m.tex_color[0] = h.textureColor.r; m.tex_color[1] = h.textureColor.g;
    m.tex_color[2] = h.textureColor.b; m.tex_color[3] = h.textureColor.a;
    m.is_outer_layer = h.isOuterLayer ? 1 : 0;
    m.box = m.face = -1;
    return m;
}
McConfig configOf(const RayTracer::Config* config, int maxBounces, const ShadingParams& p) {
    McConfig c = config ? mcskin::flattenConfig(*config) : mcskin::flattenConfig(RayTracer::Config{});
    c.max_bounces = maxBounces;
    c.kd = p.kd; c.ks = p.ks; c.ambient = p.ambient; c.shininess = p.shininess;
    return c;
}
Color toColor(const float* c) { return Color(c[0], c[1], c[2], c[3]); }

}  // namespace

HitResult intersectScene(const Ray& ray, const Scene& scene) {
    mcskin::FlatScene flat;
    mcskin::flattenScene(scene, flat);
    const McRay r = toRay(ray);
    McHit h{};
    if (mcskin_cuda_intersect(&flat.scene, 0, -1, &r, 1, &h) != MC_OK) return HitResult();
    return toHit(h);
}

HitResult intersectMesh(const Ray& ray, const Mesh& mesh) {
    Scene one;
    one.meshes.push_back(mesh);
    mcskin::FlatScene flat;
    mcskin::flattenScene(one, flat);
    const McRay r = toRay(ray);
    McHit h{};
    if (mcskin_cuda_intersect(&flat.scene, 0, 0, &r, 1, &h) != MC_OK) return HitResult();
    return toHit(h);
}

bool isInShadow(const Vec3& point, const Vec3& normal, const Vec3& lightPos, const Scene& scene) {
    mcskin::FlatScene flat;
    mcskin::flattenScene(scene, flat);
    const float p[3] = {point.x, point.y, point.z}, n[3] = {normal.x, normal.y, normal.z},
                l[3] = {lightPos.x, lightPos.y, lightPos.z};
    int32_t out = 0;
    if (mcskin_cuda_in_shadow(&flat.scene, 0, p, n, l, 1, &out) != MC_OK) return false;
    return out != 0;
}

float computeSoftShadow(const Vec3& point, const Vec3& normal, const Light& light, const Scene& scene, int samples,
                        unsigned int seed) {
    Scene lit = scene;  // the light is a parameter of its own in the reference's signature
    lit.light = light;
    mcskin::FlatScene flat;
    mcskin::flattenScene(lit, flat);
    const float p[3] = {point.x, point.y, point.z}, n[3] = {normal.x, normal.y, normal.z};
    float out = 1.0f;
    mcskin_cuda_soft_shadow(&flat.scene, 0, p, n, &seed, samples, 1, &out);
    return out;
}

Color shade(const HitResult& hit, const Vec3& viewDir, const Light& light, const Scene& scene,
            const ShadingParams& params, float shadowFactor) {
    Scene lit = scene;
    lit.light = light;
    mcskin::FlatScene flat;
    mcskin::flattenScene(lit, flat);
    const McConfig cfg = configOf(nullptr, 0, params);
    const McHit h = fromHit(hit);
    const float v[3] = {viewDir.x, viewDir.y, viewDir.z};
    float out[4] = {0, 0, 0, 1};
    mcskin_cuda_shade(&flat.scene, &cfg, 0, &h, v, &shadowFactor, 1, out);
    return toColor(out);
}

Color RayTracer::traceRay(const Ray& ray, const Scene& scene, int depth, int maxBounces, const ShadingParams& params,
                          const Config* config) {
    mcskin::FlatScene flat;
    mcskin::flattenScene(scene, flat);
    const McConfig cfg = configOf(config, maxBounces, params);
    const McRay r = toRay(ray);
    float out[4] = {0, 0, 0, 1};
    mcskin_cuda_trace(&flat.scene, &cfg, 0, config ? 1 : 0, depth, &r, 1, out);
    return toColor(out);
}

Color RayTracer::backgroundColor(const Scene& scene, float u, float v, const Config* config) {
    mcskin::FlatScene flat;
    mcskin::flattenScene(scene, flat);
    const McConfig cfg = configOf(config, 0, ShadingParams{});
    const float uv[2] = {u, v};
    float out[4] = {0, 0, 0, 1};
    mcskin_cuda_background(&flat.scene, &cfg, 0, config ? 1 : 0, uv, 1, out);
    return toColor(out);
}

float RayTracer::computeAO(const Vec3& point, const Vec3& normal, const Scene& scene, int samples, float radius,
                           unsigned int seed) {
    mcskin::FlatScene flat;
    mcskin::flattenScene(scene, flat);
    const float p[3] = {point.x, point.y, point.z}, n[3] = {normal.x, normal.y, normal.z};
    float out = 1.0f;
    mcskin_cuda_ambient_occlusion(&flat.scene, 0, p, n, &seed, samples, radius, 1, &out);
    return out;
}

Ray Camera::generateRay(float u, float v, float aspectRatio) const {
    Scene s;
    s.camera = *this;
    mcskin::FlatScene flat;
    mcskin::flattenScene(s, flat);
    const float uv[2] = {u, v};
    McRay r{};
    if (mcskin_cuda_generate_rays(&flat.scene, 0, aspectRatio, uv, 1, &r) != MC_OK) return Ray(position, Vec3());
    return Ray(Vec3(r.origin[0], r.origin[1], r.origin[2]), Vec3(r.dir[0], r.dir[1], r.dir[2]));
}
