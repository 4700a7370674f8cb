// ref_shim.cpp — TEST INFRASTRUCTURE ONLY.  Never linked into the product.
//
// A C-ABI shim around the UNMODIFIED reference core (the 10 Qt-free sources of
// /root/reference/src/CMakeLists.txt:2-14 minus gui/camera_controller.cpp),
// compiled where they lie by oracle/build_ref.sh into oracle/_ref/libmcskin_ref.so.
// It lets the tests (a) run the reference's own TileRenderer::render /
// intersectScene / shade / traceRay on exactly the flat scenes the CUDA path and
// the C restatement (oracle/mcskin_oracle.c) consume, (b) build scenes with the
// reference's SkinParser + MeshBuilder to check mcskin_build_skin_scene, and
// (c) time the reference's multithreaded CPU renderer (bench.py --impl reference).
//
// No reference source text is copied here: the file only #includes the
// reference headers at build time and calls their public functions.
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <random>
#include <string>
#include <thread>
#include <unistd.h>
#include <vector>

// reference headers (-I /root/reference/src -I /root/reference/third_party)
#include "math/vec3.h"
#include "output/image_writer.h"
#include "raytracer/intersection.h"
#include "raytracer/raytracer.h"
#include "raytracer/shading.h"
#include "raytracer/tile_renderer.h"
#include "scene/mesh_builder.h"
#include "scene/pose.h"
#include "scene/scene.h"
#include "skin/skin_parser.h"
#include <stb/stb_image_write.h>

// product-side templates over the scene types (include/mcskin/detail/flatten.hpp)
#include "mcskin/detail/flatten.hpp"

// ---- intersectScene call counter (ld --wrap; see build_ref.sh) ----------------
static std::atomic<long long> g_calls{0};
static std::atomic<int> g_counting{0};
extern "C" HitResult __real__Z14intersectSceneRK3RayRK5Scene(const Ray&, const Scene&);
extern "C" HitResult __wrap__Z14intersectSceneRK3RayRK5Scene(const Ray& r, const Scene& s) {
    if (g_counting.load(std::memory_order_relaxed)) g_calls.fetch_add(1, std::memory_order_relaxed);
    return __real__Z14intersectSceneRK3RayRK5Scene(r, s);
}

namespace {

struct RefScene {
    Scene scene;
    // storage for textures of scenes rebuilt from a flat description
    std::vector<std::unique_ptr<TextureRegion>> textures;
};

Pose poseFrom(const float* p) {
    Pose pose;
    if (p) {
        pose.head = {p[0], p[1]};
        pose.body = {p[2], p[3]};
        pose.rightArm = {p[4], p[5]};
        pose.leftArm = {p[6], p[7]};
        pose.rightLeg = {p[8], p[9]};
        pose.leftLeg = {p[10], p[11]};
    }
    return pose;
}

RayTracer::Config configFrom(const McConfig& c) {
    RayTracer::Config k;
    k.width = c.width;
    k.height = c.height;
    k.maxBounces = c.max_bounces;
    k.samplesPerPixel = c.samples_per_pixel;
    k.tileSize = c.tile_size;
    k.threadCount = c.thread_count;
    k.softShadows = c.soft_shadows != 0;
    k.shadowSamples = c.shadow_samples;
    k.aoEnabled = c.ao_enabled != 0;
    k.aoSamples = c.ao_samples;
    k.aoRadius = c.ao_radius;
    k.aoIntensity = c.ao_intensity;
    k.dofEnabled = c.dof_enabled != 0;
    k.aperture = c.aperture;
    k.focusDistance = c.focus_distance;
    k.gradientBg = c.gradient_bg != 0;
    k.gradientScale = c.gradient_scale;
    k.bgCenter = Color(c.bg_center[0], c.bg_center[1], c.bg_center[2], c.bg_center[3]);
    k.bgEdge = Color(c.bg_edge[0], c.bg_edge[1], c.bg_edge[2], c.bg_edge[3]);
    return k;
}

ShadingParams paramsFrom(const McConfig& c) {
    ShadingParams p;
    p.kd = c.kd;
    p.ks = c.ks;
    p.ambient = c.ambient;
    p.shininess = c.shininess;
    return p;
}

// The 12 triangles of a box whose vertex min/max are exactly (lo, hi); vertex
// order follows the face order the reference's face index refers to
// (-Z,+Z,+X,-X,+Y,-Y), two triangles per face.
void boxTriangles(const float lo[3], const float hi[3], const TextureRegion* const faceTex[6],
                  int nTriangles, std::vector<Triangle>& out) {
    const Vec3 c[8] = {
        Vec3(lo[0], lo[1], lo[2]), Vec3(hi[0], lo[1], lo[2]), Vec3(lo[0], hi[1], lo[2]), Vec3(hi[0], hi[1], lo[2]),
        Vec3(lo[0], lo[1], hi[2]), Vec3(hi[0], lo[1], hi[2]), Vec3(lo[0], hi[1], hi[2]), Vec3(hi[0], hi[1], hi[2])};
    // corner index = x + 2y + 4z
    static const int quad[6][4] = {{2, 3, 1, 0}, {7, 6, 4, 5}, {3, 7, 5, 1}, {6, 2, 0, 4}, {6, 7, 3, 2}, {0, 1, 5, 4}};
    static const float nrm[6][3] = {{0, 0, -1}, {0, 0, 1}, {1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}};
    out.clear();
    for (int f = 0; f < 6; ++f) {
        for (int half = 0; half < 2; ++half) {
            if (static_cast<int>(out.size()) >= nTriangles) return;
            Triangle t;
            t.v0 = c[quad[f][0]];
            t.v1 = c[quad[f][half ? 2 : 1]];
            t.v2 = c[quad[f][half ? 3 : 2]];
            t.normal = Vec3(nrm[f][0], nrm[f][1], nrm[f][2]);
            t.u0 = 0; t.v0_uv = 0; t.u1 = 1; t.v1_uv = half ? 1.f : 0.f; t.u2 = half ? 0.f : 1.f; t.v2_uv = 1;
            t.texture = faceTex[f];
            out.push_back(t);
        }
    }
}

RefScene* sceneFromFlat(const McScene& f) {
    auto rs = std::make_unique<RefScene>();
    Scene& sc = rs->scene;
    sc.meshes.reserve(f.n_boxes);
    for (int b = 0; b < f.n_boxes; ++b) {
        const McBox& box = f.boxes[b];
        const TextureRegion* faceTex[6];
        for (int k = 0; k < 6; ++k) {
            const McFaceTex& ft = box.face[k];
            if (ft.texel_offset < 0) {
                faceTex[k] = nullptr;
                continue;
            }
            auto tr = std::make_unique<TextureRegion>();
            if (ft.width > 0 && ft.height > 0) {
                tr->width = ft.width;
                tr->height = ft.height;
                tr->pixels.resize(static_cast<size_t>(ft.width) * ft.height);
                for (size_t i = 0; i < tr->pixels.size(); ++i) {
                    const float* p = f.texels_rgba + (static_cast<size_t>(ft.texel_offset) + i) * 4;
                    tr->pixels[i] = Color(p[0], p[1], p[2], p[3]);
                }
            }
            faceTex[k] = tr.get();
            rs->textures.push_back(std::move(tr));
        }
        Mesh m;
        m.isOuterLayer = box.is_outer_layer != 0;
        m.hasRotation = box.has_rotation != 0;
        m.pivot = Vec3(box.pivot[0], box.pivot[1], box.pivot[2]);
        m.rotX = box.rot_x_deg;
        m.rotZ = box.rot_z_deg;
        // Mesh::triangles carries the texture pointers even for posed meshes
        // (intersection.cpp:124-129); the bounds list is localTriangles when posed.
        std::vector<Triangle> tris;
        boxTriangles(box.bounds_min, box.bounds_max, faceTex, 12, tris);
        m.triangles = tris;
        if (m.hasRotation) {
            boxTriangles(box.bounds_min, box.bounds_max, faceTex, box.n_triangles, tris);
            m.localTriangles = tris;
        } else if (box.n_triangles < 12) {
            boxTriangles(box.bounds_min, box.bounds_max, faceTex, box.n_triangles, tris);
            m.triangles = tris;
        }
        sc.meshes.push_back(std::move(m));
    }
    sc.light.position = Vec3(f.light_pos[0], f.light_pos[1], f.light_pos[2]);
    sc.light.color = Color(f.light_color[0], f.light_color[1], f.light_color[2], f.light_color[3]);
    sc.light.radius = f.light_radius;
    sc.camera.position = Vec3(f.cam_pos[0], f.cam_pos[1], f.cam_pos[2]);
    sc.camera.target = Vec3(f.cam_target[0], f.cam_target[1], f.cam_target[2]);
    sc.camera.up = Vec3(f.cam_up[0], f.cam_up[1], f.cam_up[2]);
    sc.camera.fov = f.cam_fov_deg;
    sc.backgroundColor = Color(f.background[0], f.background[1], f.background[2], f.background[3]);
    return rs.release();
}

Vec3 unrotate(const Mesh& m, Vec3 p, bool isPoint) {
    // world -> local, own arithmetic (only used to classify which face an AOV hit lies on)
    const double d2r = 3.14159265358979323846 / 180.0;
    Vec3 piv = isPoint ? m.pivot : Vec3(0, 0, 0);
    double x = p.x - piv.x, y = p.y - piv.y, z = p.z - piv.z;
    if (std::fabs(m.rotZ) > 0.01f) {
        double a = -m.rotZ * d2r, c = std::cos(a), s = std::sin(a);
        double nx = x * c - y * s, ny = x * s + y * c;
        x = nx; y = ny;
    }
    if (std::fabs(m.rotX) > 0.01f) {
        double a = -m.rotX * d2r, c = std::cos(a), s = std::sin(a);
        double ny = y * c - z * s, nz = y * s + z * c;
        y = ny; z = nz;
    }
    return Vec3(float(x + piv.x), float(y + piv.y), float(z + piv.z));
}

// (box, face) of a hit the reference reports only as point+normal: the face is the
// one whose plane the local hit point lies on, along the dominant axis of the local normal.
void classifyHit(const Scene& sc, const Ray& ray, int onlyBox, HitResult& best, int& boxOut, int& faceOut) {
    best = HitResult();
    best.hit = false;
    best.t = std::numeric_limits<float>::max();
    boxOut = -1;
    faceOut = -1;
    for (int b = 0; b < static_cast<int>(sc.meshes.size()); ++b) {
        if (onlyBox >= 0 && b != onlyBox) continue;
        HitResult h = intersectMesh(ray, sc.meshes[b]);
        if (h.hit && (onlyBox >= 0 || h.t < best.t)) {
            best = h;
            boxOut = b;
        }
    }
    if (boxOut < 0) {
        best.hit = false;
        return;
    }
    const Mesh& m = sc.meshes[boxOut];
    Vec3 lp = best.point, ln = best.normal;
    if (m.hasRotation) {
        lp = unrotate(m, lp, true);
        ln = unrotate(m, ln, false);
    }
    float lo[3], hi[3];
    mcskin::triangleBounds(m.hasRotation ? m.localTriangles : m.triangles, lo, hi);
    const float n[3] = {std::fabs(ln.x), std::fabs(ln.y), std::fabs(ln.z)};
    const float p[3] = {lp.x, lp.y, lp.z};
    int axis = 0;
    if (n[1] > n[axis]) axis = 1;
    if (n[2] > n[axis]) axis = 2;
    const bool maxSide = std::fabs(p[axis] - hi[axis]) <= std::fabs(p[axis] - lo[axis]);
    static const int faceOf[3][2] = {{3, 2}, {5, 4}, {0, 1}};  // [axis][maxSide]
    faceOut = faceOf[axis][maxSide ? 1 : 0];
}

void fillHit(const HitResult& h, int box, int face, McHit& o) {
    std::memset(&o, 0, sizeof(o));
    o.hit = h.hit ? 1 : 0;
    o.box = -1;
    o.face = -1;
    if (!h.hit) return;
    o.t = h.t;
    o.point[0] = h.point.x; o.point[1] = h.point.y; o.point[2] = h.point.z;
    o.normal[0] = h.normal.x; o.normal[1] = h.normal.y; o.normal[2] = h.normal.z;
    o.tex_color[0] = h.textureColor.r; o.tex_color[1] = h.textureColor.g;
    o.tex_color[2] = h.textureColor.b; o.tex_color[3] = h.textureColor.a;
    o.is_outer_layer = h.isOuterLayer ? 1 : 0;
    o.box = box;
    o.face = face;
}

HitResult hitFrom(const McHit& h) {
    HitResult r;
    r.hit = h.hit != 0;
    r.t = h.t;
    r.point = Vec3(h.point[0], h.point[1], h.point[2]);
    r.normal = Vec3(h.normal[0], h.normal[1], h.normal[2]);
    r.textureColor = Color(h.tex_color[0], h.tex_color[1], h.tex_color[2], h.tex_color[3]);
    r.isOuterLayer = h.is_outer_layer != 0;
    return r;
}

}  // namespace

extern "C" {

void* mcref_scene_from_flat(const McScene* flat) { return flat ? sceneFromFlat(*flat) : nullptr; }

// SkinParser::parse needs a file: the atlas is written with the reference's own stb.
void* mcref_scene_from_atlas(const uint8_t* rgba, int w, int h, const float* pose12) {
    char path[128];
    static std::atomic<int> serial{0};
    std::snprintf(path, sizeof(path), "/tmp/mcref_atlas_%d_%d.png", static_cast<int>(getpid()), serial.fetch_add(1));
    if (!stbi_write_png(path, w, h, 4, rgba, w * 4)) return nullptr;
    auto parsed = SkinParser::parse(path);
    std::remove(path);
    if (!parsed.isOk()) return nullptr;
    auto rs = std::make_unique<RefScene>();
    rs->scene = MeshBuilder::buildScene(*parsed.value, poseFrom(pose12));
    return rs.release();
}

void* mcref_default_scene(const float* pose12) {
    auto rs = std::make_unique<RefScene>();
    rs->scene = MeshBuilder::buildDefaultScene(poseFrom(pose12));
    return rs.release();
}

void mcref_scene_free(void* s) { delete static_cast<RefScene*>(s); }

int mcref_builtin_pose_count(void) { return static_cast<int>(getBuiltinPoses().size()); }
int mcref_builtin_pose(int index, float* pose12) {
    auto poses = getBuiltinPoses();
    if (index < 0 || index >= static_cast<int>(poses.size())) return -1;
    const Pose& p = poses[index];
    const PartPose* parts[6] = {&p.head, &p.body, &p.rightArm, &p.leftArm, &p.rightLeg, &p.leftLeg};
    for (int i = 0; i < 6; ++i) {
        pose12[2 * i] = parts[i]->rotX;
        pose12[2 * i + 1] = parts[i]->rotZ;
    }
    return 0;
}

// Flattens with the product's walk.  Returns 0, or -1 when a capacity is too small
// (n_boxes / n_texels are still written so the caller can size and retry).
int mcref_scene_flatten(void* s, McBox* boxes, int capBoxes, float* texels, int capTexels, McScene* out) {
    mcskin::FlatScene fs;
    mcskin::flattenScene(static_cast<RefScene*>(s)->scene, fs);
    *out = fs.scene;
    out->boxes = boxes;
    out->texels_rgba = texels;
    if (fs.scene.n_boxes > capBoxes || fs.scene.n_texels > capTexels) return -1;
    if (!fs.boxes.empty()) std::memcpy(boxes, fs.boxes.data(), fs.boxes.size() * sizeof(McBox));
    if (!fs.texels.empty()) std::memcpy(texels, fs.texels.data(), fs.texels.size() * sizeof(float));
    return 0;
}

// TileRenderer::render with the reference's own thread pool.  calls (nullable):
// number of intersectScene invocations (counting slows the run; pass NULL when timing).
int mcref_render(void* s, const McConfig* cfg, float* outRgba, long long* calls) {
    RayTracer::Config k = configFrom(*cfg);
    g_calls.store(0);
    g_counting.store(calls ? 1 : 0);
    Image img = TileRenderer::render(static_cast<RefScene*>(s)->scene, k);
    g_counting.store(0);
    if (calls) *calls = g_calls.load();
    if (outRgba && !img.pixels.empty()) std::memcpy(outRgba, img.pixels.data(), img.pixels.size() * sizeof(Color));
    return static_cast<int>(TileRenderer::lastErrors().size());
}

int mcref_render_tile(void* s, const McConfig* cfg, const McTile* tile, float* imageRgba) {
    RayTracer::Config k = configFrom(*cfg);
    Image img(k.width, k.height);
    std::memcpy(img.pixels.data(), imageRgba, img.pixels.size() * sizeof(Color));
    Tile t{tile->x, tile->y, tile->width, tile->height};
    TileRenderer::renderTile(t, static_cast<RefScene*>(s)->scene, k, img);
    std::memcpy(imageRgba, img.pixels.data(), img.pixels.size() * sizeof(Color));
    return 0;
}

int mcref_generate_tiles(int w, int h, int tileSize, McTile* out, int capacity) {
    auto tiles = TileRenderer::generateTiles(w, h, tileSize);
    for (int i = 0; i < static_cast<int>(tiles.size()) && i < capacity && out; ++i)
        out[i] = McTile{tiles[i].x, tiles[i].y, tiles[i].width, tiles[i].height};
    return static_cast<int>(tiles.size());
}

int mcref_intersect(void* s, int box, const McRay* rays, int n, McHit* out) {
    const Scene& sc = static_cast<RefScene*>(s)->scene;
    for (int i = 0; i < n; ++i) {
        Ray r(Vec3(rays[i].origin[0], rays[i].origin[1], rays[i].origin[2]),
              Vec3(rays[i].dir[0], rays[i].dir[1], rays[i].dir[2]));
        // the reported HitResult is the reference's own; classifyHit only adds (box, face)
        HitResult viaScene = box >= 0 ? intersectMesh(r, sc.meshes[box]) : intersectScene(r, sc);
        HitResult viaMeshes;
        int b, f;
        classifyHit(sc, r, box, viaMeshes, b, f);
        fillHit(viaScene, b, f, out[i]);
    }
    return 0;
}

int mcref_trace(void* s, const McConfig* cfg, int useConfig, int depth, const McRay* rays, int n, float* outRgba) {
    const Scene& sc = static_cast<RefScene*>(s)->scene;
    RayTracer::Config k = configFrom(*cfg);
    ShadingParams p = paramsFrom(*cfg);
    for (int i = 0; i < n; ++i) {
        Ray r(Vec3(rays[i].origin[0], rays[i].origin[1], rays[i].origin[2]),
              Vec3(rays[i].dir[0], rays[i].dir[1], rays[i].dir[2]));
        Color c = RayTracer::traceRay(r, sc, depth, k.maxBounces, p, useConfig ? &k : nullptr);
        outRgba[4 * i] = c.r; outRgba[4 * i + 1] = c.g; outRgba[4 * i + 2] = c.b; outRgba[4 * i + 3] = c.a;
    }
    return 0;
}

int mcref_shade(void* s, const McConfig* cfg, const McHit* hits, const float* viewDirs, const float* shadowFactors,
                int n, float* outRgba) {
    const Scene& sc = static_cast<RefScene*>(s)->scene;
    ShadingParams p = paramsFrom(*cfg);
    for (int i = 0; i < n; ++i) {
        Color c = shade(hitFrom(hits[i]), Vec3(viewDirs[3 * i], viewDirs[3 * i + 1], viewDirs[3 * i + 2]), sc.light,
                        sc, p, shadowFactors ? shadowFactors[i] : -1.0f);
        outRgba[4 * i] = c.r; outRgba[4 * i + 1] = c.g; outRgba[4 * i + 2] = c.b; outRgba[4 * i + 3] = c.a;
    }
    return 0;
}

int mcref_in_shadow(void* s, const float* pts, const float* nrms, const float* lights, int n, int* out) {
    const Scene& sc = static_cast<RefScene*>(s)->scene;
    for (int i = 0; i < n; ++i)
        out[i] = isInShadow(Vec3(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]),
                            Vec3(nrms[3 * i], nrms[3 * i + 1], nrms[3 * i + 2]),
                            Vec3(lights[3 * i], lights[3 * i + 1], lights[3 * i + 2]), sc) ? 1 : 0;
    return 0;
}

int mcref_soft_shadow(void* s, const float* pts, const float* nrms, const uint32_t* seeds, int samples, int n,
                      float* out) {
    const Scene& sc = static_cast<RefScene*>(s)->scene;
    for (int i = 0; i < n; ++i)
        out[i] = computeSoftShadow(Vec3(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]),
                                   Vec3(nrms[3 * i], nrms[3 * i + 1], nrms[3 * i + 2]), sc.light, sc, samples,
                                   seeds[i]);
    return 0;
}

int mcref_ambient_occlusion(void* s, const float* pts, const float* nrms, const uint32_t* seeds, int samples,
                            float radius, int n, float* out) {
    const Scene& sc = static_cast<RefScene*>(s)->scene;
    for (int i = 0; i < n; ++i)
        out[i] = RayTracer::computeAO(Vec3(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]),
                                      Vec3(nrms[3 * i], nrms[3 * i + 1], nrms[3 * i + 2]), sc, samples, radius,
                                      seeds[i]);
    return 0;
}

int mcref_generate_rays(void* s, float aspect, const float* uv, int n, McRay* out) {
    const Scene& sc = static_cast<RefScene*>(s)->scene;
    for (int i = 0; i < n; ++i) {
        Ray r = sc.camera.generateRay(uv[2 * i], uv[2 * i + 1], aspect);
        out[i].origin[0] = r.origin.x; out[i].origin[1] = r.origin.y; out[i].origin[2] = r.origin.z;
        out[i].dir[0] = r.direction.x; out[i].dir[1] = r.direction.y; out[i].dir[2] = r.direction.z;
    }
    return 0;
}

int mcref_background(void* s, const McConfig* cfg, int useConfig, const float* uv, int n, float* outRgba) {
    const Scene& sc = static_cast<RefScene*>(s)->scene;
    RayTracer::Config k = configFrom(*cfg);
    for (int i = 0; i < n; ++i) {
        Color c = RayTracer::backgroundColor(sc, uv[2 * i], uv[2 * i + 1], useConfig ? &k : nullptr);
        outRgba[4 * i] = c.r; outRgba[4 * i + 1] = c.g; outRgba[4 * i + 2] = c.b; outRgba[4 * i + 3] = c.a;
    }
    return 0;
}

// Hit mask + triangle id of the pinhole ray through every pixel centre
// (SURVEY.md §8c): tri id = box*12 + face*2, -1 on miss.
int mcref_aov(void* s, const McConfig* cfg, int32_t* outTriId) {
    const Scene& sc = static_cast<RefScene*>(s)->scene;
    const float aspect = static_cast<float>(cfg->width) / static_cast<float>(cfg->height);
    for (int py = 0; py < cfg->height; ++py)
        for (int px = 0; px < cfg->width; ++px) {
            float u = (static_cast<float>(px) + 0.5f) / static_cast<float>(cfg->width);
            float v = (static_cast<float>(py) + 0.5f) / static_cast<float>(cfg->height);
            Ray r = sc.camera.generateRay(u, v, aspect);
            HitResult h;
            int b, f;
            classifyHit(sc, r, -1, h, b, f);
            outTriId[py * cfg->width + px] = h.hit ? b * 12 + f * 2 : -1;
        }
    return 0;
}

// ImageWriter::writePNG (image_writer.cpp:6-28); returns 1 on success like the reference's bool.
int mcref_write_png(const float* rgba, int w, int h, const char* path) {
    Image img(w, h);
    if (w > 0 && h > 0) std::memcpy(img.pixels.data(), rgba, img.pixels.size() * sizeof(Color));
    return ImageWriter::writePNG(img, path ? path : "") ? 1 : 0;
}

// The libstdc++ generator and distribution the reference draws from (tile_renderer.cpp:78-79,
// shading.cpp:43-44): raw engine words and uniform_real_distribution<float>(0,1) values of one seed.
void mcref_mt19937(uint32_t seed, int n, uint32_t* outU32, float* outCanonical) {
    std::mt19937 a(seed), b(seed);
    std::uniform_real_distribution<float> dist(0.0f, 1.0f);
    for (int i = 0; i < n; ++i) {
        if (outU32) outU32[i] = static_cast<uint32_t>(a());
        if (outCanonical) outCanonical[i] = dist(b);
    }
}

// static_cast<unsigned int>(float) as the reference's compiler lowers it (raytracer.cpp:110-112)
uint32_t mcref_seed_cast(float f) {
    volatile float v = f;
    return static_cast<unsigned int>(v);
}

int mcref_hardware_threads(void) { return static_cast<int>(std::thread::hardware_concurrency()); }

}  // extern "C"
