// dropin_test.cpp — the reference's TileRenderer tests (tests/test_tile_renderer.cpp,
// tests/test_intersection.cpp, tests/test_raytracer.cpp) against the drop-in C++ API.
// Plain asserts (gtest is not available offline).  Exit code 0 = all passed.
// Usage: dropin_test <out.f32>   (also dumps a 80x64 2-spp frame for the Python side to compare)
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "mcskin/detail/unflatten.hpp"
#include "mcskin_cuda.h"
#include "raytracer/intersection.h"
#include "raytracer/raytracer.h"
#include "raytracer/tile_renderer.h"
#include "scene/scene.h"

MCSKIN_DEFINE_UNFLATTEN()

#define CHECK(cond)                                                                   \
    do {                                                                              \
        if (!(cond)) {                                                                \
            std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);    \
            std::exit(1);                                                             \
        }                                                                             \
    } while (0)

static Scene simpleScene() {  // tests/test_raytracer.cpp:84-96
    Scene s;
    s.backgroundColor = Color(0.2f, 0.3f, 0.5f, 1.0f);
    s.light.position = Vec3(10, 10, -10);
    s.light.color = Color(1, 1, 1, 1);
    s.camera.position = Vec3(0, 0, -10);
    s.camera.target = Vec3(0, 0, 0);
    s.camera.up = Vec3(0, 1, 0);
    s.camera.fov = 60.0f;
    return s;
}

// one 2x2x2 box whose 12 triangles share ONE external texture (tests/test_raytracer.cpp:98-146)
static Mesh testBox(const TextureRegion* tex) {
    McBox b{};
    for (int k = 0; k < 3; ++k) { b.bounds_min[k] = -1; b.bounds_max[k] = 1; }
    b.n_triangles = 12;
    for (auto& f : b.face) f = McFaceTex{-1, 0, 0};
    McScene f{};
    f.n_boxes = 1;
    f.boxes = &b;
    Scene tmp = unflattenScene(f);
    Mesh m = tmp.meshes[0];
    for (Triangle& t : m.triangles) t.texture = tex;
    return m;
}

int main(int argc, char** argv) {
    // generateTiles (test_tile_renderer.cpp:9-57)
    CHECK(TileRenderer::generateTiles(64, 64, 32).size() == 4);
    auto tiles = TileRenderer::generateTiles(100, 70, 32);
    CHECK(tiles.size() == 12 && tiles[3].x == 96 && tiles[3].width == 4 && tiles[11].height == 6);
    CHECK(TileRenderer::generateTiles(0, 10, 8).empty() && TileRenderer::generateTiles(10, 10, 0).empty());

    Scene scene = simpleScene();
    // ProgressCallbackInvoked (:85-104)
    {
        RayTracer::Config c;
        c.width = 32; c.height = 32; c.maxBounces = 0; c.tileSize = 16; c.threadCount = 1;
        std::atomic<int> calls{0};
        int lastTotal = 0, lastDone = 0;
        Image img = TileRenderer::render(scene, c, [&](int done, int total) { calls++; lastTotal = total; lastDone = done; });
        CHECK(TileRenderer::lastErrors().empty());
        CHECK(calls.load() == 4 && lastTotal == 4 && lastDone == 4);
        CHECK(img.width == 32 && img.height == 32 && img.pixels.size() == 32u * 32u);
    }
    // threadCount is irrelevant; 1 == N bit for bit (:106-145); null callback (:147-159)
    {
        RayTracer::Config c1;
        c1.width = 16; c1.height = 16; c1.maxBounces = 1; c1.tileSize = 8; c1.threadCount = 1;
        RayTracer::Config cN = c1;
        cN.threadCount = 4;
        Image a = TileRenderer::render(scene, c1), b = TileRenderer::render(scene, cN, nullptr);
        CHECK(a.pixels.size() == b.pixels.size());
        for (size_t i = 0; i < a.pixels.size(); ++i) CHECK(a.pixels[i] == b.pixels[i]);
    }
    // zero-tile configs return an empty image without errors (tile_renderer.cpp:144-146)
    {
        RayTracer::Config c;
        c.width = 0; c.height = 16;
        Image img = TileRenderer::render(scene, c);
        CHECK(img.pixels.empty() && TileRenderer::lastErrors().empty());
    }
    // external shared texture + traceRay / intersectScene through the API (test_raytracer.cpp:148-224)
    static TextureRegion tex;
    tex.width = 2; tex.height = 2;
    tex.pixels = {Color(1, 0, 0, 1), Color(0, 1, 0, 1), Color(0, 0, 1, 1), Color(1, 1, 0, 1)};
    Scene boxed = simpleScene();
    boxed.meshes.push_back(testBox(&tex));
    {
        Ray ray(Vec3(0, 0, -10), Vec3(0, 0, 1));
        HitResult h = intersectScene(ray, boxed);
        CHECK(h.hit && std::fabs(h.t - 9.0f) < 1e-4f && std::fabs(h.normal.z + 1.0f) < 1e-4f);
        Color miss = RayTracer::traceRay(Ray(Vec3(0, 0, -10), Vec3(0, 1, 0)), boxed, 0, 3);
        CHECK(miss == boxed.backgroundColor);
        Color deep = RayTracer::traceRay(ray, boxed, 5, 3);
        CHECK(deep == boxed.backgroundColor);
        Color hit = RayTracer::traceRay(ray, boxed, 0, 3);
        CHECK(!(std::fabs(hit.r - 0.2f) < 1e-5f && std::fabs(hit.g - 0.3f) < 1e-5f && std::fabs(hit.b - 0.5f) < 1e-5f));
        Ray centre = boxed.camera.generateRay(0.5f, 0.5f, 1.0f);
        CHECK(std::fabs(centre.direction.z - 1.0f) < 1e-5f && centre.origin == boxed.camera.position);
    }
    // renderTile over generateTiles == render, and a frame dump for the Python comparison
    {
        std::vector<uint8_t> atlas(64 * 64 * 4);
        for (size_t i = 0; i < atlas.size(); ++i) atlas[i] = static_cast<uint8_t>((i * 2654435761u) >> 13);
        for (size_t i = 3; i < atlas.size(); i += 4) atlas[i] = (atlas[i] & 1) ? 255 : 0;
        std::vector<McBox> boxes(MCSKIN_MAX_SKIN_BOXES);
        std::vector<float> texels(static_cast<size_t>(MCSKIN_MAX_SKIN_TEXELS) * 4);
        const float pose[12] = {0, 0, 0, 0, 30, 0, -30, 0, -25, 0, 25, 0};
        McScene flat;
        CHECK(mcskin_build_skin_scene(atlas.data(), 64, 64, pose, boxes.data(), texels.data(), &flat) == MC_OK);
        Scene skin = unflattenScene(flat);
        RayTracer::Config c;
        c.width = 80; c.height = 64; c.samplesPerPixel = 2; c.maxBounces = 2;
        Image full = TileRenderer::render(skin, c);
        CHECK(TileRenderer::lastErrors().empty());
        Image byTile(c.width, c.height);
        for (const Tile& t : TileRenderer::generateTiles(c.width, c.height, c.tileSize)) TileRenderer::renderTile(t, skin, c, byTile);
        for (size_t i = 0; i < full.pixels.size(); ++i) CHECK(full.pixels[i] == byTile.pixels[i]);
        // MCSKIN_DEVICES spreads the same render(scene, settings) call over the GPUs of the process
        // (tile_renderer.h:26-28 stays the entry point); same pixels, callback count unchanged
        {
            const int nDev = mcskin_cuda_device_count();
            for (const char* spread : {"all", "2", "3"}) {
                if (std::atoi(spread) > nDev) continue;
                setenv("MCSKIN_DEVICES", spread, 1);
                int calls = 0;
                Image multi = TileRenderer::render(skin, c, [&](int, int) { ++calls; });
                CHECK(TileRenderer::lastErrors().empty());
                CHECK(calls == static_cast<int>(TileRenderer::generateTiles(c.width, c.height, c.tileSize).size()));
                for (size_t i = 0; i < full.pixels.size(); ++i) CHECK(full.pixels[i] == multi.pixels[i]);
            }
            unsetenv("MCSKIN_DEVICES");
            // a frame that fails (no such device) still reports every tile, records the error and
            // returns a default image (tile_renderer.cpp:150-171: progress is reported for tiles that threw, too)
            setenv("MCSKIN_DEVICE", "99", 1);
            int calls = 0, lastDone = 0;
            Image failed = TileRenderer::render(skin, c, [&](int done, int) { ++calls; lastDone = done; });
            unsetenv("MCSKIN_DEVICE");
            const int total = static_cast<int>(TileRenderer::generateTiles(c.width, c.height, c.tileSize).size());
            CHECK(TileRenderer::lastErrors().size() == 1 && TileRenderer::lastErrors()[0].tileIndex == -1);
            CHECK(calls == total && lastDone == total);
            CHECK(failed.width == c.width && failed.pixels.size() == full.pixels.size());
            Image again = TileRenderer::render(skin, c);   // and the next frame is fine again
            CHECK(TileRenderer::lastErrors().empty());
            for (size_t i = 0; i < full.pixels.size(); ++i) CHECK(full.pixels[i] == again.pixels[i]);
        }
        if (argc > 1) {
            FILE* fp = std::fopen(argv[1], "wb");
            CHECK(fp != nullptr);
            std::fwrite(atlas.data(), 1, atlas.size(), fp);
            std::fwrite(full.pixels.data(), sizeof(Color), full.pixels.size(), fp);
            std::fclose(fp);
        }
    }
    std::printf("dropin_test: all checks passed\n");
    return 0;
}
