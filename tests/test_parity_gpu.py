"""GPU parity: the CUDA path, called through the C ABI, against the CPU oracle.

Bars (BASELINE.json north_star): hit mask and (box, face) id bit-exact; geometry
(t, point, normal, texel) bit-exact; 8-bit RGBA within +-1 LSB on >= 99.9 % of pixels;
background pixels bit-exact as floats.
"""
import numpy as np
import pytest

from minecraftskin_raytracer_b200 import _abi
from tests.conftest import pixel_report
from tests.scenes import RENDER_CASES, make_config, random_rays, synth_skin

pytestmark = pytest.mark.gpu


def _scene(mclib, seed=1, kind="64x64", pose=None):
    return mclib.build_skin_scene(synth_skin(seed, kind), pose)


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.mark.parametrize("pose", [None, "walking", "dab"])
def test_intersect_bit_exact(gpu, oracle, pose):
    scene = _scene(gpu, 1, "64x64", pose)
    rays = random_rays(np.random.default_rng(7), 20000)
    want = oracle.intersect(scene, rays)
    got = gpu.intersect(scene, rays)
    assert want["hit"].mean() > 0.2
    assert np.array_equal(got["hit"], want["hit"])
    for f in ("box", "face", "is_outer_layer"):
        assert np.array_equal(got[f], want[f]), f
    for f in ("t", "point", "normal", "tex_color"):
        assert np.array_equal(_bits(got[f]), _bits(want[f])), f


def test_intersect_single_box(gpu, oracle):
    scene = _scene(gpu, 3, "64x64", "waving")
    rays = random_rays(np.random.default_rng(11), 4000)
    for box in (0, 1, 4, 5):
        want = oracle.intersect(scene, rays, box=box)
        got = gpu.intersect(scene, rays, box=box)
        assert np.array_equal(got["hit"], want["hit"])
        assert np.array_equal(_bits(got["t"]), _bits(want["t"]))
        assert np.array_equal(got["face"], want["face"])


def test_generate_rays_and_background_bit_exact(gpu, oracle):
    scene = _scene(gpu)
    uv = np.random.default_rng(3).random((5000, 2)).astype(np.float32)
    for aspect in (1.0, 16.0 / 9.0, 0.5):
        want = oracle.generate_rays(scene, aspect, uv)
        got = gpu.generate_rays(scene, aspect, uv)
        assert np.array_equal(_bits(got["dir"]), _bits(want["dir"]))
        assert np.array_equal(_bits(got["origin"]), _bits(want["origin"]))
    for kw in (dict(), dict(gradient_bg=0), dict(gradient_scale=2.5)):
        cfg = make_config(**kw)
        assert np.array_equal(_bits(gpu.background(scene, cfg, uv)), _bits(oracle.background(scene, cfg, uv)))
    cfg = make_config()
    assert np.array_equal(_bits(gpu.background(scene, cfg, uv, use_config=False)),
                          _bits(oracle.background(scene, cfg, uv, use_config=False)))


def _surface_points(oracle, scene, n, seed):
    rays = random_rays(np.random.default_rng(seed), n)
    hits = oracle.intersect(scene, rays)
    keep = hits["hit"] == 1
    return rays[keep], hits[keep]


def test_in_shadow_exact(gpu, oracle):
    scene = _scene(gpu, 1, "64x64", "walking")
    rays, hits = _surface_points(oracle, scene, 12000, 5)
    lights = np.tile(np.float32(scene.light_pos), (len(hits), 1))
    lights += np.random.default_rng(1).normal(scale=4.0, size=lights.shape).astype(np.float32)
    want = oracle.in_shadow(scene, hits["point"], hits["normal"], lights)
    got = gpu.in_shadow(scene, hits["point"], hits["normal"], lights)
    assert 0.02 < want.mean() < 0.98
    assert np.array_equal(got, want)


def test_sincos_device_equals_host_model_and_libm(gpu):
    """The device's sinf/cosf == its host model everywhere, == glibc's on an FMA host (bit for bit)."""
    from tests.test_host_side import _host_has_fma, _libm_sincos, sincos_test_angles
    a = sincos_test_angles(400_000)
    got_s, got_c = gpu.sincos(a)
    model_s, model_c = gpu.sincos_model(a)
    assert np.array_equal(_bits(got_s), _bits(model_s)) and np.array_equal(_bits(got_c), _bits(model_c))
    if _host_has_fma():
        sub = a[:: max(1, len(a) // 50_000)]
        want_s, want_c = _libm_sincos(sub)
        s2, c2 = gpu.sincos(sub)
        assert np.array_equal(_bits(s2), _bits(want_s)) and np.array_equal(_bits(c2), _bits(want_c))


def test_powf_device_equals_host_model_and_libm(gpu):
    from tests.test_host_side import _host_has_fma, _libm_powf, powf_test_values
    x, y = powf_test_values(300_000)
    got = gpu.powf(x, y)
    assert np.array_equal(_bits(got), _bits(gpu.powf_model(x, y)))
    if _host_has_fma():
        sub = slice(None, None, max(1, len(x) // 60_000))
        assert np.array_equal(_bits(gpu.powf(x[sub], y[sub])), _bits(_libm_powf(x[sub], y[sub])))


def test_sincos_and_powf_device_equal_golden_libm_vectors(gpu):
    """The device functions against committed glibc outputs (host-independent pin of the libm parity)."""
    from tests.test_host_side import libm_golden
    g = libm_golden()
    sn, cs = gpu.sincos(g["angles"])
    assert np.array_equal(_bits(sn), _bits(g["sin"])) and np.array_equal(_bits(cs), _bits(g["cos"]))
    assert np.array_equal(_bits(gpu.powf(g["pow_x"], g["pow_y"])), _bits(g["pow"]))


# With sinf/cosf bit-identical to the host's (FMA build of glibc) the sampled light points are the
# reference's own, so shadow factors match exactly; on a host whose glibc picks the non-FMA
# variant a last-ulp difference can move a shadow ray across an edge now and then.
def _flip_budget():
    from tests.test_host_side import _host_has_fma
    return 0.0 if _host_has_fma() else 2e-3


@pytest.mark.parametrize("samples", [2, 8, 64, 120])
def test_soft_shadow(gpu, oracle, samples):
    scene = _scene(gpu, 1, "64x64", None)
    rays, hits = _surface_points(oracle, scene, 3000 if samples < 100 else 600, 9)
    seeds = np.random.default_rng(2).integers(0, 2**32, size=len(hits), dtype=np.uint32)
    want = oracle.soft_shadow(scene, hits["point"], hits["normal"], seeds, samples)
    got = gpu.soft_shadow(scene, hits["point"], hits["normal"], seeds, samples)
    differ = got != want
    assert differ.mean() <= _flip_budget(), differ.mean()
    assert np.abs(got - want).max() <= 1.0 / samples + 1e-6


def test_ambient_occlusion(gpu, oracle):
    scene = _scene(gpu, 4, "64x64", "waving")
    rays, hits = _surface_points(oracle, scene, 3000, 13)
    seeds = np.random.default_rng(4).integers(0, 2**32, size=len(hits), dtype=np.uint32)
    for samples in (8, 16):
        want = oracle.ambient_occlusion(scene, hits["point"], hits["normal"], seeds, samples, 3.0)
        got = gpu.ambient_occlusion(scene, hits["point"], hits["normal"], seeds, samples, 3.0)
        differ = got != want
        assert differ.mean() <= _flip_budget(), differ.mean()


def test_shade_matches(gpu, oracle):
    scene = _scene(gpu, 1, "64x64", None)
    rays, hits = _surface_points(oracle, scene, 6000, 17)
    view = -rays["dir"]
    cfg = make_config()
    for sf in (None, np.random.default_rng(0).random(len(hits)).astype(np.float32)):
        want = oracle.shade(scene, cfg, hits, view, sf)
        got = gpu.shade(scene, cfg, hits, view, sf)
        # every operation of shade() is IEEE add/mul/div/sqrt except std::pow, which the device
        # evaluates with glibc's own algorithm: identical floats on an FMA host
        if _flip_budget() == 0.0:
            assert np.array_equal(_bits(got), _bits(want))
        assert np.abs(got - want).max() <= 2e-6


@pytest.mark.parametrize("use_config,depth", [(True, 0), (False, 0), (True, 2), (True, 9)])
def test_trace_matches(gpu, oracle, use_config, depth):
    scene = _scene(gpu, 5, "64x64", "fighting")
    rays = random_rays(np.random.default_rng(23), 3000)
    cfg = make_config(max_bounces=4)
    want = oracle.trace(scene, cfg, rays, depth=depth, use_config=use_config)
    got = gpu.trace(scene, cfg, rays, depth=depth, use_config=use_config)
    if _flip_budget() == 0.0:
        assert np.array_equal(_bits(got), _bits(want))   # whole paths, bounces and soft shadows included
    close = np.abs(got - want).max(axis=1) <= 1e-5
    assert close.mean() >= 0.998, close.mean()


# default: classified primary pass + wavefront shading.  The other modes run the alternative
# code paths the library keeps (and falls back to) and must give the same image.
RENDER_MODES = {
    "default": {},
    "all_active": {"force_all_active": 1},
    "megakernel": {"shade_mode": 1},
    "megakernel_warp": {"shade_mode": 2},
    "tiny_wave_budget": {"wave_budget_bytes": 1 << 20},   # most pixels overflow to the megakernel
    "small_queue": {"wave_queue_pct": 20},                # part of the paths overflow the hit queue: redone in-thread
    "tiny_queue": {"wave_queue_pct": 1},                  # nearly all of them do
    "split_tiles": {"primary_blocks_per_sm": 100000},     # every tile split over one block per 256-pixel round
    "split_heavy_tiles": {"heavy_tiles_per_sm": 100000},  # the figure's tiles split over one block per round, the rest whole
    "one_lane_no_graph": {"frame_lanes": 1, "use_graphs": 0, "cache_tile_seeds": 0},
    "five_lanes": {"frame_lanes": 5},
    "two_soft_blocks": {"soft_blocks_per_sm": 1, "shade_blocks_per_sm": 2, "frame_lanes": 2},
}


@pytest.mark.parametrize("case", RENDER_CASES, ids=[c[0] for c in RENDER_CASES])
@pytest.mark.parametrize("mode", list(RENDER_MODES), ids=list(RENDER_MODES))
def test_render_parity(gpu, oracle, case, mode):
    name, seed, kind, pose, over = case
    scene = _scene(gpu, seed, kind, pose)
    cfg = make_config(**over)
    want = oracle.render(scene, cfg)
    ctx_opts = RENDER_MODES[mode]
    got = _render_with_options(gpu, scene, cfg, ctx_opts)
    rep = pixel_report(got, want, oracle.quantize)
    # hit mask / triangle id at pixel centres
    tri_want = oracle.aov(scene, cfg)
    tri_got = gpu.aov(scene, cfg)
    assert np.array_equal(tri_got, tri_want)
    assert rep["within1"] >= 0.999, rep
    if _flip_budget() == 0.0:
        # sinf, cosf and powf are evaluated with the host libm's own algorithms and everything else is
        # IEEE arithmetic in the reference's order: the float image is the reference's, bit for bit
        assert np.array_equal(_bits(got), _bits(want)), rep
    # pixels none of whose samples can hit anything: exact float equality
    bg = _background_pixels(oracle, scene, cfg, want)
    assert np.array_equal(_bits(got[bg]), _bits(want[bg]))
    assert bg.mean() > 0.3


def _background_pixels(oracle, scene, cfg, want):
    """Pixels whose colour equals the pure-background render (an empty scene) bit for bit."""
    from minecraftskin_raytracer_b200.scene import FlatScene
    empty = FlatScene(light_pos=scene.light_pos, cam_pos=scene.cam_pos, cam_target=scene.cam_target,
                      cam_up=scene.cam_up, cam_fov_deg=scene.cam_fov_deg, background=scene.background)
    pure = oracle.render(empty, cfg)
    return (pure.view(np.uint32) == want.view(np.uint32)).all(axis=-1)


def _render_with_options(gpu, scene, cfg, opts):
    """Render through the device-resident context (so options can be set) and fetch with torch-free cudaMemcpy."""
    import ctypes as C
    if not opts:
        f32, _, _ = gpu.render(scene, cfg)
        return f32
    import torch
    ctx = gpu.Context(0)
    try:
        for k, v in opts.items():
            ctx.set_option(k, v)
        ctx.set_scene(scene, cfg)
        out = torch.zeros((cfg.height, cfg.width, 4), dtype=torch.float32, device="cuda:0")
        torch.cuda.synchronize()
        frames = []
        for _ in range(3):  # direct launches, then the call that captures the graph, then a replay
            out.zero_()
            torch.cuda.synchronize()
            ctx.render_bands(0, 1, out.data_ptr(), 0, 0)
            ctx.sync()
            frames.append(out.cpu().numpy())
        assert np.array_equal(_bits(frames[0]), _bits(frames[1])) and np.array_equal(_bits(frames[0]), _bits(frames[2]))
        return frames[2]
    finally:
        ctx.close()


def test_u8_output_and_progress(gpu, oracle):
    scene = _scene(gpu, 1)
    cfg = make_config(width=100, height=70, samples_per_pixel=2, max_bounces=2)
    calls = []
    f32, u8, stats = gpu.render(scene, cfg, want_u8=True, progress=lambda d, t: calls.append((d, t)))
    total = len(gpu.generate_tiles(100, 70, 32))
    assert calls == [(i, total) for i in range(1, total + 1)]  # tests/test_tile_renderer.cpp:85-104
    assert np.array_equal(u8, oracle.quantize(f32))             # image_writer.cpp:18-22
    assert stats["n_kernel_launches"] >= 2 and stats["n_tiles"] == total


def test_render_into_page_locked_buffers_overlaps_copy_out(gpu, oracle):
    """With page-locked destinations the image leaves for the host while the shading pass runs
    (whole frame after the primary pass, the figure's rectangle again at the end): same bytes."""
    import torch
    scene = _scene(gpu, 6, "64x64", "waving")
    for over in (dict(width=320, height=200, samples_per_pixel=4, max_bounces=3),
                 dict(width=96, height=128, samples_per_pixel=1, max_bounces=1)):
        cfg = make_config(**over)
        want, want_u8, _ = gpu.render(scene, cfg, want_u8=True)            # pageable destinations: plain copy
        f32 = torch.empty((cfg.height, cfg.width, 4), dtype=torch.float32, pin_memory=True).numpy()
        u8 = torch.empty((cfg.height, cfg.width, 4), dtype=torch.uint8, pin_memory=True).numpy()
        for _ in range(3):  # direct launches, graph capture, graph replay
            f32.fill(-1.0)
            u8.fill(7)
            gpu.render(scene, cfg, out_f32=f32, out_u8=u8)
            assert np.array_equal(_bits(f32), _bits(want)) and np.array_equal(u8, want_u8)
    assert pixel_report(want, oracle.render(scene, cfg), oracle.quantize)["within1"] >= 0.999


def test_render_tile_matches_full_frame(gpu, oracle):
    scene = _scene(gpu, 2, "64x64", "walking")
    cfg = make_config(width=80, height=72, samples_per_pixel=3, max_bounces=2)
    full, _, _ = gpu.render(scene, cfg)
    img = np.zeros((72, 80, 4), dtype=np.float32)
    img[..., 3] = 1
    for t in gpu.generate_tiles(80, 72, 32):
        img = gpu.render_tile(scene, cfg, tuple(t), img)
    assert np.array_equal(_bits(img), _bits(full))


def test_deterministic_run_to_run(gpu):
    scene = _scene(gpu, 1, "64x64", "dab")
    cfg = make_config(width=96, height=96, samples_per_pixel=4)
    a, _, _ = gpu.render(scene, cfg)
    b, _, _ = gpu.render(scene, cfg)
    assert np.array_equal(_bits(a), _bits(b))  # tests/test_tile_renderer_props.cpp:89-134


def test_graph_replay_and_seed_cache_survive_interleaved_frames(gpu, oracle):
    """A context replays a captured frame graph and keeps seeded tile engines between frames; other
    frame sizes, scenes and partitions rendered in between must not leak into a later replay."""
    import torch
    scene_a, scene_b = _scene(gpu, 1, "64x64", "walking"), _scene(gpu, 7, "legacy", None)
    cfg_a = make_config(width=160, height=192, samples_per_pixel=4, max_bounces=3)
    cfg_b = make_config(width=160, height=128, samples_per_pixel=16, max_bounces=1)
    want_a, want_b = oracle.render(scene_a, cfg_a), oracle.render(scene_b, cfg_a)
    ctx = gpu.Context(0)
    try:
        out = torch.zeros((192, 160, 4), dtype=torch.float32, device="cuda:0")

        def frame(scene, cfg, first=0, stride=1):
            ctx.set_scene(scene, cfg)
            out.zero_()
            torch.cuda.synchronize()
            ctx.render_bands(first, stride, out.data_ptr(), 0, 0)
            ctx.sync()
            return out.cpu().numpy()[: cfg.height]

        first_a = frame(scene_a, cfg_a)
        assert pixel_report(first_a, want_a, oracle.quantize)["within1"] >= 0.999
        for _ in range(3):
            assert np.array_equal(_bits(frame(scene_a, cfg_a)), _bits(first_a))      # capture, replay, replay
        frame(scene_b, cfg_b)                                                        # other size and spp: reseeds the tiles
        frame(scene_a, cfg_a, 1, 2)                                                  # other partition
        assert np.array_equal(_bits(frame(scene_a, cfg_a)), _bits(first_a))
        got_b = frame(scene_b, cfg_a)                                                # same frame shape, other scene
        assert pixel_report(got_b, want_b, oracle.quantize)["within1"] >= 0.999
        for _ in range(2):
            assert np.array_equal(_bits(frame(scene_b, cfg_a)), _bits(got_b))
        assert np.array_equal(_bits(frame(scene_a, cfg_a)), _bits(first_a))
    finally:
        ctx.close()


def test_empty_scene_and_degenerate_sizes(gpu, oracle):
    from minecraftskin_raytracer_b200.scene import FlatScene
    empty = FlatScene()
    cfg = make_config(width=64, height=48, samples_per_pixel=2)
    got, _, _ = gpu.render(empty, cfg)
    assert np.array_equal(_bits(got), _bits(oracle.render(empty, cfg)))
    for w, h, ts in ((0, 10, 32), (10, 0, 32), (10, 10, 0), (-5, 10, 32)):
        cfg = make_config(width=w, height=h, tile_size=ts)
        f32, _, stats = gpu.render(empty, cfg)   # tile_renderer.cpp:144-146: empty image, no error
        assert f32.size == max(w, 0) * max(h, 0) * 4


# ---------------------------------------------------------------- against the committed reference outputs
def test_cuda_vs_golden_reference_renders(gpu, oracle):
    """CUDA frames vs frames rendered by the UNMODIFIED reference (tests/golden/reference_vectors.npz)."""
    from tests.golden_data import golden_render_cases, golden_scene, vectors
    v = vectors()
    for name, cfg in golden_render_cases():
        scene = golden_scene(name)
        want = v[f"{name}/image"]
        got, _, _ = gpu.render(scene, cfg)
        rep = pixel_report(got, want, oracle.quantize)
        assert rep["within1"] >= 0.999, (name, rep)
        if _flip_budget() == 0.0:   # the golden frames come from the reference on an FMA host (same glibc)
            assert np.array_equal(_bits(got), _bits(want)), (name, rep)
        assert np.array_equal(gpu.aov(scene, cfg), v[f"{name}/tri_id"]), name


def test_cuda_vs_golden_reference_rays(gpu):
    from tests.golden_data import ray_scene, vectors
    v = vectors()
    scene = ray_scene(gpu.build_skin_scene)
    rays, want = v["rays/rays"], v["rays/hits"]
    got = gpu.intersect(scene, rays)
    for f in ("hit", "box", "face", "is_outer_layer"):
        assert np.array_equal(got[f], want[f]), f
    for f in ("t", "point", "normal", "tex_color"):
        assert np.array_equal(_bits(got[f]), _bits(want[f])), f
    keep = want["hit"] == 1
    cfg = make_config(max_bounces=3)
    assert np.abs(gpu.shade(scene, cfg, want[keep], -rays["dir"][keep], v["rays/shade_sf"]) - v["rays/shade_soft"]).max() <= 2e-6
    assert np.abs(gpu.shade(scene, cfg, want[keep], -rays["dir"][keep], None) - v["rays/shade_hard"]).max() <= 2e-6
    for key, use_cfg in (("rays/trace_cfg", True), ("rays/trace_nocfg", False)):
        close = np.abs(gpu.trace(scene, cfg, rays[:1500], 0, use_cfg) - v[key]).max(axis=1) <= 1e-5
        assert close.mean() >= 0.998, key
    assert (gpu.soft_shadow(scene, want["point"][keep], want["normal"][keep], v["rays/seeds"], 8) != v["rays/soft8"]).mean() <= _flip_budget()
    cam = gpu.generate_rays(scene, 16.0 / 9.0, v["rays/uv"])
    assert np.array_equal(_bits(cam["dir"]), _bits(v["rays/camera"]["dir"]))
    assert np.array_equal(_bits(gpu.background(scene, cfg, v["rays/uv"])), _bits(v["rays/background"]))


# ---------------------------------------------------------------- BASELINE.json configurations at full size
@pytest.mark.parametrize("name,seed,kind,over", [
    ("C1", 1, "64x64", dict(width=512, height=512, samples_per_pixel=1, max_bounces=2)),
    ("C2", 2, "legacy", dict(width=1920, height=1080, samples_per_pixel=4, max_bounces=4)),
    ("headline", 0, "64x64", dict(width=1920, height=1080, samples_per_pixel=16, max_bounces=4)),
], ids=["C1", "C2", "headline"])
def test_baseline_configs_at_full_size(gpu, oracle, name, seed, kind, over):
    """BASELINE.json configs[0], configs[1] and the headline frame, whole, against the CPU path run here on
    the box's host cores (the unmodified reference when its library travelled, else the C restatement):
    the north star's bar (>= 99.9 % of pixels within 1 LSB, hit ids exact) — and, on an FMA host, the same bits."""
    from oracle.harness import Reference
    scene = _scene(gpu, seed, kind, None)
    cfg = make_config(**over)
    ref = Reference.load()
    want = ref.render(scene, cfg) if ref is not None else oracle.render(scene, cfg)
    got, _, _ = gpu.render(scene, cfg)
    rep = pixel_report(got, want, oracle.quantize)
    assert rep["within1"] >= 0.999, (name, rep)
    if _flip_budget() == 0.0:
        assert np.array_equal(_bits(got), _bits(want)), (name, rep)
    assert np.array_equal(gpu.aov(scene, cfg), oracle.aov(scene, cfg)), name


# ---------------------------------------------------------------- full-size properties (no oracle run)
def test_headline_frame_properties(gpu, oracle):
    """BASELINE headline size (1080p / 16 spp / 4 bounces): properties that need no CPU frame."""
    import json
    from pathlib import Path
    counts = json.loads((Path(__file__).parent / "golden" / "work_counts.json").read_text())["headline_1080p_16spp_4b"]
    scene = _scene(gpu, 0)
    cfg = make_config(**counts["config"])
    f32, u8, stats = gpu.render(scene, cfg, want_u8=True)
    assert np.array_equal(u8, oracle.quantize(f32))
    # background pixels = the empty-scene frame, bit for bit; everything else stays inside the figure's box
    empty, _, _ = gpu.render(type(scene)(), cfg)
    same = (f32.view(np.uint32) == empty.view(np.uint32)).all(axis=-1)
    ys, xs = np.nonzero(~same)
    assert 0.05 < (~same).mean() < 0.10
    assert xs.min() > 700 and xs.max() < 1220 and ys.min() > 100 and ys.max() < 980
    # pixels the classification pass found active == pixels where any of the 16 samples hits;
    # the oracle counted the samples, so active pixels are bounded by it from both sides
    hit_samples = counts["counters"]["n_primary_rays"] - counts["counters"]["n_background_primary"]
    assert hit_samples / 16 <= stats["n_active_pixels"] <= hit_samples
    assert np.all(np.isfinite(f32)) and f32.min() >= 0.0 and f32.max() <= 1.0
    # tile-row partitions reassemble to the same frame (what the multi-GPU path relies on)
    import torch
    from minecraftskin_raytracer_b200 import bands
    ctx = gpu.Context(0)
    try:
        ctx.set_scene(scene, cfg)
        world = 3
        parts = []
        for r in range(world):
            band = torch.zeros((bands.padded_band_rows(cfg.height, cfg.tile_size, world), cfg.width, 4), device="cuda:0")
            ctx.render_bands(r, world, band.data_ptr(), 0, 0)
            ctx.sync()
            parts.append(band)
        frame = bands.deinterleave(parts, torch.zeros((cfg.height, cfg.width, 4), device="cuda:0"), cfg.tile_size)
        assert np.array_equal(frame.cpu().numpy().view(np.uint32), f32.view(np.uint32))
    finally:
        ctx.close()


@pytest.mark.parametrize("mode", ["grouped", "grouped_small_groups", "frame_by_frame"])
def test_batch_of_skins(gpu, oracle, mode):
    """mcskin_cuda_context_render_batch (BASELINE config 4 in miniature): many skins, one config.
    grouped: scenes with equal frame descriptions share launches (gridDim.y = scene);
    frame_by_frame: one frame at a time over the context's lanes."""
    import torch
    n = 24
    cfg = make_config(width=64, height=64, samples_per_pixel=4, max_bounces=2)
    scenes = [_scene(gpu, 100 + i, "legacy" if i % 5 == 0 else "64x64", [None, "walking", "dab"][i % 3]) for i in range(n)]
    ctx = gpu.Context(0)
    try:
        if mode == "frame_by_frame":
            ctx.set_option("batch_mode", 0)
        if mode == "grouped_small_groups":
            ctx.set_option("batch_group", 7)   # chunks of 7 scenes: 24 = 7 + 7 + 7 + 3, several groups per chunk
        out = torch.zeros((n, 64, 64, 4), dtype=torch.float32, device="cuda:0")
        out_u8 = torch.zeros((n, 64, 64, 4), dtype=torch.uint8, device="cuda:0")
        torch.cuda.synchronize()
        for _ in range(2):  # twice: buffers and staging are reused
            ctx.render_batch(scenes, cfg, out.data_ptr(), out_u8.data_ptr(), 0)
            ctx.sync()
        got = out.cpu().numpy()
        for i in (0, 1, 5, 6, 7, 11, 20, 23):
            single, _, _ = gpu.render(scenes[i], cfg)
            assert np.array_equal(_bits(got[i]), _bits(single)), i          # batched == one at a time
        for i in (0, 5, 23):
            assert pixel_report(got[i], oracle.render(scenes[i], cfg), oracle.quantize)["within1"] >= 0.999
        assert np.array_equal(out_u8.cpu().numpy(), oracle.quantize(got))
    finally:
        ctx.close()


def test_batch_headline_shape(gpu, oracle):
    """BASELINE config 4's frame shape (256x256, 4 spp, 2 bounces) for a few dozen skins in one grouped batch."""
    import torch
    n = 40
    cfg = make_config(width=256, height=256, samples_per_pixel=4, max_bounces=2)
    scenes = [_scene(gpu, i) for i in range(n)]
    ctx = gpu.Context(0)
    try:
        out = torch.zeros((n, 256, 256, 4), dtype=torch.float32, device="cuda:0")
        torch.cuda.synchronize()
        ctx.render_batch(scenes, cfg, out.data_ptr(), 0, 0)
        ctx.sync()
        got = out.cpu().numpy()
        for i in (0, 17, 39):
            single, _, _ = gpu.render(scenes[i], cfg)
            assert np.array_equal(_bits(got[i]), _bits(single)), i
        assert pixel_report(got[3], oracle.render(scenes[3], cfg), oracle.quantize)["within1"] >= 0.999
    finally:
        ctx.close()


def test_rows_into_frame_partitions(gpu, oracle):
    """render_rows_into_frame: tile-row partitions written at their own place in one full frame (what the
    multi-GPU path does over peer memory) reassemble the single-device frame; the buffer is a plain
    allocation of the C ABI viewed by torch through __cuda_array_interface__."""
    import torch
    scene = _scene(gpu, 8, "64x64", "running")
    cfg = make_config(width=200, height=330, samples_per_pixel=4, max_bounces=3)
    want, want_u8, _ = gpu.render(scene, cfg, want_u8=True)
    buf = gpu.DeviceBuffer(0, (cfg.height, cfg.width, 4))
    frame = torch.as_tensor(buf, device="cuda:0")
    frame_u8 = torch.zeros((cfg.height, cfg.width, 4), dtype=torch.uint8, device="cuda:0")
    ctx = gpu.Context(0)
    try:
        ctx.set_scene(scene, cfg)
        for world in (1, 2, 5):
            for rep in range(3):  # direct, capture, replay
                frame.fill_(-1.0)
                frame_u8.zero_()
                torch.cuda.synchronize()
                for r in range(world):
                    ctx.render_rows_into_frame(r, world, buf.ptr, frame_u8.data_ptr(), 0)
                    ctx.sync()
                assert np.array_equal(_bits(frame.cpu().numpy()), _bits(want)), (world, rep)
                assert np.array_equal(frame_u8.cpu().numpy(), want_u8), (world, rep)
        assert len(buf.ipc_handle()) == 64
    finally:
        ctx.close()
        del frame
        buf.free()


def test_peer_flags_signal_and_wait(gpu):
    """The barrier of the peer-store exchange: stream-ordered release / acquire of 32-bit flags, with a
    bounded wait (a flag that never arrives sets the timeout word instead of hanging the device)."""
    import torch
    buf = gpu.DeviceBuffer(0, (64,), dtype="uint32")
    words = torch.as_tensor(buf, device="cuda:0")
    try:
        words.zero_()
        torch.cuda.synchronize()
        for epoch in (1, 2, 3):
            for r in (1, 2, 3):
                gpu.peer_signal(0, buf.ptr + 4 * r, epoch)
            gpu.peer_wait(0, buf.ptr + 4, 3, epoch, buf.ptr + 4 * 32)
            torch.cuda.synchronize()
            assert words[1:4].tolist() == [epoch] * 3 and int(words[32]) == 0
        gpu.peer_wait(0, buf.ptr + 4, 3, 2)              # already reached: returns at once
        gpu.peer_wait(0, buf.ptr + 4, 3, 9, buf.ptr + 4 * 32)   # never reached: gives up after ~2 s
        torch.cuda.synchronize()
        assert int(words[32]) == 1
    finally:
        del words
        buf.free()


def test_render_multi_in_process(gpu):
    """mcskin_cuda_render_multi over however many devices this process sees (1 on the test box)."""
    scene = _scene(gpu, 3, "64x64", "running")
    cfg = make_config(width=96, height=80, samples_per_pixel=2, max_bounces=2)
    single, _, _ = gpu.render(scene, cfg)
    n = gpu.device_count()
    multi, multi_u8, stats = gpu.render(scene, cfg, want_u8=True, multi_devices=n)
    assert np.array_equal(_bits(multi), _bits(single))
    assert stats["n_tiles"] == len(gpu.generate_tiles(96, 80, 32))
