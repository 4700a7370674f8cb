"""Generates tests/golden/work_counts.json: exact event counts of the reference algorithm
(oracle/mcskin_oracle.c counters, cross-checked against the unmodified reference's
intersectScene call count when oracle/_ref is present) for the benchmark workloads.

bench.py turns these fixed integers into "unique rays" (Mrays/s) and algorithmic
lane-ops (roofline); they depend only on scene + config, not on the machine.

    python tests/golden/make_work_counts.py            # ~1-2 min of CPU
"""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from minecraftskin_raytracer_b200 import _abi, lib  # noqa: E402
from minecraftskin_raytracer_b200.scene import synth_skin  # noqa: E402
from oracle.harness import Oracle, Reference  # noqa: E402

# name -> (skin seed, kind, pose, config overrides)      (SURVEY.md §8d "Configs restated")
WORKLOADS = {
    "headline_1080p_16spp_4b": (0, "64x64", None, dict(width=1920, height=1080, samples_per_pixel=16, max_bounces=4)),
    "headline_1080p_4spp_4b": (0, "64x64", None, dict(width=1920, height=1080, samples_per_pixel=4, max_bounces=4)),
    "headline_1080p_2spp_4b": (0, "64x64", None, dict(width=1920, height=1080, samples_per_pixel=2, max_bounces=4)),
    "c1_512_1spp_2b": (1, "64x64", None, dict(width=512, height=512, samples_per_pixel=1, max_bounces=2)),
    "c2_legacy_1080p_4spp_4b": (2, "legacy", None, dict(width=1920, height=1080, samples_per_pixel=4, max_bounces=4)),
    "headline_walking_1080p_4spp_4b": (0, "64x64", "walking", dict(width=1920, height=1080, samples_per_pixel=4, max_bounces=4)),
    "c4_item_256_4spp_2b": (0, "64x64", None, dict(width=256, height=256, samples_per_pixel=4, max_bounces=2)),
}


def main():
    orc = Oracle()
    ref = Reference.load()
    out = {}
    for name, (seed, kind, pose, over) in WORKLOADS.items():
        scene = lib.build_skin_scene(synth_skin(seed, kind), pose)
        cfg = _abi.default_config(**over)
        t = time.time()
        _, cnt = orc.render(scene, cfg, counters=True)
        dt = time.time() - t
        entry = dict(skin_seed=seed, skin_kind=kind, pose=pose, config=over, counters=cnt,
                     n_boxes_plain=int((scene.boxes["has_rotation"] == 0).sum()),
                     n_boxes_rotated=int((scene.boxes["has_rotation"] != 0).sum()),
                     unique_rays=cnt["n_intersect_scene"] - cnt["n_retests"])
        if ref is not None and over["width"] * over["height"] * over["samples_per_pixel"] <= 1920 * 1080 * 4:
            _, calls = ref.render(scene, cfg, counters=True)
            assert calls == cnt["n_intersect_scene"], (name, calls, cnt["n_intersect_scene"])
            entry["reference_intersect_scene_calls"] = calls
        out[name] = entry
        print(f"{name}: {dt:.1f}s unique_rays={entry['unique_rays']}", flush=True)
    path = Path(__file__).with_name("work_counts.json")
    path.write_text(json.dumps(out, indent=1, sort_keys=True) + "\n")
    print("wrote", path)


if __name__ == "__main__":
    main()
