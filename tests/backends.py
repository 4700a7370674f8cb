"""Adapters giving the CUDA library the same call surface as the oracle harness."""
from __future__ import annotations


class CudaBackend:
    name = "cuda"

    def __init__(self, lib):
        self.lib = lib

    def intersect(self, scene, rays, box=-1):
        return self.lib.intersect(scene, rays, box=box)

    def shade(self, scene, cfg, hits, view_dirs, shadow_factors=None):
        return self.lib.shade(scene, cfg, hits, view_dirs, shadow_factors)

    def in_shadow(self, scene, p, n, l):
        return self.lib.in_shadow(scene, p, n, l)

    def trace(self, scene, cfg, rays, depth=0, use_config=True):
        return self.lib.trace(scene, cfg, rays, depth=depth, use_config=use_config)

    def generate_rays(self, scene, aspect, uv):
        return self.lib.generate_rays(scene, aspect, uv)

    def background(self, scene, cfg, uv, use_config=True):
        return self.lib.background(scene, cfg, uv, use_config=use_config)

    def render(self, scene, cfg):
        return self.lib.render(scene, cfg)[0]

    def generate_tiles(self, w, h, ts):
        return self.lib.generate_tiles(w, h, ts)
