"""mcskin-b200: B200-native render hot path of MCSkin RaytraceRenderer (TileRenderer::render)."""
from . import _abi  # noqa: F401
from .scene import BUILTIN_POSES, FlatScene, synth_skin  # noqa: F401

__all__ = ["FlatScene", "synth_skin", "BUILTIN_POSES"]
