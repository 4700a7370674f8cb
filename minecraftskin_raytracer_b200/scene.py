"""Host-side scene description for the render hot path.

`FlatScene` is the Python face of the C ABI's McScene (include/mcskin_cuda.h):
the reference's Scene (src/scene/scene.h:10-34) reduced to what the ray tracer
reads — one axis-aligned box per Mesh, six face-texture windows per box, light,
camera, flat background colour.

`synth_skin` is the deterministic synthetic atlas SURVEY.md §8(d) defines for the
benchmark configs; `build_skin_scene` reaches the product's C implementation of
SkinParser::parse + MeshBuilder::buildScene (skin_parser.cpp:11-132,
mesh_builder.cpp:66-202).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _abi

# Built-in pose library (src/scene/pose.h:25-92): (rotX, rotZ) degrees for
# head, body, rightArm, leftArm, rightLeg, leftLeg.
BUILTIN_POSES = {
    "standing": [(0, 0)] * 6,
    "walking": [(0, 0), (0, 0), (30, 0), (-30, 0), (-25, 0), (25, 0)],
    "running": [(-5, 0), (5, 0), (50, 0), (-50, 0), (-45, 0), (45, 0)],
    "waving": [(5, 0), (0, 0), (-140, -20), (0, 0), (0, 0), (0, 0)],
    "sitting": [(0, 0), (0, 0), (-10, 0), (-10, 0), (-90, 0), (-90, 0)],
    "fighting": [(-10, 0), (5, 0), (-90, 10), (20, -10), (-15, 0), (20, 0)],
    "dab": [(30, 15), (0, 5), (-45, 30), (150, -10), (0, 0), (0, 0)],
}
BUILTIN_POSE_ORDER = ["standing", "walking", "running", "waving", "sitting", "fighting", "dab"]


def pose_array(pose) -> np.ndarray | None:
    """None | name | index | 6x2 / 12 numbers  ->  float32[12] (or None = standing)."""
    if pose is None:
        return None
    if isinstance(pose, str):
        pose = BUILTIN_POSES[pose]
    elif isinstance(pose, (int, np.integer)):
        pose = BUILTIN_POSES[BUILTIN_POSE_ORDER[int(pose)]]
    arr = np.asarray(pose, dtype=np.float32).reshape(12)
    return np.ascontiguousarray(arr)


def _lowbias32(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint64)
    m = np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x7FEB352D)) & m
    x ^= x >> np.uint64(15)
    x = (x * np.uint64(0x846CA68B)) & m
    x ^= x >> np.uint64(16)
    return x.astype(np.uint32)


def synth_skin(seed: int, kind: str = "64x64") -> np.ndarray:
    """Deterministic synthetic skin atlas, uint8 [H, 64, 4].

    h = lowbias32(seed*4096 + y*64 + x + 1); RGB = bytes 1,2,3 of h; alpha 255,
    except texels inside an outer-layer UV block get alpha 0 when h & 1.
    kind: "64x64" | "legacy" (64x32, outer block = head overlay only) |
          "slim" (64x64 with the columns a 3-px-arm layout leaves unused set transparent;
          the reference has no slim support and slices it as classic).
    """
    height = 32 if kind == "legacy" else 64
    ys, xs = np.mgrid[0:height, 0:64]
    h = _lowbias32((np.uint64(seed) * np.uint64(4096) + ys.astype(np.uint64) * np.uint64(64)
                    + xs.astype(np.uint64) + np.uint64(1)) & np.uint64(0xFFFFFFFF))
    img = np.empty((height, 64, 4), dtype=np.uint8)
    img[..., 0] = (h >> 8) & 0xFF
    img[..., 1] = (h >> 16) & 0xFF
    img[..., 2] = (h >> 24) & 0xFF
    img[..., 3] = 255
    outer = (xs >= 32) & (ys < 16)
    if height == 64:
        outer |= (ys >= 32) & (ys < 48)
        outer |= (ys >= 48) & ((xs < 16) | (xs >= 48))
    holes = outer & ((h & 1) == 1)
    img[holes, 3] = 0
    if kind == "slim":
        # columns unused by 3-px arms: right arm block x in [54,56), left arm block x in [46,48);
        # their outer-layer twins at (54..56, 32..48) and (62..64, 48..64)
        for (x0, x1, y0, y1) in ((54, 56, 16, 32), (46, 48, 48, 64), (54, 56, 32, 48), (62, 64, 48, 64)):
            img[y0:y1, x0:x1, 3] = 0
    elif kind not in ("64x64", "legacy"):
        raise ValueError(f"unknown skin kind {kind!r}")
    return img


@dataclass
class FlatScene:
    boxes: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=_abi.BOX_DTYPE))
    texels: np.ndarray = field(default_factory=lambda: np.zeros((0, 4), dtype=np.float32))
    light_pos: tuple = (0.0, 40.0, 30.0)
    light_color: tuple = (1.0, 1.0, 1.0, 1.0)
    light_radius: float = 3.0
    cam_pos: tuple = (0.0, 18.0, 50.0)
    cam_target: tuple = (0.0, 18.0, 0.0)
    cam_up: tuple = (0.0, 1.0, 0.0)
    cam_fov_deg: float = 60.0
    background: tuple = (0.2, 0.3, 0.5, 1.0)

    def as_c(self) -> _abi.McScene:
        """McScene whose pointers alias this object's arrays (keep `self` alive)."""
        self.boxes = np.ascontiguousarray(self.boxes, dtype=_abi.BOX_DTYPE)
        self.texels = np.ascontiguousarray(self.texels, dtype=np.float32).reshape(-1, 4)
        s = _abi.McScene()
        s.n_boxes = len(self.boxes)
        s.boxes = self.boxes.ctypes.data_as(C.POINTER(_abi.McBox)) if len(self.boxes) else None
        s.n_texels = len(self.texels)
        s.texels_rgba = self.texels.ctypes.data_as(C.POINTER(C.c_float)) if len(self.texels) else None
        s.light_pos[:] = list(self.light_pos)
        s.light_color[:] = list(self.light_color)
        s.light_radius = self.light_radius
        s.cam_pos[:] = list(self.cam_pos)
        s.cam_target[:] = list(self.cam_target)
        s.cam_up[:] = list(self.cam_up)
        s.cam_fov_deg = self.cam_fov_deg
        s.background[:] = list(self.background)
        s._keepalive = self  # noqa: SLF001
        return s

    @classmethod
    def from_c(cls, s: _abi.McScene) -> "FlatScene":
        boxes = np.zeros(s.n_boxes, dtype=_abi.BOX_DTYPE)
        if s.n_boxes:
            C.memmove(boxes.ctypes.data, s.boxes, boxes.nbytes)
        texels = np.zeros((s.n_texels, 4), dtype=np.float32)
        if s.n_texels:
            C.memmove(texels.ctypes.data, s.texels_rgba, texels.nbytes)
        return cls(boxes=boxes, texels=texels, light_pos=tuple(s.light_pos), light_color=tuple(s.light_color),
                   light_radius=s.light_radius, cam_pos=tuple(s.cam_pos), cam_target=tuple(s.cam_target),
                   cam_up=tuple(s.cam_up), cam_fov_deg=s.cam_fov_deg, background=tuple(s.background))

    def same_as(self, other: "FlatScene") -> bool:
        return (self.boxes.tobytes() == other.boxes.tobytes() and self.texels.tobytes() == other.texels.tobytes()
                and all(np.array_equal(np.float32(getattr(self, k)), np.float32(getattr(other, k))) for k in
                        ("light_pos", "light_color", "light_radius", "cam_pos", "cam_target", "cam_up",
                         "cam_fov_deg", "background")))


def make_box(lo, hi, face_textures, *, outer=False, pivot=(0, 0, 0), rot_x=0.0, rot_z=0.0, has_rotation=False,
             n_triangles=12):
    """One McBox record. face_textures: 6 x (texel_offset, width, height) in reference face order
    (-Z, +Z, +X, -X, +Y, -Y)."""
    b = np.zeros((), dtype=_abi.BOX_DTYPE)
    b["bounds_min"] = lo
    b["bounds_max"] = hi
    b["pivot"] = pivot
    b["rot_x_deg"] = rot_x
    b["rot_z_deg"] = rot_z
    b["has_rotation"] = int(has_rotation)
    b["is_outer_layer"] = int(outer)
    b["n_triangles"] = n_triangles
    b["face"] = np.asarray(face_textures, dtype=np.int32).reshape(6, 3)
    return b


def solid_box_scene(color=(1.0, 0.0, 0.0, 1.0), center=(0, 0, 0), size=(2, 2, 2), tex_wh=(4, 4), offset=0.0,
                    **scene_kwargs) -> FlatScene:
    """The fixture most reference unit tests use: one box, one solid texture on all faces
    (tests/test_intersection.cpp:7-17, MeshBuilder::buildBox mesh_builder.cpp:66-123)."""
    w, h = tex_wh
    texels = np.tile(np.asarray(color, dtype=np.float32), (w * h, 1))
    half = np.float32(size) / np.float32(2.0) + np.float32(offset)
    c = np.float32(center)
    box = make_box(c - half, c + half, [(0, w, h)] * 6, outer=offset > 0)
    return FlatScene(boxes=np.array([box]), texels=texels, **scene_kwargs)
