#!/usr/bin/env python
"""Golden vectors for the libm calls inside the path: sinf / cosf / powf of THIS container's glibc
(2.39, FMA variants — the build the reference was run with for reference_vectors.npz), on the
inputs the path produces plus the branch boundaries of the algorithms.

    python tests/golden/make_libm_golden.py      ->  tests/golden/libm_vectors.npz

The device functions (csrc/dev_shade.cuh: sincos_ref, powf_ref) and their host models must reproduce
these bits on any host, FMA or not (tests/test_host_side.py, tests/test_parity_gpu.py)."""
import ctypes as C
import ctypes.util
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from tests.test_host_side import powf_test_values, sincos_test_angles  # noqa: E402


def main():
    if " fma " not in Path("/proc/cpuinfo").read_text():
        raise SystemExit("this CPU makes glibc pick its non-FMA sinf/cosf/powf: not the variant the goldens pin")
    libm = C.CDLL(ctypes.util.find_library("m") or "libm.so.6")
    for f in (libm.sinf, libm.cosf):
        f.restype, f.argtypes = C.c_float, [C.c_float]
    libm.powf.restype, libm.powf.argtypes = C.c_float, [C.c_float, C.c_float]
    angles = sincos_test_angles(30_000, seed=21)
    x, y = powf_test_values(30_000, seed=23)
    out = {
        "angles": angles,
        "sin": np.array([libm.sinf(float(a)) for a in angles], dtype=np.float32),
        "cos": np.array([libm.cosf(float(a)) for a in angles], dtype=np.float32),
        "pow_x": x, "pow_y": y,
        "pow": np.array([libm.powf(float(a), float(b)) for a, b in zip(x, y)], dtype=np.float32),
    }
    path = Path(__file__).resolve().parent / "libm_vectors.npz"
    np.savez_compressed(path, **out)
    print(path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
