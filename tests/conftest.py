import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_report_header(config):
    """Which parity bar the GPU tests apply on this host (tests/test_parity_gpu.py::_flip_budget)."""
    try:
        from tests.test_host_side import _host_has_fma
        fma = _host_has_fma()
    except Exception as exc:  # noqa: BLE001
        return f"parity bar: unknown ({exc})"
    return ("parity bar: BIT-EXACT float images (this host's glibc selects its FMA sinf/cosf/powf variants, the ones the device "
            "functions restate)" if fma else
            "parity bar: 8-bit RGBA within +-1 LSB on >= 99.9 % of pixels, shadow-ray flips <= 2e-3 (non-FMA glibc host: "
            "sinf/cosf/powf differ in the last ulp from the device restatement)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU parity oracle (oracle/mcskin_oracle.c), built on demand."""
    from oracle.harness import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    """The unmodified reference behind oracle/_ref/libmcskin_ref.so; skipped when it was never built."""
    from oracle.harness import Reference, build_reference
    if os.environ.get("MCSKIN_SKIP_REF_BUILD") != "1":
        try:
            build_reference()
        except Exception as exc:  # noqa: BLE001
            print("reference build failed:", exc)
    ref = Reference.load()
    if ref is None:
        pytest.skip("oracle/_ref/libmcskin_ref.so not present")
    return ref


@pytest.fixture(scope="session")
def mclib():
    """The product library front-end; building it is __graft_entry__.build()'s job."""
    from minecraftskin_raytracer_b200 import build
    build.build()
    from minecraftskin_raytracer_b200 import lib
    return lib


@pytest.fixture(scope="session")
def gpu(mclib):
    if mclib.device_count() <= 0:
        pytest.fail("no CUDA device visible: -m gpu tests must run on a GPU box")
    return mclib


def pixel_report(a_f32: np.ndarray, b_f32: np.ndarray, quantize):
    """8-bit comparison of two float RGBA images: fraction of pixels with all channels within 1 LSB."""
    qa = quantize(a_f32).astype(np.int16)
    qb = quantize(b_f32).astype(np.int16)
    d = np.abs(qa - qb).max(axis=-1)
    return {
        "within1": float((d <= 1).mean()),
        "exact8": float((d == 0).mean()),
        "max_lsb": int(d.max()) if d.size else 0,
        "bitexact_f32": float((a_f32.view(np.uint32) == b_f32.view(np.uint32)).all(axis=-1).mean()) if d.size else 1.0,
    }
