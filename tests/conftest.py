import pathlib
import sys

import numpy as np
import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """The strict CPU oracle (test infrastructure; the product never loads it)."""
    from oracle.binding import Oracle

    return Oracle("strict")


@pytest.fixture(scope="session")
def scenes():
    sys.path.insert(0, str(ROOT / "tests" / "tools"))
    import gen_golden

    return {k: v for k, v in gen_golden.SCENES.items()}


@pytest.fixture(scope="session")
def ctx():
    """One rtcu context on cuda:0.  Fails loudly (no skip, no fallback) when the CUDA path is unavailable."""
    from rt_b200 import build
    from rt_b200.renderer import Context

    build.build_cuda()
    c = Context(0)
    yield c
    c.close()


@pytest.fixture
def knobs(ctx, monkeypatch):
    """Sets / clears RTCU_* experiment knobs for the session context: knobs(RTCU_BVH_DIRECT="0"), knobs(RTCU_BVH_DIRECT=None).
    The library reads its environment once in rtcu_create (never on the launch path), so every change is followed by
    rtcu_reload_env; the defaults are back when the test ends."""
    def set_knobs(**kv):
        for k, v in kv.items():
            if v is None:
                monkeypatch.delenv(k, raising=False)
            else:
                monkeypatch.setenv(k, str(v))
        ctx.reload_env()

    yield set_knobs
    monkeypatch.undo()
    ctx.reload_env()


def ulp_diff(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """distance in units of last place between two float32 arrays (same sign assumed where it matters)"""
    ia = a.astype(np.float32).view(np.int32).astype(np.int64)
    ib = b.astype(np.float32).view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, -(ia & 0x7FFFFFFF), ia)
    ib = np.where(ib < 0, -(ib & 0x7FFFFFFF), ib)
    return np.abs(ia - ib)


def unpack_rgba(rgba8: np.ndarray) -> np.ndarray:
    return np.stack([(rgba8 >> 24) & 255, (rgba8 >> 16) & 255, (rgba8 >> 8) & 255, rgba8 & 255], axis=-1).astype(np.int32)
