"""pytest configuration. `-m "not gpu"` covers the oracle against the golden vectors / the compiled reference,
the host-side logic and the C-ABI surface; `-m gpu` are the parity tests proper (they need a B200)."""
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def native_libs():
    """Make sure the in-tree libraries exist (built by __graft_entry__.build(); rebuilt here if missing)."""
    from xraytracer_b200 import build, capi
    need = [capi.GPU_LIB, capi.HOST_LIB, capi.ORACLE_LIB]
    if not all(p.exists() for p in need):
        build.build_all()
    elif build.REFERENCE_SRC.exists() and not capi.REF_LIB.exists():
        build.build_oracle()
    return True


@pytest.fixture(scope="session")
def cornell():
    from xraytracer_b200 import scenes
    s = scenes.cornell_box("quad")
    return s, s.flatten()


def require_gpu():
    from xraytracer_b200 import capi
    if capi.gpu().xrtg_device_count() < 1:
        pytest.fail("GPU test selected but no CUDA device is visible (the render path has no CPU fallback)")


@pytest.fixture(autouse=True)
def guard_bands(monkeypatch):
    """Every workspace buffer of a GPU scene carries guard bands (csrc/scene_impl.h: DevBuf). After every render / trace call of
    a GPU test the guards are verified: a kernel that wrote past a queue fails the test that ran it (compute-sanitizer is
    closed on this pool). No-op for tests that never create a GpuScene."""
    from xraytracer_b200 import api

    def wrap(name):
        orig = getattr(api.GpuScene, name)

        def checked(self, *a, **kw):
            out = orig(self, *a, **kw)
            bad = self.check_guards()
            assert bad == 0, f"{name}: {bad} guard bytes around the workspace buffers were overwritten"
            return out
        monkeypatch.setattr(api.GpuScene, name, checked)
    for n in ("render", "render_device", "trace_primary", "trace_rays", "render_u8"):
        wrap(n)
    yield
