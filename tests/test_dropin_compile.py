"""Drop-in check of the C++ API: the reference's OWN example sources must compile UNCHANGED against include/xrt (same
header names, class names, constructor signatures), and our GPU example must link against libxrthost/libxrtgpu.
Compiled where /root/reference exists; nothing from it is copied."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
REF_EXAMPLES = Path("/root/reference/Src/examples")
CXX = "/usr/bin/g++"
FLAGS = ["-std=c++17", "-fsyntax-only", "-w", f"-I{ROOT / 'include' / 'xrt'}", f"-I{ROOT / 'include'}", f"-I{ROOT / 'tests' / 'stubs'}",
         "-include", str(ROOT / "oracle" / "shim" / "compat.h"), "-DXRT_WITH_OPENCV", '-DDATA_DIR="/tmp/"']


@pytest.mark.skipif(not REF_EXAMPLES.exists(), reason="/root/reference not present")
@pytest.mark.parametrize("example", ["cornellbox.cpp", "example.cpp", "vpt.cpp"])
def test_reference_examples_compile_unchanged_against_our_headers(example):
    r = subprocess.run([CXX] + FLAGS + [str(REF_EXAMPLES / example)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_our_gpu_example_builds_and_links(tmp_path):
    exe = tmp_path / "cornellbox_gpu"
    r = subprocess.run([CXX, "-std=c++17", "-O1", f"-I{ROOT / 'include'}", str(ROOT / "examples" / "cornellbox_gpu.cpp"), "-o", str(exe),
                        f"-L{ROOT / 'xraytracer_b200' / 'host'}", f"-L{ROOT / 'xraytracer_b200' / 'csrc'}", "-lxrthost", "-lxrtgpu",
                        f"-Wl,-rpath,{ROOT / 'xraytracer_b200' / 'host'}:{ROOT / 'xraytracer_b200' / 'csrc'}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    # without a GPU the program must fail loudly (exit code 2 + the ABI's error text), never render on the CPU
    from xraytracer_b200 import capi
    run = subprocess.run([str(exe), str(tmp_path / "out.ppm"), "32", "24", "2"], capture_output=True, text=True)
    if capi.gpu().xrtg_device_count() == 0:
        assert run.returncode == 2 and "no CUDA device" in run.stderr
        assert not (tmp_path / "out.ppm").exists()
    else:
        assert run.returncode == 0, run.stderr
        assert (tmp_path / "out.ppm").exists()
