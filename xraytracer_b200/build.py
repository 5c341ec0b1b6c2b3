"""In-tree build of every native artefact (no JIT cache: the .so files travel with the repo snapshot).

  xraytracer_b200/csrc/libxrtgpu.so   CUDA kernels + C ABI          nvcc, sm_100a only
  xraytracer_b200/host/libxrthost.so  host C++ drop-in API          g++
  oracle/libxrtoracle.so              CPU restatement (checker)     g++     (tests / baselines only)
  oracle/_ref/libxrtref.so            compiled reference (checker)  g++     only where /root/reference exists
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "xraytracer_b200" / "csrc"
HOST = ROOT / "xraytracer_b200" / "host"
ORACLE = ROOT / "oracle"
REFERENCE_SRC = Path("/root/reference/Src")
CXX = os.environ.get("XRT_CXX", "/usr/bin/g++")
NVCC = os.environ.get("XRT_NVCC", "nvcc")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_COMMON = ARCH + ["-lineinfo", "-O3", "-std=c++17", "-ccbin", CXX, "-Xcompiler", "-fPIC,-fopenmp,-fvisibility=hidden",
                      f"-I{ROOT / 'include'}", f"-I{CSRC}"]


def _run(cmd, **kw):
    r = subprocess.run([str(c) for c in cmd], capture_output=True, text=True, **kw)
    if r.returncode != 0:
        raise RuntimeError("command failed: " + " ".join(str(c) for c in cmd) + "\n" + r.stdout + r.stderr)
    return r.stdout + r.stderr


def _stale(target: Path, sources) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(s).stat().st_mtime > t for s in sources)


def build_gpu(force=False, verbose=False) -> Path:
    """libxrtgpu.so. kernels_exact.cu is compiled with -fmad=false (parity), kernels_fast.cu with FMA."""
    out = CSRC / "libxrtgpu.so"
    headers = list(CSRC.glob("*.h")) + list(CSRC.glob("*.cuh")) + [ROOT / "include" / "xrtgpu.h"]
    # exact: no FMA contraction, IEEE div/sqrt (parity). fast: FMA + approximate div/sqrt/sin/cos/log/exp — it is an independent
    # Monte-Carlo estimate anyway (counter RNG) and is validated statistically against the oracle.
    fast_flags = os.environ.get("XRT_FAST_FLAGS", "-use_fast_math").split()
    units = [("kernels_exact.cu", ["-fmad=false"]), ("kernels_fast.cu", fast_flags), ("api.cu", []), ("multi.cu", []), ("lbvh.cu", []), ("gpu_build.cu", []), ("bvh.cpp", []), ("small_scene.cpp", [])]
    objs = []
    jobs = []
    for src, extra in units:
        obj = CSRC / (Path(src).stem + ".o")
        objs.append(obj)
        if force or _stale(obj, [CSRC / src] + headers):
            cmd = [NVCC] + NVCC_COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", CSRC / src, "-o", obj]
            jobs.append(cmd)
    if jobs:
        with ThreadPoolExecutor(max_workers=4) as ex:
            for log in ex.map(_run, jobs):
                if verbose:
                    print(log)
    if force or jobs or _stale(out, objs):
        _run([NVCC] + ARCH + ["-shared", "-ccbin", CXX, "-Xcompiler", "-fopenmp", "-o", out] + objs + ["-lgomp", "-ldl"])
    return out


def build_host(force=False) -> Path:
    out = HOST / "libxrthost.so"
    srcs = sorted(HOST.glob("*.cpp"))
    deps = srcs + list((ROOT / "include" / "xrt").glob("*.h")) + [ROOT / "include" / "xrtgpu.h"]
    if force or _stale(out, deps):
        _run([CXX, "-std=c++17", "-O2", "-fPIC", "-shared", "-Wall", "-fvisibility=hidden", "-fvisibility-inlines-hidden",
              "-Wl,-Bsymbolic", f"-I{ROOT / 'include'}"] + srcs + ["-o", out, f"-L{CSRC}", "-lxrtgpu",
                                                                 "-Wl,-rpath,$ORIGIN/../csrc"])
    return out


def build_oracle() -> list:
    """The checkers. `ref` only where the reference sources exist (never copied into the repo)."""
    built = []
    _run(["make", "-C", ORACLE, "port"])
    built.append(ORACLE / "libxrtoracle.so")
    if REFERENCE_SRC.exists():
        _run(["make", "-C", ORACLE, "-j8", "ref"])
        built.append(ORACLE / "_ref" / "libxrtref.so")
        built.append(ORACLE / "_ref" / "libxrtrefgpu.so")
    return built


def build_all(force=False, verbose=False):
    g = build_gpu(force, verbose)
    h = build_host(force)
    o = build_oracle()
    return [g, h] + o


if __name__ == "__main__":
    for p in build_all(force="--force" in sys.argv, verbose="-v" in sys.argv):
        print("built", p)
