"""In-tree build of libbhmel.so (nvcc, sm_100a only).  `python -m beatheritage_b200.build`.

The library is written next to this file so it travels with the repo snapshot to the GPU box
(built artefacts are git-ignored, not gpurun-ignored)."""
from __future__ import annotations

import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libbhmel.so")
LIB_STEM = os.path.join(PKG, "libbhstem.so")      # conv stem: its own library, its own header
# bounds-asserting debug builds (-DBHMEL_BOUNDS / -DBHSTEM_BOUNDS), used by tests/test_gpu_bounds.py only
LIB_BOUNDS = os.path.join(PKG, "libbhmel_bounds.so")
LIB_STEM_BOUNDS = os.path.join(PKG, "libbhstem_bounds.so")
GEN = os.path.join(CSRC, "fft32_gen.h")

NVCC_FLAGS = [
    *os.environ.get("BHMEL_EXTRA_NVCC", "").split(),
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v",
]


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def generate() -> str:
    gen_py = os.path.join(CSRC, "gen_fft32.py")
    if not _newer(GEN, [gen_py]):
        out = subprocess.run([sys.executable, gen_py], check=True, capture_output=True, text=True).stdout
        with open(GEN, "w") as f:
            f.write(out)
    mel_py, mel_h = os.path.join(CSRC, "gen_mel_static.py"), os.path.join(CSRC, "mel_static_gen.h")
    gen_args = os.environ.get("BHMEL_GEN_MEL_ARGS", "").split()   # A/B experiments only
    if gen_args or not _newer(mel_h, [mel_py, os.path.join(CSRC, "bhmel_fb_baked.h")]):
        subprocess.run([sys.executable, mel_py, "-o", mel_h, *gen_args], check=True)
    return GEN


def _nvcc(out: str, src: str, verbose: bool, extra: tuple = ()) -> None:
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    res = subprocess.run([nvcc, *extra, *NVCC_FLAGS, "-o", out, src], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed building {os.path.basename(out)}")


def build_stem(force: bool = False, verbose: bool = False) -> str:
    src = os.path.join(CSRC, "bhstem.cu")
    if force or not _newer(LIB_STEM, [src, os.path.join(os.path.dirname(PKG), "include", "bhstem.h")]):
        _nvcc(LIB_STEM, src, verbose)
    return LIB_STEM


def _mel_sources() -> list[str]:
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)
            if f.endswith((".cu", ".cuh", ".h", ".py")) and f != "bhstem.cu"]
    srcs.append(os.path.join(os.path.dirname(PKG), "include", "bhmel.h"))
    return srcs


def build_bounds(force: bool = False, verbose: bool = False) -> tuple[str, str]:
    """The bounds-asserting debug libraries (never loaded by the product)."""
    generate()
    stem_src = os.path.join(CSRC, "bhstem.cu")
    if force or not _newer(LIB_STEM_BOUNDS, [stem_src, os.path.join(os.path.dirname(PKG), "include", "bhstem.h")]):
        _nvcc(LIB_STEM_BOUNDS, stem_src, verbose, ("-DBHSTEM_BOUNDS",))
    if force or not _newer(LIB_BOUNDS, _mel_sources()):
        _nvcc(LIB_BOUNDS, os.path.join(CSRC, "bhmel.cu"), verbose, ("-DBHMEL_BOUNDS",))
    return LIB_BOUNDS, LIB_STEM_BOUNDS


def build(force: bool = False, verbose: bool = False, bounds: bool = False) -> str:
    """libbhstem.so, libbhmel.so and (bounds=True) their bounds-asserting debug twins, concurrently."""
    from concurrent.futures import ThreadPoolExecutor
    generate()
    jobs = [lambda: build_stem(force, verbose)]
    if force or not _newer(LIB, _mel_sources()):
        jobs.append(lambda: _nvcc(LIB, os.path.join(CSRC, "bhmel.cu"), verbose))
    if bounds:
        jobs.append(lambda: build_bounds(force, verbose))
    with ThreadPoolExecutor(max_workers=len(jobs)) as ex:
        for f in [ex.submit(j) for j in jobs]:
            f.result()
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True, bounds="--bounds" in sys.argv))
